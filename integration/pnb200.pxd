# Cython declarations of the C ABI in include/pnb200.h -- the file a PyNucleus maintainer would add as
# nl/PyNucleus_nl/pnb200.pxd.  Only what the dense path binds is declared.
from libc.stdint cimport int32_t, int64_t, uint8_t

cdef extern from "pnb200.h":
    ctypedef struct pnb_mesh_t:
        int32_t dim
        int32_t num_vertices
        int32_t num_cells
        const double *vertices
        const int32_t *cells
        const double *vol
        const double *h
        double diam
        int32_t num_bfacets
        const int32_t *bfacets
        int32_t num_blocks
        const int32_t *block_cell_ptr
        const int32_t *block_dof_ptr
        const int32_t *block_facet_ptr
    ctypedef struct pnb_dofmap_t:
        int32_t dofs_per_element
        int32_t num_dofs
        const int32_t *dofs
    ctypedef struct pnb_kernel_t:
        int32_t kernel_type
        int32_t dim
        double s
        double scaling
        double bscaling
        double singularity
        double bsingularity
        double horizon2
        double target_order
        double btarget_order
        int32_t order_num_dofs
        const uint8_t *cell_labels
        const uint8_t *bfacet_labels
        int32_t active_class
        uint8_t pair_class[16]
        uint8_t bpair_class[16]
        int32_t pair_orientation
        int32_t pair_filter
    ctypedef struct pnb_rule_t:
        int32_t n
        int32_t rows
        const double *bary
        const double *w
    ctypedef struct pnb_rules_t:
        pnb_rule_t identical
        pnb_rule_t edge
        pnb_rule_t vertex
        pnb_rule_t bedge
        pnb_rule_t bvertex
        int32_t max_order
        const pnb_rule_t *cell
        const pnb_rule_t *facet
    ctypedef struct pnb_problem:
        pass
    int PNB_ERR_ORDER
    const char *pnb_last_error()
    int pnb_device_count()
    int pnb_problem_create(const pnb_mesh_t *, const pnb_dofmap_t *, const pnb_kernel_t *, const pnb_rules_t *, int device,
                           pnb_problem **) nogil
    int pnb_problem_set_rules(pnb_problem *, const pnb_rules_t *) nogil
    void pnb_problem_destroy(pnb_problem *) nogil
    int pnb_max_order(pnb_problem *, int zero_exterior, int32_t *out) nogil
    int pnb_dense_assemble(pnb_problem *, int zero_exterior, int32_t row_begin, int32_t row_end, double *A, int64_t ld,
                           int a_on_device) nogil
    int pnb_dense_assemble_element(pnb_problem *, int polynomial_order, int dofs_per_element, int num_dofs, const int32_t *dofs,
                                   int zero_exterior, double *A, int64_t ld, int a_on_device) nogil
    int pnb_dense_assemble_element_tempered(pnb_problem *, double tempered, int polynomial_order, int dofs_per_element, int num_dofs,
                                            const int32_t *dofs, int zero_exterior, double *A, int64_t ld, int a_on_device) nogil
    int pnb_dense_assemble_element_smooth(pnb_problem *, int mode, double a, int bmode, double ba, int polynomial_order,
                                          int dofs_per_element, int num_dofs, const int32_t *dofs, int zero_exterior, double *A,
                                          int64_t ld, int a_on_device) nogil
    int pnb_dense_matvec(int device, const double *A, int64_t num_rows, int64_t num_cols, int64_t ld, const double *x,
                         double *y) nogil
