/* pnb200 -- C ABI of the B200-native nonlocal assembly library (libpnb200.so).
 *
 * Drop-in boundary for the Cython inner loops of the reference's nonlocal
 * operator assembly (sandialabs/PyNucleus, paths relative to the reference
 * root).  Plain pointers and sizes only; no Python or torch types.
 *
 * Conventions
 *   - every function returns 0 on success and a negative code on failure;
 *     pnb_last_error() returns a human readable message for the calling thread
 *   - all caller arrays are C-contiguous; the library never frees caller memory
 *   - INDEX = int32, REAL = double  (base/PyNucleus_base/myTypes64.pxd)
 *   - a pointer documented as "device" must point into CUDA memory of the
 *     problem's device (e.g. torch.Tensor.data_ptr()); "host" pointers are
 *     ordinary memory
 *   - there is no CPU fallback: without a CUDA device every compute entry point
 *     fails with PNB_ERR_NO_DEVICE
 */
#ifndef PNB200_H
#define PNB200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PNB_OK 0
#define PNB_ERR_NO_DEVICE (-1)
#define PNB_ERR_CUDA (-2)
#define PNB_ERR_ARG (-3)
#define PNB_ERR_UNSUPPORTED (-4)
#define PNB_ERR_ORDER (-5) /* a pair needs a regular rule of higher order than was supplied */

/* panel types, nl/PyNucleus_nl/panelTypes.pxi */
#define PNB_DISTANT 0
#define PNB_COMMON_VERTEX (-1)
#define PNB_COMMON_EDGE (-2)
#define PNB_COMMON_FACE (-3)
#define PNB_IGNORED (-6)

/* kernel types, nl/PyNucleus_nl/kernel_params.pxi:87-98 */
#define PNB_KERNEL_FRACTIONAL 0
#define PNB_KERNEL_INDICATOR 1      /* constant kernel C chi(|x-y| <= delta), kernelsCy.pyx:273-295 */
#define PNB_KERNEL_PERIDYNAMIC 2    /* C / |x-y| chi(|x-y| <= delta), kernelsCy.pyx:321-359 */

/* Simplicial mesh: the arrays nonlocalBuilder reads from `dm.mesh`
 * (nonlocalOperator_{SCALAR}.pxi:111-126,144-152: vertices, cells, volVector,
 * hVector, diam) plus the surface mesh of the zero-exterior loop
 * (nonlocalAssembly_{SCALAR}.pxi:1432-1436, mesh.get_surface_mesh()). */
typedef struct {
    int32_t dim;             /* 1 or 2 (= manifold dimension) */
    int32_t num_vertices;
    int32_t num_cells;
    const double *vertices;  /* host, num_vertices x dim */
    const int32_t *cells;    /* host, num_cells x (dim+1) */
    const double *vol;       /* host, num_cells  (mesh.volVector) */
    const double *h;         /* host, num_cells  (mesh.hVector)   */
    double diam;             /* mesh.diam; H0 = diam/sqrt(8), nonlocalOperator_{SCALAR}.pxi:435 */
    int32_t num_bfacets;
    const int32_t *bfacets;  /* host, num_bfacets x dim: boundary vertices (1D) / oriented boundary edges (2D) */
    /* Batch of independent sub-meshes in one problem (H2 near field: one block per near cluster pair,
     * assembleClusters, nonlocalAssembly_{SCALAR}.pxi:1663-1889).  num_blocks > 0: the cells, DoFs and boundary facets
     * of block k are the contiguous ranges [block_*_ptr[k], block_*_ptr[k+1]); cells of different blocks never
     * interact, every block has its own surface terms.  DoF ranges start at multiples of pnb_block_alignment()
     * (unused DoFs pad the range).  pnb_dense_assemble then writes, instead of one num_dofs x num_dofs matrix, the
     * dense operators of the blocks one after the other (block k: n_k x n_k doubles, row-major, n_k = its DoF range,
     * at offset sum_{j<k} n_j^2; device output only, ld is ignored).  num_blocks == 0: one mesh (default). */
    int32_t num_blocks;
    const int32_t *block_cell_ptr;   /* host, num_blocks+1 */
    const int32_t *block_dof_ptr;    /* host, num_blocks+1 */
    const int32_t *block_facet_ptr;  /* host, num_blocks+1 */
} pnb_mesh_t;

/* P1 DoFMap: dm.dofs, negative = boundary DoF (fem/PyNucleus_fem/DoFMaps.pyx:157-210) */
typedef struct {
    int32_t dofs_per_element; /* dim+1 (P1) */
    int32_t num_dofs;
    const int32_t *dofs;      /* host, num_cells x dofs_per_element */
} pnb_dofmap_t;

/* Kernel parameter block, the device-side equivalent of
 * kernel_params.pxi:14-28 for piecewise-constant symmetric kernels. */
typedef struct {
    int32_t kernel_type;   /* PNB_KERNEL_*: all are C |x-y|^singularity inside the horizon */
    int32_t dim;
    double s;              /* fS        */
    double scaling;        /* fSCALING of the interior kernel  C(d,s)          */
    double bscaling;       /* fSCALING of kernel.getBoundaryKernel() = C/s     */
    double singularity;    /* fSINGULARITY of the interior kernel, -d-2s       */
    double bsingularity;   /* of the boundary kernel, 1-d-2s                   */
    double horizon2;       /* fHORIZON2; +inf for the infinite horizon.  Finite: interaction domain = l2 ball
                            * (ball2_retriangulation, interactionDomains.pyx:866-965): remote pairs are skipped, pairs cut
                            * by the horizon are re-triangulated, no zero-exterior surface terms                */
    double target_order;   /* local_matrix.target_order                         */
    double btarget_order;  /* local_matrix_zeroExterior.target_order            */
    int32_t order_num_dofs; /* num_dofs entering getQuadOrder (local_matrix.num_dofs, fractionalLaplacian2D.pyx:629);
                             * 0 = dm.num_dofs.  Differs when two DoFMaps are combined
                             * (nonlocalAssembly_{SCALAR}.pxi:1366-1378: the local matrix keeps the first map's count) */
    /* Piecewise constant variable kernels (kernel.evalParams at the cell centres per cell pair,
     * nonlocalOperator_{SCALAR}.pxi:509-513; leftRightFractionalOrder, fractionalOrders.pyx:285-335): the parameters above
     * hold for the cell pairs whose label pair maps to active_class; the other pairs are skipped (they belong to the
     * problem instance of their own class; the operator is the sum over the classes).  cell_labels == NULL: constant kernel. */
    const uint8_t *cell_labels;    /* host, num_cells, values 0..3 */
    const uint8_t *bfacet_labels;  /* host, num_bfacets */
    int32_t active_class;
    uint8_t pair_class[16];        /* [label of the cell with the SMALLER index * 4 + label of the other cell] -> class of a
                                    * cell pair.  An unsymmetric table (or pair_orientation != 0) sends the problem to the
                                    * DoF-tile kernels; the cell-group kernels look the class up in either order */
    uint8_t bpair_class[16];       /* [cell label * 4 + facet label] -> class of a (cell, boundary facet) pair of the
                                    * zero-exterior surface terms (kernel.evalParams(cell centre, facet centre)); need not
                                    * be symmetric.  Unsymmetric piecewise orders s(x,y) != s(y,x) are assembled as sums of
                                    * such instances, see nonlocalBuilder._getDenseClasses */
    int32_t pair_orientation;      /* singular (touching) cell pairs: 0 = evaluated as (smaller cell index, larger), the
                                    * order of the reference's symmetric assembly loop; 1 = as (larger, smaller), the second
                                    * visit of its unsymmetric loop (nonlocalAssembly_{SCALAR}.pxi:1419-1428).  The singular
                                    * rules are not symmetric in their two cells: the results differ by the quadrature error */
    int32_t pair_filter;           /* 0 = all cell pairs of the class; 1 = only the touching (singular) ones: the orientation
                                    * of a pair only matters for those, so an unsymmetric order takes the bulk of the pairs from
                                    * ONE instance on the fast path and corrects the touching pairs with two cheap instances
                                    * (+1/2 orientation 1, -1/2 orientation 0), see nonlocalBuilder.setKernel */
} pnb_kernel_t;

/* One quadrature table: rows x n barycentric coordinates (x point first, then
 * y point) and n weights, exactly the `nodes`/`weights` of the reference's
 * quadratureRule objects (fem/PyNucleus_fem/quadrature.pxd:25-26). */
typedef struct {
    int32_t n;
    int32_t rows;
    const double *bary; /* host, rows x n */
    const double *w;    /* host, n */
} pnb_rule_t;

/* All tables of one problem.  Singular tables replace specialQuadRules[...]
 * (fractionalLaplacian2D.pyx:644-813,1255-1314; fractionalLaplacian1D.pyx:255-339,
 * 671-709); regular tables replace distantQuadRulesPtr[order]
 * (nonlocalOperator_{SCALAR}.pxi:549-600, 988-1020). */
typedef struct {
    pnb_rule_t identical;  /* COMMON_FACE (2D) / COMMON_EDGE (1D) */
    pnb_rule_t edge;       /* COMMON_EDGE (2D only) */
    pnb_rule_t vertex;     /* COMMON_VERTEX */
    pnb_rule_t bedge;      /* boundary COMMON_EDGE (2D only) */
    pnb_rule_t bvertex;    /* boundary COMMON_VERTEX */
    int32_t max_order;     /* regular tables are given for orders 1..max_order */
    const pnb_rule_t *cell;  /* host, max_order+1 entries, index = order: rule on a cell  */
    const pnb_rule_t *facet; /* host, max_order+1 entries: rule on a boundary facet        */
} pnb_rules_t;

typedef struct pnb_problem pnb_problem;

/* mesh.hVector, mesh.h, mesh.hmin (hdeltaCy, fem/PyNucleus_fem/meshCy.pyx:1654-1732): per cell the longest edge,
 * h_max / h_min the longest / SHORTEST edge of the mesh; edge lengths are sqrt(mydot(e,e)) with mydot = BLAS ddot
 * (base/PyNucleus_base/opt_true_blas.pxi:125-141), which accumulates with fused multiply-adds.  Host arithmetic,
 * no device needed.  vertices: num_vertices x dim, cells: num_cells x (dim+1), h: num_cells (out). */
int pnb_mesh_edge_lengths(int32_t dim, int32_t num_cells, const double *vertices, const int32_t *cells,
                          double *h, double *h_max, double *h_min);

const char *pnb_last_error(void);
int pnb_version(void);
int pnb_device_count(void);

/* Uploads mesh, DoFMap, kernel parameters and tables to `device`; the assembly
 * schedules (cell groups, unit lists, near pair list; DoF tiles) are built on the
 * first assembly.  Replaces nonlocalBuilder.__init__/setKernel
 * (nonlocalAssembly_{SCALAR}.pxi:879-975).  `rules` may carry max_order = 0
 * when only pnb_max_order / pnb_classify_pairs are used afterwards. */
int pnb_problem_create(const pnb_mesh_t *mesh, const pnb_dofmap_t *dm, const pnb_kernel_t *kernel,
                       const pnb_rules_t *rules, int device, pnb_problem **out);
/* Replaces the lazily grown distantQuadRules cache: (re)uploads the regular tables. */
int pnb_problem_set_rules(pnb_problem *p, const pnb_rules_t *rules);
void pnb_problem_destroy(pnb_problem *p);

/* Sparsity pattern of nonlocalBuilder.getDense(trySparsification=True) (nonlocalAssembly_{SCALAR}.pxi:1293-1332): byte
 * mask (device memory, num_dofs x num_dofs, leading dimension ld) with a one for every pair of DoFs of every cell pair
 * that getPanelType does not ignore (finite horizon: pairs within reach of each other). */
int pnb_sparsity_mask(pnb_problem *p, unsigned char *device_mask, int64_t ld);

/* Assembly path of whole 2D operators with an infinite horizon: 0 (default) the cell-group kernels, 1 the DoF-tile
 * kernels that also serve 1D problems, row blocks and finite horizons.  Both produce the same operator (summation
 * order differs); the parity tests compare them. */
int pnb_problem_set_path(pnb_problem *p, int path);

/* alignment of the DoF ranges of the blocks of a batched problem (pnb_mesh_t.num_blocks) */
int pnb_block_alignment(void);

/* Largest regular quadrature order any cell pair / cell-facet pair requests
 * (getQuadOrder, fractionalLaplacian2D.pyx:622-642,1226-1253;
 * fractionalLaplacian1D.pyx:234-253,644-669).  Lets the host build exactly the
 * tables that addQuadRule would have created on demand. */
int pnb_max_order(pnb_problem *p, int zero_exterior, int32_t *max_order_out);

/* getPanelType() for a list of cell pairs (nonlocalOperator_{SCALAR}.pxi:280-378,
 * 493-540).  pairs: host, npairs x 2.  Outputs (host): panel[npairs],
 * perm1/perm2[npairs x (dim+1)] (may be NULL).  boundary != 0: second index is a
 * boundary facet (local_matrix_zeroExterior). */
int pnb_classify_pairs(pnb_problem *p, int boundary, int64_t npairs, const int32_t *pairs,
                       int32_t *panel, int32_t *perm1, int32_t *perm2);

/* Histogram of getPanelType over all cell pairs c1<=c2: hist[3+panel] for
 * panel in [-3, 255].  hist: host, 259 entries. */
int pnb_panel_histogram(pnb_problem *p, int64_t *hist);

/* local_matrix.eval(contrib, panel) for a list of cell pairs
 * (fractionalLaplacian2D.pyx:823-891, fractionalLaplacian1D.pyx:349-407,
 * nonlocalOperator_{SCALAR}.pxi:722-789; boundary: :1022-1108, 2D :1324-1407,
 * 1D :719-783).  contrib: host, npairs x nloc with
 * nloc = (2*dpe)(2*dpe+1)/2 (interior) or dpe(dpe+1)/2 (boundary), the
 * reference's flattened upper-triangular layout.
 * path = 0: warp-cooperative evaluator; path = 1: thread-per-pair low-order
 * evaluator (regular pairs of order <= pnb_far_max_order() only). */
int pnb_local_matrices(pnb_problem *p, int boundary, int path, int64_t npairs, const int32_t *pairs,
                       int32_t *panel, double *contrib);
int pnb_far_max_order(void);

/* nonlocalBuilder.getDense() (nonlocalAssembly_{SCALAR}.pxi:1262-1473) for the
 * rows [row_begin, row_end) of the operator: A is (row_end-row_begin) x num_dofs
 * with leading dimension ld (in doubles).  a_on_device != 0: A is device
 * memory; else A is host memory and the copy back is part of the call.
 * The result is deterministic (no floating point atomics): bitwise
 * reproducible from run to run, and bitwise symmetric.  2D, whole operator: cell-group kernels (every cell pair
 * evaluated once, U + U^T); 1D and row ranges: DoF-tile kernels. */
int pnb_dense_assemble(pnb_problem *p, int zero_exterior, int32_t row_begin, int32_t row_end,
                       double *A, int64_t ld, int a_on_device);

/* 2D, several GPUs of one node (one problem instance per GPU).  Replaces the reference's split of the cell loop over MPI
 * ranks followed by an Allreduce of the whole N x N matrix (nonlocalAssembly_{SCALAR}.pxi:1280-1285, 1449-1450).
 * Part `part` of `nparts` owns a contiguous range of cell groups (cells ordered along a Hilbert curve; the ranges are
 * balanced by estimated work) and the rows of the dofs that first appear in its groups.  Every cell pair is evaluated
 * exactly once over all parts: a unit of two groups by the part of its row group or of its column group, alternating.
 * The unit block reaches the owners of its rows as row fragments stored into their staging buffers (peer memory over
 * NVLink: plain stores, no read-modify-write, no atomics); the owner sums the fragments of its rows in a fixed order.
 * Per GPU only num_rows x num_dofs entries and a staging buffer of about twice that size exist; no collective moves
 * matrix entries.  The caller exchanges the per-cell diagonal blocks (pnb_dense_cell_blocks_copy, num_cells x 6 doubles)
 * and one status word.
 *   1. pnb_dist_plan      partition, schedules and tables; returns num_rows and the staging size of this part
 *      pnb_dist_rows      global dof of every owned row (ascending)
 *   2. pnb_dist_eval      asynchronous; stage_ptrs[o] = staging buffer of part o as seen from this device (own
 *                         allocation, or peer memory opened with pnb_ipc_import)
 *   3. pnb_dist_status    synchronises the device; order_needed > 0: a pair needs a regular rule beyond the supplied
 *                         tables -- the caller takes the maximum over ALL parts, extends the tables on all of them
 *                         (pnb_problem_set_rules) and repeats step 2
 *   4. the caller sums the cell-block buffers over the parts, after which all parts are known to have finished step 2
 *   5. pnb_dist_apply     A_rows (device, num_rows x num_dofs): row k = global row rows[k] */
int pnb_dist_plan(pnb_problem *p, int32_t nparts, int32_t part, int32_t *num_rows, int64_t *staging_doubles);
int pnb_dist_rows(pnb_problem *p, int32_t *rows);
int pnb_dist_eval(pnb_problem *p, int zero_exterior, double *const *stage_ptrs);
int pnb_dist_status(pnb_problem *p, int32_t *order_needed);
int pnb_dist_apply(pnb_problem *p, int use_cell_blocks, double *A_rows, int64_t ld);
/* plain device allocations and their CUDA IPC handles (64 bytes), so that the staging buffers of the other processes of
 * the node can be written through peer memory */
int pnb_device_alloc(int device, int64_t bytes, void **dptr);
int pnb_device_free(int device, void *dptr);
int pnb_ipc_export(int device, void *dptr, unsigned char *handle);
int pnb_ipc_import(int device, const unsigned char *handle, void **dptr);
int pnb_ipc_close(int device, void *dptr);

/* Row-block (multi-GPU) form of pnb_dense_assemble: one problem instance per GPU assembles the rows
 * [row_begin, row_end) (multiples of pnb_row_granularity(), or num_dofs) into device memory A_rows of
 * (row_end-row_begin) x num_dofs.  Replaces the reference's cell-range split + Allreduce of the full matrix
 * (nonlocalAssembly_{SCALAR}.pxi:1280-1285, 1449-1450): every entry is written by exactly one GPU.
 *   1. pnb_dense_rows_begin   all pair integrals that touch the owned rows
 *   2. the caller sums the buffers returned by pnb_dense_cell_blocks over all row blocks (NCCL allreduce;
 *      supports are disjoint, so the sum is exact and order independent).  Not needed for a single block.
 * Used for 1D problems and explicit row ranges (DoF-tile kernels); 2D operators on several GPUs: pnb_dist_*.
 *   3. pnb_dense_rows_end     adds the cell-diagonal blocks to the owned rows */
int pnb_row_granularity(void);
int pnb_dense_rows_begin(pnb_problem *p, int zero_exterior, int32_t row_begin, int32_t row_end, double *A_rows, int64_t ld);
int pnb_dense_cell_blocks(pnb_problem *p, double **device_ptr, int64_t *count);
/* copy between that buffer and caller device memory of `count` doubles (to_problem != 0: caller -> problem) */
int pnb_dense_cell_blocks_copy(pnb_problem *p, double *device_buf, int to_problem);
int pnb_dense_rows_end(pnb_problem *p, int32_t row_begin, int32_t row_end, double *A_rows, int64_t ld);

/* Zero-exterior surface terms alone (the facet loop nonlocalAssembly_{SCALAR}.pxi:1430-1448 without the cell pairs):
 * per cell the upper triangle (row-major) of its symmetric (dim+1) x (dim+1) block, num_cells x (dim+1)(dim+2)/2
 * doubles written to HOST memory.  The H2 near field of the regional operator (zeroExterior=False) subtracts them
 * (assembleClusters, nonlocalAssembly_{SCALAR}.pxi:1889-1912). */
int pnb_boundary_cell_blocks(pnb_problem *p, double *host_out);

/* Counters of the last pnb_dense_assemble call: [0] evaluated cell pairs (2D whole-operator path: every pair
 * once; DoF-tile path: with tile-halo redundancy), [1] distinct cell pairs c1<=c2 that are not skipped,
 * [2] kernel launches, [3] pairs of the near evaluator, [4] pairs of the uniform order-2 units, [5..] reserved.
 * stats: host, 8 entries. */
int pnb_dense_stats(pnb_problem *p, int64_t *stats);
/* device time in ms of the phases of the last assembly:
 * [0] pair kernels (2D: near evaluator + unit kernels + symmetrisation; 1D / row ranges: tile kernels),
 * [1] boundary kernel, [2] reduce+scatter of the cell-diagonal blocks, [3] total */
int pnb_dense_timings(pnb_problem *p, double *ms);
/* device milliseconds of the last 2D assembly per kernel: [0] uniform order-2 units, [1] near pair list (first
 * assembly only) + near evaluator, [2] all other units, [3] symmetrisation */
int pnb_dense_kernel_timings(pnb_problem *p, double *ms);

/* H2 far-field kernel blocks, assembleFarFieldInteractions (nl/PyNucleus_nl/clusterMethodCy.pyx:2153-2238):
 * for each admissible cluster pair b an m1^d x m2^d block  -2 gamma(xi_i, xi_j)  at the tensor Chebyshev
 * nodes of the two cluster boxes (boxes: nblk x d x 2 = [lo, hi] per axis; multi-index with the last
 * dimension fastest).  eta / eta_ptr: the 1D Chebyshev nodes cos((2(m-p)-1)pi/(2m)), p = 0..m-1, for every
 * m <= max_m, stored back to back (eta_ptr[m] = start of the m nodes; eta_ptr has max_m+2 entries) -- computed
 * by the caller exactly as the reference does (np.cos) so that the node coordinates are bit-identical.
 * offsets: nblk+1 prefix sums of the block sizes; out: host, offsets[nblk] doubles.  All inputs host memory. */
int pnb_farfield_blocks(pnb_problem *p, int64_t nblk, const double *boxes1, const double *boxes2,
                        const int32_t *m1, const int32_t *m2, int32_t max_m, const double *eta,
                        const int32_t *eta_ptr, const int64_t *offsets, double *out);

/* Dense operator for a DoFMap that is not P1 (P0 and P2 on intervals and triangles, P3 on intervals; P1 accepted for
 * cross-checks): the
 * reference runs the same assembly loop with (2 dpe)(2 dpe + 1)/2 local entries built from the DoFMap's shape functions
 * (nonlocalAssembly_{SCALAR}.pxi:1386-1448 with fractionalLaplacian2D.pyx:644-891).  `p` carries the mesh, the kernel and
 * the tables (create it with the vertex dofs of the map as a P1 table and kernel.order_num_dofs = num_dofs); `dofs` is the
 * element's cell -> dof table (host, num_cells x dofs_per_element, the reference's local order: vertices, then edges
 * (0,1), (1,2), (0,2); 1D: vertices, then the cell; P0: the cell).  polynomial_order 0 expects singular tables built for
 * discontinuous elements (no cancellation across elements, fractionalLaplacian2D.pyx:595-600).  One warp owns one row of the operator (no atomics, bitwise
 * reproducible); infinite horizon, constant kernels.  A: num_dofs x num_dofs, row-major, leading dimension ld; device
 * memory if a_on_device != 0, else host memory (assembled on the device and copied back). */
int pnb_dense_assemble_element(pnb_problem *p, int polynomial_order, int dofs_per_element, int num_dofs, const int32_t *dofs,
                               int zero_exterior, double *A, int64_t ld, int a_on_device);
/* The same for TEMPERED fractional kernels gamma(x,y) = C |x-y|^(-d-2s) exp(-tempered |x-y|) (temperedFracKernelInfinite*,
 * kernelsCy.pyx:186-213; any element incl. P1).  `p` carries the power law with the tempered scaling constant
 * (constantFractionalLaplacianScaling, kernelNormalization.pyx:84-88) as kernel.scaling; the surface terms use the boundary
 * kernel of `p` as it is: the reference's getBoundaryKernel (kernelsCy.pyx:1982-2027) does not hand the tempering on, so its
 * zero-exterior terms are the untempered power law with the tempered constant / s.  tempered = 0: the plain kernel. */
int pnb_dense_assemble_element_tempered(pnb_problem *p, double tempered, int polynomial_order, int dofs_per_element,
                                        int num_dofs, const int32_t *dofs, int zero_exterior, double *A, int64_t ld,
                                        int a_on_device);
/* The general form: kernels gamma(x,y) = C |x-y|^e f(|x-y|) whose power law (C, e) is the interior kernel of `p` and whose
 * boundary kernel is the boundary power law of `p` times another smooth factor.  Serves the integrable kernels with infinite
 * horizon of the reference's driver tests (tests/test_drivers_intFracLapl.py:42-43), with e = 0:
 *   Gaussian    C exp(-a r^2), a = 1/(2 variance^d) (gaussianKernel*, kernelsCy.pyx:388-415; Kernel.__init__ :690-695);
 *               boundary 1D  C sqrt(pi/a) erfc(sqrt(a) r)  (:418-430),  2D  C/a exp(-a r^2)/r  (:433-445)
 *   exponential C exp(-a r), boundary 2C/a exp(-a r)  (:448-477)
 * Create `p` as a fractional problem with singularity = bsingularity = 0, scaling = C, bscaling = the constant of the
 * boundary form, infinite horizon. */
#define PNB_SMOOTH_NONE 0
#define PNB_SMOOTH_EXP_R 1          /* exp(-a |x-y|)   */
#define PNB_SMOOTH_EXP_R2 2         /* exp(-a |x-y|^2) */
#define PNB_SMOOTH_ERFC_R 3         /* erfc(sqrt(a) |x-y|): boundary factor only */
#define PNB_SMOOTH_EXP_R2_OVER_R 4  /* exp(-a |x-y|^2) / |x-y|: boundary factor only (2D) */
int pnb_dense_assemble_element_smooth(pnb_problem *p, int mode, double a, int bmode, double ba, int polynomial_order,
                                      int dofs_per_element, int num_dofs, const int32_t *dofs, int zero_exterior, double *A,
                                      int64_t ld, int a_on_device);

/* Dense operator for a fractional order that VARIES INSIDE A CELL: s(x, y) = sFun(x), kernel.piecewise == False
 * (singleVariableUnsymmetricFractionalOrder, fractionalOrders.pyx:153-183; smoothedLeftRightFractionalOrder :641-645 is the
 * driver's `--s twoDomainNonSym(sl,sr)`).  The reference then assembles with the unsymmetric local matrices
 * fractionalLaplacian{1,2}D_nonsym over both orientations of every cell pair (nonlocalAssembly_{SCALAR}.pxi:1412-1428),
 * re-evaluates order, scaling constant and kernel at every quadrature node (updateAndEvalFractional, kernelsCy.pyx:596-622;
 * variableFractionalLaplacianScaling.evalPtr, kernelNormalization.pyx:421-440) and uses, per cell pair, the singularity
 * -d - 2 max(s) over the centres and vertices of both cells (evalParamsOnSimplices, kernelsCy.pyx:1826-1850) for the
 * regular order and for a singular rule of its own (getNearQuadRule).  The caller evaluates s at the centres and vertices
 * and hands over the distinct maxima with one set of singular tables per value. */
#define PNB_ORDERFUN_CONST 0
#define PNB_ORDERFUN_SMOOTHSTEP 1          /* smoothStep, fractionalOrders.pyx:389-416 (first coordinate) */
#define PNB_ORDERFUN_LINEARSTEP 2          /* linearStep, :447-470 */
#define PNB_ORDERFUN_SMOOTHSTEP_RADIAL 3   /* smoothStepRadial, :497-535 (interface = radius) */
#define PNB_ORDERFUN_FE 4                  /* feFractionalOrder, :660-668: a P1 function on the assembly mesh (vertex_values);
                                            * sl / sr = bounds of the order (values equal to them use the power tables) */
typedef struct {
    int32_t fun;                   /* PNB_ORDERFUN_* */
    double sl, sr, r, slope, interface;
    int32_t num_values;            /* distinct values of max(s) over a cell (centre and vertices) or boundary facet */
    const double *values;          /* host, num_values, ascending */
    const int32_t *cell_value;     /* host, num_cells: index into values */
    const int32_t *bfacet_value;   /* host, num_bfacets */
    /* singular tables for the singularities -d - 2 values[k] (boundary: 1 - d - 2 values[k]); host, num_values entries each;
     * edge / bedge: 2D only (may be NULL in 1D) */
    const pnb_rule_t *identical, *edge, *vertex, *bedge, *bvertex;
    const double *vertex_values;   /* host, num_vertices: PNB_ORDERFUN_FE only (else NULL) */
} pnb_varorder_t;
/* `p`: mesh, regular tables and the quadrature-order constants (target orders, order_num_dofs) of the local matrices; its
 * own kernel parameters and singular tables are not used.  Elements as in pnb_dense_assemble_element.  One warp owns one
 * row (no atomics, bitwise reproducible); normalised kernels with infinite horizon. */
int pnb_dense_assemble_varorder(pnb_problem *p, const pnb_varorder_t *order, int polynomial_order, int dofs_per_element,
                                int num_dofs, const int32_t *dofs, int zero_exterior, double *A, int64_t ld, int a_on_device);

/* Several GPUs for the row-owner kernels (pnb_dense_assemble_element / _tempered / _smooth / _varorder): one problem instance per
 * GPU assembles the rows of one part.  The rows, sorted by descending number of cells around their dof, are dealt to the parts
 * in turn; a row is complete on its owner (one warp per row, no exchange of matrix entries, no collective).  After
 * pnb_problem_set_row_part(p, part, nparts) the assembly calls write num_rows x num_dofs entries (device memory only):
 * row k of the output = global row rows[k] of pnb_element_rows (ascending).  nparts = 1 restores the whole operator.
 * pnb_element_rows is host arithmetic (rows may be NULL to query the count). */
int pnb_problem_set_row_part(pnb_problem *p, int32_t part, int32_t nparts);
int pnb_element_rows(pnb_problem *p, int dofs_per_element, int num_dofs, const int32_t *dofs, int32_t part, int32_t nparts,
                     int32_t *rows, int32_t *num_rows);
/* the same without a problem instance (no device needed): every rank can list the rows of all parts */
int pnb_element_rows_host(int32_t num_cells, int dofs_per_element, int num_dofs, const int32_t *dofs, int32_t part,
                          int32_t nparts, int32_t *rows, int32_t *num_rows);

/* ---- H2 operator on the device -------------------------------------------------------------------------
 * Replaces H2Matrix.matvec (nl/PyNucleus_nl/clusterMethodCy.pyx:2269-2295) with its upwardPass / downwardPass
 * (:1093-1176) and tree_node.enterLeafValues (:1205-1325).  The caller describes the cluster tree node by node
 * (ids = positions in the arrays; the root has parent -1), hands over the transfer operators, the far-field kernel blocks
 * (pnb_farfield_blocks) and the near field as a CSR matrix; the leaf moments are either supplied or computed on the
 * device from the mesh.  All arrays are host memory unless stated; the library copies what it keeps. */
typedef struct pnb_h2 pnb_h2;
typedef struct {
    int32_t dim, num_dofs, num_nodes;
    const int32_t *coef_ptr;       /* num_nodes+1: node n owns coef_ptr[n+1]-coef_ptr[n] = m_n^dim coefficients */
    const int32_t *parent;         /* num_nodes, -1 for the root */
    const int32_t *level;          /* num_nodes, root = 0 */
    /* leaves */
    int32_t num_leaves;
    const int32_t *leaf_node;      /* num_leaves: node id */
    const int32_t *leaf_dof_ptr;   /* num_leaves+1 */
    const int32_t *leaf_dofs;      /* dofs of the leaves (every dof in exactly one leaf) */
    const double *leaf_values;     /* V[dof][alpha] of all leaves back to back (row-major), or NULL: computed on the
                                    * device from the fields below */
    const int32_t *leaf_cell_ptr;  /* num_leaves+1: cells around the dofs of a leaf, ascending */
    const int32_t *leaf_cells;
    const int32_t *leaf_cell_pos;  /* (dim+1) per listed cell: position of its dofs in the leaf's dof list, -1 = elsewhere */
    const double *leaf_boxes;      /* num_leaves x dim x 2 */
    const int32_t *leaf_orders;    /* num_leaves: interpolation order m */
    int32_t num_vertices, num_cells;
    const double *vertices;        /* num_vertices x dim */
    const int32_t *cells;          /* num_cells x (dim+1) */
    const double *vol;             /* num_cells */
    int32_t max_m;                 /* largest interpolation order */
    const int32_t *rule_n;         /* max_m+1: nodes of the rule of order m+2 on the reference simplex used by leaves of
                                    * interpolation order m (0 where no leaf has that order) */
    const int64_t *rule_bary_ptr;  /* max_m+1: start of the (dim+1) x rule_n[m] barycentric coordinates in rule_bary */
    const int64_t *rule_w_ptr;     /* max_m+1: start of the rule_n[m] weights in rule_w */
    const double *rule_bary;
    const double *rule_w;
    int64_t rule_bary_size, rule_w_size;
    const double *eta;             /* 1D Chebyshev nodes as in pnb_farfield_blocks */
    const int32_t *eta_ptr;        /* max_m+2 */
    /* transfer operators: T[m_parent^dim][m_node^dim] (row-major) at transfer_ptr[n]; -1 for the root */
    const int64_t *transfer_ptr;
    const double *transfer;
    int64_t transfer_size;
    /* admissible pairs (n1, n2) with their blocks K[m1^dim][m2^dim] at far_ptr[k] */
    int32_t num_far;
    const int32_t *far_n1, *far_n2;
    const int64_t *far_ptr;
    const double *far_blocks;
    int64_t far_size;
    /* near field, CSR, DEVICE memory (kept by the caller while the handle lives is NOT required: copied) */
    const int64_t *near_indptr;    /* device, num_dofs+1, or NULL */
    const int64_t *near_indices;   /* device */
    const double *near_data;       /* device */
} pnb_h2_desc_t;

int pnb_h2_create(int device, const pnb_h2_desc_t *desc, pnb_h2 **out);
/* leaf moments as computed or supplied: host buffer of the size of all V blocks */
int pnb_h2_leaf_values(pnb_h2 *h, double *out);
/* y = H x; x, y device memory, stream a cudaStream_t.  far_only != 0: without the near field */
int pnb_h2_matvec(pnb_h2 *h, const double *x, double *y, int far_only, void *stream);
int pnb_h2_destroy(pnb_h2 *h);

/* Dense_LinearOperator.matvec (base/PyNucleus_base/DenseLinearOperator_{SCALAR}.pxi:14-18
 * -> dgemv, opt_true_blas.pxi:159): y = A x for a row block.  All pointers are
 * device memory on `device`; stream is a cudaStream_t (0 = default stream). */
int pnb_dense_matvec(int device, const double *A, int64_t num_rows, int64_t num_cols, int64_t ld,
                     const double *x, double *y, void *stream);

/* ---- BLAS-1 of the device CG loop (cg_solver.solve, base/PyNucleus_base/solvers.pyx:364-445), fused; all vectors and the
 * workspace are device memory, stream a cudaStream_t.  The workspace holds the scalars of the iteration: work[0] = <r,z> of
 * the previous iteration, work[1] = <p,Ap>, work[2] = <r,z>, work[3] = <r,r>; pnb_krylov_workspace_doubles() doubles.
 * Reductions have a fixed shape: bitwise reproducible. */
int pnb_krylov_workspace_doubles(void);
int pnb_krylov_dot(int device, int64_t n, const double *a, const double *b, double *work, double *out, void *stream);
/* alpha = work[0] / <p,Ap>;  x += alpha p;  r -= alpha Ap;  z = Minv .* r (Minv == NULL: z = r, z may alias r);
 * work[2] = <r,z>, work[3] = <r,r> */
int pnb_krylov_cg_update(int device, int64_t n, const double *p, const double *Ap, const double *Minv, double *x, double *r,
                         double *z, double *work, void *stream);
/* p = z + (work[2] / work[0]) p;  work[0] = work[2] */
int pnb_krylov_cg_direction(int device, int64_t n, const double *z, double *p, double *work, void *stream);

/* FP64 FMA throughput microbenchmark (roofline denominator): returns TFLOP/s */
int pnb_fp64_peak(int device, double *tflops);

/* Device buffers of destroyed problems are cached for the next problem (the reference allocates its
 * matrices per call, nonlocalAssembly_{SCALAR}.pxi:1286-1290; cudaMalloc of GB-sized buffers is slow).
 * Returns the cached buffers to the driver.  PNB_POOL_LIMIT_GB bounds the cache (default 48). */
int pnb_release_cached_memory(void);

#ifdef __cplusplus
}
#endif
#endif
