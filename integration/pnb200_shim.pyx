# cython: language_level=3, boundscheck=False, wraparound=False
"""Thin Cython shim between PyNucleus' own objects and libpnb200 (C ABI, include/pnb200.h).

This is the file a PyNucleus maintainer would add next to nl/PyNucleus_nl/nonlocalAssembly.pyx: it takes the reference's
`nonlocalBuilder` (its mesh, DoFMap, kernel and the quadrature tables its local matrices already hold), fills the C
parameter blocks, and replaces the Cython cell-pair loop of `getDense` (nonlocalAssembly_{SCALAR}.pxi:1386-1448) by
one call into the CUDA library.  The result is written straight into the `data` array of a reference
`Dense_LinearOperator`.  Nothing of this repository's Python package is imported: the shim talks to the .so only.

    from pnb200_shim import nonlocalBuilderB200          # subclass of PyNucleus_nl.nonlocalBuilder
    A = nonlocalBuilderB200(dm, kernel, params).getDense()

Tempered fractional kernels and the Gaussian / exponential kernels on the full space go to the row-owner kernel with their smooth
factors (pnb_dense_assemble_element_smooth).  Unsupported configurations (non-symmetric or variable
kernels, two DoFMaps, P3 on triangles, vector-valued kernels) fall through
to the reference's own getDense, so the subclass is a drop-in.
"""
import numpy as np
cimport numpy as np
from libc.stdint cimport int32_t, int64_t
from libc.stdlib cimport malloc, free
from libc.string cimport memset
from pnb200 cimport *

np.import_array()


class PNB200Error(RuntimeError):
    pass


cdef object _last_error():
    return pnb_last_error().decode('utf-8', 'replace')


cdef void _fill_rule(pnb_rule_t *r, double[:, ::1] nodes, double[::1] weights):
    r.n = <int32_t>nodes.shape[1]
    r.rows = <int32_t>nodes.shape[0]
    r.bary = &nodes[0, 0]
    r.w = &weights[0]


def _regular_rules(int dim, int max_order):
    """the reference's regular rules of orders 1..max_order (nonlocalOperator_{SCALAR}.pxi:561, 998-999)"""
    from PyNucleus_fem.quadrature import simplexXiaoGimbutas, simplexDuffyTransformation
    cell = [None]
    facet = [None]
    for p in range(1, max_order+1):
        qr = simplexXiaoGimbutas(p, dim)
        cell.append((np.ascontiguousarray(qr.nodes, dtype=np.float64), np.ascontiguousarray(qr.weights, dtype=np.float64)))
        if dim == 1:
            facet.append((np.ones((1, 1)), np.ones(1)))
        else:
            qf = simplexDuffyTransformation(p, dim, dim-1)
            facet.append((np.ascontiguousarray(qf.nodes, dtype=np.float64), np.ascontiguousarray(qf.weights, dtype=np.float64)))
    return cell, facet


def _tempered(kernel):
    """tempering rate of a fractional kernel (temperedFracKernelInfinite*, kernelsCy.pyx:186-213), 0 for all others"""
    if int(kernel.kernelType) != 0:
        return 0.
    return float(getattr(kernel, 'temperedValue', 0.) or 0.)


def _smooth(kernel):
    """(mode, a, boundary mode, boundary a, constant of the boundary form) of a Gaussian / exponential kernel on the full space
    (gaussianKernel*, exponentialKernel*: kernelsCy.pyx:388-477; fEXPONENTINVERSE: Kernel.__init__ :690-697), None otherwise"""
    kt = int(kernel.kernelType)
    if kt not in (3, 8) or kernel.finiteHorizon:
        return None
    C = kernel.scalingValue
    if kt == 8:
        a = float(kernel.getKernelParam('exponentialRate'))
        return 1, a, 1, a, 2.0*C/a
    a = 0.5/float(kernel.getKernelParam('variance'))**kernel.dim
    if kernel.dim == 1:
        return 2, a, 3, a, C*np.sqrt(np.pi/a)
    return 2, a, 4, a, C/a


def supported(builder):
    """configurations the accelerated path covers (everything else stays with the reference's Cython loops)"""
    from PyNucleus_fem.DoFMaps import P0_DoFMap, P1_DoFMap, P2_DoFMap, P3_DoFMap
    k = builder.kernel
    if int(k.kernelType) in (3, 8):
        # Gaussian (1D / 2D) and exponential (1D) kernels on the full space, P1 / P2 (row-owner kernel with smooth factors)
        return (not k.finiteHorizon and builder.dm2 is None and isinstance(builder.dm, (P1_DoFMap, P2_DoFMap)) and k.symmetric
                and not k.variable and k.valueSize == 1 and builder.dm.mesh.dim in (1, 2) and builder.dm.mesh.manifold_dim == builder.dm.mesh.dim
                and (builder.comm is None or builder.comm.size == 1) and not k.complement)
    if _tempered(k) != 0. and k.finiteHorizon:
        return False        # tempered kernels: infinite horizon (row-owner kernel)
    if isinstance(builder.dm, (P0_DoFMap, P2_DoFMap, P3_DoFMap)) and (k.finiteHorizon or int(k.kernelType) != 0):
        return False        # P0 / P2 / P3: fractional kernels with infinite horizon (row-owner kernel)
    if isinstance(builder.dm, P3_DoFMap) and builder.dm.mesh.dim != 1:
        return False        # cubic elements: intervals only
    return (builder.dm2 is None and isinstance(builder.dm, (P0_DoFMap, P1_DoFMap, P2_DoFMap, P3_DoFMap)) and k.symmetric and not k.variable
            and k.valueSize == 1 and builder.dm.mesh.dim in (1, 2) and builder.dm.mesh.manifold_dim == builder.dm.mesh.dim
            and (builder.comm is None or builder.comm.size == 1) and int(k.kernelType) in (0, 1, 2)
            and not k.complement and (int(k.kernelType) == 0 or k.finiteHorizon))


def getDense(builder, zeroExterior=True, int device=0, int max_regular_order=24):
    """nonlocalBuilder.getDense() (nonlocalAssembly_{SCALAR}.pxi:1262-1473) on the GPU through the C ABI.
    `zeroExterior`: the constructor argument of the builder (the attribute itself is private to the cdef class)"""
    from PyNucleus_base.linear_operators import Dense_LinearOperator
    cdef:
        pnb_mesh_t m
        pnb_dofmap_t d
        pnb_kernel_t k
        pnb_rules_t r
        pnb_problem *prob = NULL
        pnb_rule_t *cell = NULL
        pnb_rule_t *facet = NULL
        int rc, o, dim, N, zero_exterior, porder, dpe
        int32_t[:, ::1] edofs
        int32_t need = 0
        double tempered, sa = 0., sba = 0.
        int smode = 0, sbmode = 0
        double[:, ::1] vertices, nodes
        double[::1] vol, h, weights
        int32_t[:, ::1] cells, dofs, bfacets
        double[:, ::1] data
    if not supported(builder):
        raise NotImplementedError('configuration outside the accelerated path')
    mesh = builder.dm.mesh
    dm = builder.dm
    kernel = builder.kernel
    # tempered kernels: the power law with the tempered constant in the problem, the exponential factor in the row-owner
    # kernel; the surface terms stay as they are (the reference's boundary kernel is not tempered, kernelsCy.pyx:2011-2020)
    tempered = _tempered(kernel)
    smooth = _smooth(kernel)
    if tempered != 0.:
        smode, sa = 1, tempered
    lm = builder.local_matrix
    lmb = builder.local_matrix_zeroExterior
    dim = mesh.dim
    N = dm.num_dofs
    vertices = np.ascontiguousarray(mesh.vertices, dtype=np.float64)
    cells = np.ascontiguousarray(mesh.cells, dtype=np.int32)
    # P2: the problem is created over the vertex dofs (first dim+1 columns of the table); the element's whole table goes
    # to pnb_dense_assemble_element
    porder = dm.polynomialOrder
    dpe = dm.dofs_per_element
    edofs = np.ascontiguousarray(dm.dofs, dtype=np.int32)
    if porder == 0:
        # P0 has no vertex dofs: any valid P1-shaped table will do (the element kernel never reads it)
        tab = np.full((mesh.num_cells, dim+1), -1, dtype=np.int32)
        tab[:, 0] = np.arange(mesh.num_cells)
        dofs = np.ascontiguousarray(tab)
    else:
        dofs = np.ascontiguousarray(np.asarray(dm.dofs)[:, :dim+1], dtype=np.int32)
    vol = np.ascontiguousarray(mesh.volVector, dtype=np.float64)
    h = np.ascontiguousarray(mesh.hVector, dtype=np.float64)
    if dim == 2:
        bfacets = np.ascontiguousarray(mesh.boundaryEdges, dtype=np.int32).reshape(-1, 2)
    else:
        bfacets = np.ascontiguousarray(mesh.boundaryVertices, dtype=np.int32).reshape(-1, 1)
    # every block is zeroed first: fields this shim does not set keep their neutral value when the header grows
    memset(&m, 0, sizeof(m))
    memset(&d, 0, sizeof(d))
    memset(&k, 0, sizeof(k))
    memset(&r, 0, sizeof(r))
    m.dim = dim
    m.num_vertices = mesh.num_vertices
    m.num_cells = mesh.num_cells
    m.vertices = &vertices[0, 0]
    m.cells = &cells[0, 0]
    m.vol = &vol[0]
    m.h = &h[0]
    m.diam = mesh.diam
    m.num_bfacets = <int32_t>bfacets.shape[0]
    m.bfacets = &bfacets[0, 0] if bfacets.shape[0] > 0 else NULL
    d.dofs_per_element = dim+1
    d.num_dofs = N
    d.dofs = &dofs[0, 0]
    k.kernel_type = int(kernel.kernelType)
    k.dim = kernel.dim
    k.s = kernel.sValue if k.kernel_type == 0 else 0.
    k.scaling = kernel.scalingValue
    k.singularity = kernel.singularityValue
    k.horizon2 = kernel.horizonValue**2 if kernel.finiteHorizon else np.inf
    k.target_order = lm.target_order
    k.order_num_dofs = N
    if smooth is not None:
        # the library sees the power law C |x-y|^0 (a fractional problem with singularity 0) and the smooth factors
        smode, sa, sbmode, sba, bconst = smooth
        k.kernel_type = 0
        k.s = -0.5*kernel.dim
        k.bscaling = bconst
        k.bsingularity = lmb.kernel.singularityValue
        k.btarget_order = lmb.target_order
    elif k.kernel_type == 0:
        k.bscaling = lmb.kernel.scalingValue
        k.bsingularity = lmb.kernel.singularityValue
        k.btarget_order = lmb.target_order
    else:
        # integrable kernels only exist with a finite horizon, i.e. without surface terms (:918-921)
        k.bscaling = 0.
        k.bsingularity = kernel.singularityValue+1.
        k.btarget_order = lm.target_order
    # a finite horizon switches the surface terms off (:918-921)
    zero_exterior = 1 if (zeroExterior and not kernel.finiteHorizon) else 0
    # singular tables: the objects the reference's local matrices already hold (quadrature.pxd:25-26)
    keep = []

    def arrays(qr):
        a = (np.ascontiguousarray(qr.nodes, dtype=np.float64), np.ascontiguousarray(qr.weights, dtype=np.float64))
        keep.append(a)
        return a
    if dim == 2:
        nodes, weights = arrays(lm.qrId); _fill_rule(&r.identical, nodes, weights)
        nodes, weights = arrays(lm.qrEdge); _fill_rule(&r.edge, nodes, weights)
        nodes, weights = arrays(lm.qrVertex); _fill_rule(&r.vertex, nodes, weights)
        nodes, weights = arrays(lmb.qrEdge); _fill_rule(&r.bedge, nodes, weights)
        nodes, weights = arrays(lmb.qrVertex); _fill_rule(&r.bvertex, nodes, weights)
    else:
        nodes, weights = arrays(lm.qrId); _fill_rule(&r.identical, nodes, weights)
        nodes, weights = arrays(lm.qrVertex); _fill_rule(&r.vertex, nodes, weights)
        nodes, weights = arrays(lmb.qrVertex); _fill_rule(&r.bvertex, nodes, weights)
    A = Dense_LinearOperator(np.empty((N, N), dtype=np.float64))
    data = A.data
    try:
        for attempt in range(2):
            cl, fc = _regular_rules(dim, max_regular_order)
            keep.append((cl, fc))
            free(cell)
            free(facet)
            cell = <pnb_rule_t *>malloc((max_regular_order+1)*sizeof(pnb_rule_t))
            facet = <pnb_rule_t *>malloc((max_regular_order+1)*sizeof(pnb_rule_t))
            memset(cell, 0, (max_regular_order+1)*sizeof(pnb_rule_t))
            memset(facet, 0, (max_regular_order+1)*sizeof(pnb_rule_t))
            for o in range(1, max_regular_order+1):
                nodes, weights = cl[o]
                _fill_rule(&cell[o], nodes, weights)
                nodes, weights = fc[o]
                _fill_rule(&facet[o], nodes, weights)
            r.max_order = max_regular_order
            r.cell = cell
            r.facet = facet
            if prob == NULL:
                rc = pnb_problem_create(&m, &d, &k, &r, device, &prob)
            else:
                rc = pnb_problem_set_rules(prob, &r)
            if rc != 0:
                raise PNB200Error(_last_error())
            with nogil:
                if porder == 1 and smode == 0:
                    rc = pnb_dense_assemble(prob, zero_exterior, 0, N, &data[0, 0], N, 0)
                else:
                    rc = pnb_dense_assemble_element_smooth(prob, smode, sa, sbmode, sba, porder, dpe, N, &edofs[0, 0], zero_exterior,
                                                           &data[0, 0], N, 0)
            if rc == 0:
                break
            if rc != -5 or attempt == 1:       # PNB_ERR_ORDER: the reference grows its rule cache lazily (addQuadRule)
                raise PNB200Error(_last_error())
            if pnb_max_order(prob, zero_exterior, &need) != 0:
                raise PNB200Error(_last_error())
            max_regular_order = max(need, max_regular_order+1)
    finally:
        if prob != NULL:
            pnb_problem_destroy(prob)
        free(cell)
        free(facet)
    return A


_subclass = None


def builder_class():
    """PyNucleus_nl.nonlocalBuilder with getDense served by libpnb200 where the configuration is supported"""
    global _subclass
    if _subclass is None:
        from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder

        class nonlocalBuilderB200(nonlocalBuilder):
            def __init__(self, dm, kernel, params={}, zeroExterior=True, *args, **kwargs):
                super().__init__(dm, kernel, params, zeroExterior, *args, **kwargs)
                self._zeroExterior = kwargs.get('boundary', zeroExterior)

            def getDense(self, trySparsification=False):
                if trySparsification or not supported(self):
                    return super().getDense(trySparsification)
                return getDense(self, self._zeroExterior, self.params.get('device', 0), self.params.get('max_regular_order', 24))
        _subclass = nonlocalBuilderB200
    return _subclass


def install():
    """make `PyNucleus_nl.nonlocalBuilder` (and hence DoFMap.assembleNonlocal, fem/PyNucleus_fem/DoFMaps.pyx:877-899,
    and the drivers) resolve to the accelerated subclass"""
    import PyNucleus_nl
    import PyNucleus_nl.nonlocalAssembly as NA
    cls = builder_class()
    PyNucleus_nl.nonlocalBuilder = cls
    NA.nonlocalBuilder_py = cls
    return cls
