"""P1 / P2 DoFMaps: the cell -> DoF table consumed by the assembly path
(fem/PyNucleus_fem/DoFMaps.pyx:61-330; numbering :157-322)."""
import numpy as np

from .mesh import INDEX


def _shape_values(polynomialOrder, manifold_dim, bary):
    """local shape functions at barycentric points, [dofs_per_element, n] (DoFMaps.pyx:1854-1880, 1932-2005)"""
    lam = np.asarray(bary)
    if polynomialOrder == 0:
        return np.ones((1, lam.shape[1]))
    if polynomialOrder == 1:
        return lam.copy()
    if polynomialOrder == 3:
        assert manifold_dim == 1
        return np.array([4.5*lam[0]*(lam[0]-1./3.)*(lam[0]-2./3.), 4.5*lam[1]*(lam[1]-1./3.)*(lam[1]-2./3.),
                         13.5*lam[0]*lam[1]*(lam[0]-1./3.), 13.5*lam[1]*lam[0]*(lam[1]-1./3.)])
    phi = [lam[k]*(2.*lam[k]-1.) for k in range(manifold_dim+1)]
    phi.append(4.*lam[0]*lam[1])
    if manifold_dim == 2:
        phi += [4.*lam[1]*lam[2], 4.*lam[0]*lam[2]]
    return np.array(phi)


def _assembleRHS(dm, fun, qr_order=None, rule=None):
    """b_i = int f phi_i (DoFMap.assembleRHS, DoFMaps.pyx:766-785): simplex rule of order 2 p + 2 per cell unless a rule
    (barycentric nodes [dim+1, n], weights) is given; host side"""
    from . import quadrature
    mesh = dm.mesh
    md = mesh.manifold_dim
    bary, w = quadrature.regular(2*dm.polynomialOrder+2 if qr_order is None else qr_order, md) if rule is None else rule
    phi = _shape_values(dm.polynomialOrder, md, bary)                          # [dpe, nq]
    x = np.einsum('kq,ckd->cqd', bary, mesh.vertices[mesh.cells])                # [nc, nq, dim]
    f = np.array([[fun(x[c, q]) for q in range(x.shape[1])] for c in range(x.shape[0])]) if callable(fun) else \
        np.full(x.shape[:2], float(fun))
    loc = mesh.volVector[:, None]*np.einsum('cq,q,kq->ck', f, w, phi)            # [nc, dpe]
    b = np.zeros(dm.num_dofs)
    m = dm.dofs >= 0
    np.add.at(b, dm.dofs[m], loc[m])
    return b


class P1_DoFMap:
    polynomialOrder = 1

    def __init__(self, mesh, tag=None):
        """Boundary vertices (all of them for the default tag) get negative
        DoFs, interior vertices are numbered by first appearance in cell
        order.  `tag` may be a boolean vertex mask marking boundary vertices."""
        self.mesh = mesh
        self.dim = mesh.dim
        self.dofs_per_vertex = 1
        self.dofs_per_edge = 0
        self.dofs_per_element = mesh.manifold_dim+1
        nv = mesh.num_vertices
        if tag is None:
            isb = np.zeros(nv, dtype=bool)
            isb[mesh.boundaryVertices] = True
        else:
            isb = np.asarray(tag, dtype=bool)
        flat = mesh.cells.ravel()
        uniq, first = np.unique(flat, return_index=True)
        order = uniq[np.argsort(first, kind='stable')]
        interior = order[~isb[order]]
        num = np.empty(nv, dtype=np.int64)
        num[interior] = np.arange(interior.shape[0])
        # boundary DoFs: -1, -2, ... in the order of mesh.boundaryVertices (DoFMaps.pyx:158-163)
        bv = np.asarray(mesh.boundaryVertices) if tag is None else np.nonzero(isb)[0]
        num[bv] = -1-np.arange(bv.shape[0])
        self.dofs = np.ascontiguousarray(num[mesh.cells], dtype=INDEX)
        self.num_dofs = int(interior.shape[0])
        self.num_boundary_dofs = int(bv.shape[0])
        self._vertex2dof = num

    @classmethod
    def fromArrays(cls, mesh, dofs, num_dofs, num_boundary_dofs=None):
        """DoFMap over `mesh` with a given cell -> DoF table (negative = boundary DoF -1-k), e.g. the `dofs` array of a
        PyNucleus DoFMap built with an indicator tag (fem/PyNucleus_fem/DoFMaps.pyx:358-)"""
        dm = cls.__new__(cls)
        dm.mesh = mesh
        dm.dim = mesh.dim
        dm.dofs_per_vertex, dm.dofs_per_edge, dm.dofs_per_element = 1, 0, mesh.manifold_dim+1
        dm.dofs = np.ascontiguousarray(dofs, dtype=INDEX)
        assert dm.dofs.shape == mesh.cells.shape
        dm.num_dofs = int(num_dofs)
        v2d = np.zeros(mesh.num_vertices, dtype=np.int64)
        v2d[mesh.cells.ravel()] = dm.dofs.ravel()
        assert np.array_equal(v2d[mesh.cells], dm.dofs), 'P1: one DoF per vertex'
        dm._vertex2dof = v2d
        dm.num_boundary_dofs = int((v2d < 0).sum()) if num_boundary_dofs is None else int(num_boundary_dofs)
        return dm

    def getComplementDoFMap(self):
        """DoFs and boundary DoFs swapped (fem/PyNucleus_fem/DoFMaps.pyx:1170-1184)"""
        from copy import copy
        bdm = copy(self)
        bdm.dofs = np.ascontiguousarray(-self.dofs-1, dtype=INDEX)
        bdm.num_dofs, bdm.num_boundary_dofs = self.num_boundary_dofs, self.num_dofs
        bdm._vertex2dof = -self._vertex2dof-1
        return bdm

    def combine(self, other):
        """all DoFs of two complementary maps, the second map's DoFs after the first's (DoFMaps.pyx:1563-1588)"""
        from copy import copy
        assert type(self) is type(other), "Cannot combine DoFMaps of different type"
        assert self.mesh is other.mesh, "Both DoFMaps need to have the same mesh"
        assert self.num_dofs == other.num_boundary_dofs and self.num_boundary_dofs == other.num_dofs, "DoFMaps need to be complementary"
        if np.any((self.dofs >= 0) == (other.dofs >= 0)):
            raise NotImplementedError()
        dmc = copy(self)
        dmc.dofs = np.ascontiguousarray(np.where(self.dofs >= 0, self.dofs, self.num_dofs+other.dofs), dtype=INDEX)
        dmc.num_dofs = self.num_dofs+other.num_dofs
        dmc.num_boundary_dofs = 0
        dmc._vertex2dof = np.where(self._vertex2dof >= 0, self._vertex2dof, self.num_dofs+other._vertex2dof)
        return dmc

    def __repr__(self):
        return 'P1 DoFMap with {} DoFs and {} boundary DoFs.'.format(self.num_dofs, self.num_boundary_dofs)

    def getDoFCoordinates(self):
        coords = np.zeros((self.num_dofs, self.dim))
        m = self._vertex2dof >= 0
        coords[self._vertex2dof[m]] = self.mesh.vertices[m]
        return coords

    def ones(self):
        return np.ones(self.num_dofs)

    def zeros(self):
        return np.zeros(self.num_dofs)

    def assembleRHS(self, fun, qr_order=None, rule=None):
        return _assembleRHS(self, fun, qr_order, rule)


class P0_DoFMap:
    """piecewise constants (DoFMaps.pyx:1788-1806): one dof per cell, numbered like the cells; no boundary dofs.  The reference
    accepts them for kernels with s < 1/2 (the space is not conforming otherwise, fractionalLaplacian2D.pyx:596-598)."""
    polynomialOrder = 0

    def __init__(self, mesh, tag=None):
        if tag is not None:
            raise NotImplementedError('P0_DoFMap: default tag only')
        self.mesh = mesh
        self.dim = mesh.dim
        self.dofs_per_vertex = self.dofs_per_edge = 0
        self.dofs_per_element = 1
        self.dofs = np.ascontiguousarray(np.arange(mesh.num_cells).reshape(-1, 1), dtype=INDEX)
        self.num_dofs = int(mesh.num_cells)
        self.num_boundary_dofs = 0

    def vertexPart(self):
        """a P1-shaped table for the device problem behind the element kernel (mesh, kernel, tables): the cell's dof in the
        first slot, nothing in the others -- the element kernel never reads it"""
        from copy import copy
        v = copy(self)
        nvc = self.mesh.manifold_dim+1
        t = np.full((self.mesh.num_cells, nvc), -1, dtype=INDEX)
        t[:, 0] = np.arange(self.mesh.num_cells)
        v.dofs = np.ascontiguousarray(t)
        v.dofs_per_element = nvc
        return v

    def __repr__(self):
        return 'P0 DoFMap with {} DoFs and {} boundary DoFs.'.format(self.num_dofs, self.num_boundary_dofs)

    def getDoFCoordinates(self):
        return self.mesh.vertices[self.mesh.cells].mean(axis=1)

    def assembleRHS(self, fun, qr_order=None, rule=None):
        return _assembleRHS(self, fun, qr_order, rule)

    def ones(self):
        return np.ones(self.num_dofs)

    def zeros(self):
        return np.zeros(self.num_dofs)


class P3_DoFMap:
    """continuous piecewise cubic elements on intervals (DoFMaps.pyx:2106-2122): one dof per vertex, two per cell (at 1/3 and
    2/3 of the cell); numbered cell by cell, the vertices of a cell before its own two dofs.  (Triangles: not built.)"""
    polynomialOrder = 3

    def __init__(self, mesh, tag=None):
        if tag is not None or mesh.manifold_dim != 1:
            raise NotImplementedError('P3_DoFMap: intervals, default tag')
        self.mesh = mesh
        self.dim = mesh.dim
        self.dofs_per_vertex, self.dofs_per_edge, self.dofs_per_element = 1, 0, 4
        cells = np.asarray(mesh.cells)
        nc = cells.shape[0]
        vdof = {}
        nb = -1
        for v in np.asarray(mesh.boundaryVertices).tolist():
            vdof[v] = nb
            nb -= 1
        dofs = np.empty((nc, 4), dtype=np.int64)
        n = 0
        for i in range(nc):
            for k in range(2):
                v = int(cells[i, k])
                if v not in vdof:
                    vdof[v] = n
                    n += 1
                dofs[i, k] = vdof[v]
            dofs[i, 2], dofs[i, 3] = n, n+1
            n += 2
        self.dofs = np.ascontiguousarray(dofs, dtype=INDEX)
        self.num_dofs = int(n)
        self.num_boundary_dofs = int(-nb-1)

    def vertexPart(self):
        from copy import copy
        v = copy(self)
        v.dofs = np.ascontiguousarray(self.dofs[:, :2], dtype=INDEX)
        v.dofs_per_element = 2
        return v

    def __repr__(self):
        return 'P3 DoFMap with {} DoFs and {} boundary DoFs.'.format(self.num_dofs, self.num_boundary_dofs)

    def assembleRHS(self, fun, qr_order=None, rule=None):
        return _assembleRHS(self, fun, qr_order, rule)

    def ones(self):
        return np.ones(self.num_dofs)

    def zeros(self):
        return np.zeros(self.num_dofs)


class P2_DoFMap:
    """continuous piecewise quadratic elements (DoFMaps.pyx:1978-2031): one dof per vertex and per edge (1D: per vertex
    and per cell).  Local order on a cell: vertices, then the edges (0,1), (1,2), (0,2).  Numbering as in the reference's
    constructor (:157-322): boundary vertices -1, -2, ... in the order of mesh.boundaryVertices, boundary edges continue
    the negative numbers in the order of mesh.boundaryEdges; the other dofs are numbered cell by cell by first appearance,
    the vertices of a cell before its edges."""
    polynomialOrder = 2

    def __init__(self, mesh, tag=None):
        if tag is not None:
            raise NotImplementedError('P2_DoFMap: default tag (whole boundary) only')
        self.mesh = mesh
        self.dim = mesh.dim
        self.dofs_per_vertex = 1
        nvc = mesh.manifold_dim+1
        cells = np.asarray(mesh.cells)
        nc = cells.shape[0]
        vdof = {}
        nb = -1
        for v in np.asarray(mesh.boundaryVertices).tolist():
            vdof[v] = nb
            nb -= 1
        if mesh.manifold_dim == 1:
            self.dofs_per_edge, self.dofs_per_element = 0, 3
            dofs = np.empty((nc, 3), dtype=np.int64)
            n = 0
            for i in range(nc):
                for k in range(2):
                    v = int(cells[i, k])
                    if v not in vdof:
                        vdof[v] = n
                        n += 1
                    dofs[i, k] = vdof[v]
                dofs[i, 2] = n
                n += 1
        else:
            self.dofs_per_edge, self.dofs_per_element = 1, 6
            edof = {}
            for e in np.asarray(mesh.boundaryFacets).reshape(-1, 2).tolist():
                edof[(min(e), max(e))] = nb
                nb -= 1
            dofs = np.empty((nc, 6), dtype=np.int64)
            n = 0
            for i in range(nc):
                c = [int(x) for x in cells[i]]
                for k in range(3):
                    if c[k] not in vdof:
                        vdof[c[k]] = n
                        n += 1
                    dofs[i, k] = vdof[c[k]]
                for k, (a, b) in enumerate(((0, 1), (1, 2), (0, 2))):
                    e = (min(c[a], c[b]), max(c[a], c[b]))
                    if e not in edof:
                        edof[e] = n
                        n += 1
                    dofs[i, 3+k] = edof[e]
        self.dofs = np.ascontiguousarray(dofs, dtype=INDEX)
        self.num_dofs = int(n)
        self.num_boundary_dofs = int(-nb-1)
        self._nvc = nvc

    def vertexPart(self):
        """the vertex dofs as a P1-shaped table over the same numbering (the device problem keeps the mesh, the kernel and
        the tables behind it; the element's own table goes to pnb_dense_assemble_element)"""
        from copy import copy
        v = copy(self)
        v.dofs = np.ascontiguousarray(self.dofs[:, :self._nvc], dtype=INDEX)
        v.dofs_per_element = self._nvc
        return v

    def __repr__(self):
        return 'P2 DoFMap with {} DoFs and {} boundary DoFs.'.format(self.num_dofs, self.num_boundary_dofs)

    def getDoFCoordinates(self):
        nodes = (np.array([[1., 0.], [0., 1.], [0.5, 0.5]]) if self.mesh.manifold_dim == 1 else
                 np.array([[1., 0., 0.], [0., 1., 0.], [0., 0., 1.], [0.5, 0.5, 0.], [0., 0.5, 0.5], [0.5, 0., 0.5]]))
        coords = np.zeros((self.num_dofs, self.dim))
        x = np.einsum('kv,cvd->ckd', nodes, self.mesh.vertices[self.mesh.cells])
        m = self.dofs >= 0
        coords[self.dofs[m]] = x[m]
        return coords

    def ones(self):
        return np.ones(self.num_dofs)

    def zeros(self):
        return np.zeros(self.num_dofs)

    def assembleRHS(self, fun, qr_order=None, rule=None):
        return _assembleRHS(self, fun, qr_order, rule)
