"""P1 DoFMap: the cell -> DoF table consumed by the assembly path
(fem/PyNucleus_fem/DoFMaps.pyx:61-330; numbering :157-210)."""
import numpy as np

from .mesh import INDEX


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
