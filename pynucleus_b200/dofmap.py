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
        bv = np.nonzero(isb)[0]
        num[bv] = -1-np.arange(bv.shape[0])
        self.dofs = np.ascontiguousarray(num[mesh.cells], dtype=INDEX)
        self.num_dofs = int(interior.shape[0])
        self.num_boundary_dofs = int(bv.shape[0])
        self._vertex2dof = num

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
