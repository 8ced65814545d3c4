"""Simplicial meshes: the INPUT side of the assembly path.

Only what the path needs from fem/PyNucleus_fem/mesh.py and meshCy.pyx:
vertices/cells, volVector, hVector, h, hmin, diam, boundary facets, uniform
refinement with the reference's vertex/cell numbering (meshCy.pyx:506-583,
863-, 1052-1109), the radial transformer of uniform_disc (meshCy.pyx:75-89) and
the factories simpleInterval / uniform_disc (mesh.py:209-217, 946-960).
Everything is vectorised so that the 10^5..10^6 cell meshes of the bench are
generated in seconds.
"""
import numpy as np

INDEX = np.int32
REAL = np.float64


class meshNd:
    def __init__(self, vertices, cells, boundary=None, radial=False, boundaryVertices=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=REAL)
        self.cells = np.ascontiguousarray(cells, dtype=INDEX)
        self.dim = self.vertices.shape[1]
        self.manifold_dim = self.cells.shape[1]-1
        self.radial = radial
        self._boundary = None if boundary is None else np.ascontiguousarray(boundary, dtype=INDEX)
        # order of the boundary vertices = order of the negative DoF numbers (DoFMaps.pyx:158-163); the reference
        # appends the midpoints of the boundary edges in edge order when it refines (meshCy.pyx:534-555)
        self._bvertices = None if boundaryVertices is None else np.ascontiguousarray(boundaryVertices, dtype=INDEX)
        self._h = self._vol = None

    num_vertices = property(lambda self: self.vertices.shape[0])
    num_cells = property(lambda self: self.cells.shape[0])

    # ---- geometry -------------------------------------------------------
    def _edge_lengths(self):
        """hVector, h, hmin exactly as hdeltaCy rounds them (fem/PyNucleus_fem/meshCy.pyx:1654-1732): the host routine
        pnb_mesh_edge_lengths of the library (edge lengths through a fused-multiply-add dot product like the BLAS ddot
        behind the reference's mydot; getQuadOrder is sensitive to the last bit of h)"""
        import ctypes
        from . import _lib
        h = np.empty(self.num_cells)
        hmax, hmin = ctypes.c_double(0.), ctypes.c_double(0.)
        _lib.check(_lib.lib().pnb_mesh_edge_lengths(self.manifold_dim, self.num_cells, self.vertices.ctypes.data,
                                                    self.cells.ctypes.data, h.ctypes.data, ctypes.byref(hmax),
                                                    ctypes.byref(hmin)))
        self._h, self._hmax, self._hmin = h, float(hmax.value), float(hmin.value)

    @property
    def hVector(self):
        """longest edge per cell"""
        if self._h is None:
            self._edge_lengths()
        return self._h

    @property
    def volVector(self):
        if self._vol is None:
            v, c = self.vertices, self.cells
            if self.manifold_dim == 1:
                self._vol = np.abs(v[c[:, 1], 0]-v[c[:, 0], 0])
            else:
                a = v[c[:, 1]]-v[c[:, 0]]
                b = v[c[:, 2]]-v[c[:, 0]]
                self._vol = np.abs(a[:, 0]*b[:, 1]-a[:, 1]*b[:, 0])*0.5
        return self._vol

    @property
    def h(self):
        """longest edge of the mesh"""
        if self._h is None:
            self._edge_lengths()
        return self._hmax

    @property
    def hmin(self):
        """SHORTEST edge of the mesh (meshCy.pyx:1724: min over all edges, not over the cells' longest edges)"""
        if self._h is None:
            self._edge_lengths()
        return self._hmin

    volume = property(lambda self: float(self.volVector.sum()))

    @property
    def diam(self):
        return float(np.linalg.norm(self.vertices.max(axis=0)-self.vertices.min(axis=0), 2))

    # ---- boundary -------------------------------------------------------
    @property
    def boundaryFacets(self):
        """2D: boundary edges oriented as in their cell (meshCy.pyx:1826-1848);
        1D: boundary vertices as an (nb, 1) array."""
        if self._boundary is None:
            c = self.cells
            if self.manifold_dim == 1:
                cnt = np.bincount(c.ravel(), minlength=self.num_vertices)
                self._boundary = np.nonzero(cnt == 1)[0].astype(INDEX).reshape(-1, 1)
            else:
                e = np.concatenate((c[:, [0, 1]], c[:, [1, 2]], c[:, [2, 0]]))
                key = np.minimum(e[:, 0], e[:, 1]).astype(np.int64)*self.num_vertices+np.maximum(e[:, 0], e[:, 1])
                _, idx, cnt = np.unique(key, return_index=True, return_counts=True)
                self._boundary = np.ascontiguousarray(e[np.sort(idx[cnt == 1])], dtype=INDEX)
        return self._boundary

    boundaryEdges = property(lambda self: self.boundaryFacets)

    @property
    def boundaryVertices(self):
        if self._bvertices is not None:
            return self._bvertices
        return np.unique(self.boundaryFacets.ravel()).astype(INDEX)

    def get_surface_mesh(self):
        return surfaceMesh(self.vertices, self.boundaryFacets)

    # ---- refinement -----------------------------------------------------
    def refine(self):
        v, c = self.vertices, self.cells
        nv, nc = v.shape[0], c.shape[0]
        if self.manifold_dim == 1:
            mid = (v[c[:, 0]]+v[c[:, 1]])*0.5
            newv = np.vstack((v, mid))
            m = nv+np.arange(nc, dtype=INDEX)
            newc = np.empty((2*nc, 2), dtype=INDEX)
            newc[0::2, 0], newc[0::2, 1] = c[:, 0], m
            newc[1::2, 0], newc[1::2, 1] = m, c[:, 1]
            return meshNd(newv, newc, self._boundary, self.radial, self._bvertices)
        # edges in sweep order (c0c1), (c0c2), (c1c2) per cell; numbered at first appearance
        e = np.empty((nc, 3, 2), dtype=np.int64)
        e[:, 0], e[:, 1], e[:, 2] = c[:, [0, 1]], c[:, [0, 2]], c[:, [1, 2]]
        e = e.reshape(-1, 2)
        lo, hi = e.min(axis=1), e.max(axis=1)
        key = lo*nv+hi
        uniq, first, inv = np.unique(key, return_index=True, return_inverse=True)
        rank = np.empty(uniq.shape[0], dtype=np.int64)
        rank[np.argsort(first, kind='stable')] = np.arange(uniq.shape[0])
        mids = (nv+rank[inv]).reshape(nc, 3)          # midpoints of (c0c1), (c0c2), (c1c2)
        elo, ehi = uniq//nv, uniq % nv
        newv = np.empty((nv+uniq.shape[0], 2))
        newv[:nv] = v
        newv[nv+rank] = (v[elo]+v[ehi])*0.5
        m01, m02, m12 = mids[:, 0], mids[:, 1], mids[:, 2]
        newc = np.empty((4*nc, 3), dtype=INDEX)
        newc[0::4] = np.stack((c[:, 0], m01, m02), axis=1)
        newc[1::4] = np.stack((c[:, 1], m12, m01), axis=1)
        newc[2::4] = np.stack((c[:, 2], m02, m12), axis=1)
        newc[3::4] = np.stack((m01, m12, m02), axis=1)
        if self.radial:
            # radialMeshTransformer, radius=0 branch: new midpoints move to the mean radius of the edge ends
            r1 = np.sqrt(v[elo, 0]**2+v[elo, 1]**2)
            r2 = np.sqrt(v[ehi, 0]**2+v[ehi, 1]**2)
            r = 0.5*r1+0.5*r2
            n = newv[nv+rank]
            r3 = np.sqrt(n[:, 0]**2+n[:, 1]**2)
            newv[nv+rank] = n*(r/r3)[:, None]
        # boundary edges are split in place (meshCy.pyx:534-555)
        be = self.boundaryFacets.astype(np.int64)
        bkey = be.min(axis=1)*nv+be.max(axis=1)
        bm = (nv+rank[np.searchsorted(uniq, bkey)]).astype(INDEX)
        newb = np.empty((2*be.shape[0], 2), dtype=INDEX)
        newb[0::2, 0], newb[0::2, 1] = be[:, 0], bm
        newb[1::2, 0], newb[1::2, 1] = bm, be[:, 1]
        return meshNd(newv, newc, newb, self.radial, np.concatenate((self.boundaryVertices, bm)))


class surfaceMesh:
    """mesh.get_surface_mesh(): shares the vertex array (mesh.py:2055-2068)"""

    def __init__(self, vertices, cells):
        self.vertices = vertices
        self.cells = cells
        self.num_cells = cells.shape[0]


def simpleInterval(a=0., b=1., numCells=1):
    vertices = (a+(b-a)*(np.arange(numCells+1)/numCells)).reshape(-1, 1)
    vertices[-1, 0] = b
    cells = np.stack((np.arange(numCells), np.arange(1, numCells+1)), axis=1)
    return meshNd(vertices, cells)


def uniform_disc(radius=1.):
    n = 6
    ang = np.arange(n)*2*np.pi/n
    pts = np.vstack(([[0., 0.]], np.stack((radius*np.cos(ang), radius*np.sin(ang)), axis=1)))
    cells = [(0, i, i+1) for i in range(1, n)]+[(0, n, 1)]
    return meshNd(pts, np.array(cells), radial=True)


def polygon_disc(n=6, radius=1.):
    """fan of n triangles, refined radially like uniform_disc (used to hit DoF
    counts between the hexagon's 4^r steps)"""
    ang = np.arange(n)*2*np.pi/n
    pts = np.vstack(([[0., 0.]], np.stack((radius*np.cos(ang), radius*np.sin(ang)), axis=1)))
    cells = [(0, i, i+1) for i in range(1, n)]+[(0, n, 1)]
    return meshNd(pts, np.array(cells), radial=True)


def refined(mesh, noRef):
    for _ in range(noRef):
        mesh = mesh.refine()
    return mesh
