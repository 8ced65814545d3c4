"""ORACLE (test infrastructure; never imported by the product path).

Plain-loop restatement of the reference's mesh generation and P1 numbering
for the synthetic inputs of the assembly path:

* simpleInterval                 fem/PyNucleus_fem/mesh.py:209-217
* uniform_disc                   fem/PyNucleus_fem/mesh.py:946-960
* uniform refinement (1D/2D)     fem/PyNucleus_fem/meshCy.pyx:863-, 1052-1109
* radial transformer             fem/PyNucleus_fem/meshCy.pyx:75-89
* boundary edges                 fem/PyNucleus_fem/meshCy.pyx:1811-1848
* h / volume per cell            fem/PyNucleus_fem/meshCy.pyx:1654-1740
* P1 DoF numbering               fem/PyNucleus_fem/DoFMaps.pyx:157-210

Pinned against meshes dumped from the reference (tests/golden/disc_mesh_r*.npz,
interval_*.npz).
"""
from math import sqrt

import numpy as np


class Mesh:
    def __init__(self, vertices, cells, radial=False):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.radial = radial
        self.dim = self.vertices.shape[1]

    @property
    def num_cells(self):
        return self.cells.shape[0]

    @property
    def num_vertices(self):
        return self.vertices.shape[0]

    # -- geometry ---------------------------------------------------------
    def _edge_lengths(self):
        """(hVector, h, hmin) as hdeltaCy rounds them (meshCy.pyx:1654-1732): orc_edge_lengths in nonlocal_oracle.c"""
        import ctypes
        from . import lib
        if getattr(self, '_hcache', None) is None or self._hcache[0] is not self.vertices:
            v = np.ascontiguousarray(self.vertices, dtype=np.float64)
            c = np.ascontiguousarray(self.cells, dtype=np.int32)
            h = np.empty(c.shape[0])
            hmax, hmin = ctypes.c_double(0.), ctypes.c_double(0.)
            lib().orc_edge_lengths(ctypes.c_int(self.dim), ctypes.c_int(c.shape[0]), ctypes.c_void_p(v.ctypes.data),
                                   ctypes.c_void_p(c.ctypes.data), ctypes.c_void_p(h.ctypes.data), ctypes.byref(hmax),
                                   ctypes.byref(hmin))
            self._hcache = (self.vertices, h, float(hmax.value), float(hmin.value))
        return self._hcache[1:]

    @property
    def hVector(self):
        return self._edge_lengths()[0]

    @property
    def volVector(self):
        v, c = self.vertices, self.cells
        if self.dim == 1:
            return np.abs(v[c[:, 1], 0]-v[c[:, 0], 0])
        a = v[c[:, 1]]-v[c[:, 0]]
        b = v[c[:, 2]]-v[c[:, 0]]
        return np.abs(a[:, 0]*b[:, 1]-a[:, 1]*b[:, 0])*0.5

    @property
    def h(self):
        return self._edge_lengths()[1]

    @property
    def hmin(self):
        """shortest edge of the mesh (meshCy.pyx:1724)"""
        return self._edge_lengths()[2]

    @property
    def diam(self):
        return float(np.linalg.norm(self.vertices.max(axis=0)-self.vertices.min(axis=0), 2))

    # -- boundary ---------------------------------------------------------
    def boundary_facets(self):
        """2D: edges that belong to exactly one cell, oriented as in that cell
        (meshCy.pyx:1826-1848).  1D: vertices that belong to exactly one cell."""
        c = self.cells
        if self.dim == 1:
            cnt = {}
            for i in range(c.shape[0]):
                for k in range(2):
                    cnt[c[i, k]] = cnt.get(c[i, k], 0)+1
            return np.array([[v] for v, n in cnt.items() if n == 1], dtype=np.int32)
        seen = {}
        for i in range(c.shape[0]):
            for k in range(3):
                a, b = int(c[i, k]), int(c[i, (k+1) % 3])
                key = (min(a, b), max(a, b))
                if key in seen:
                    del seen[key]
                else:
                    seen[key] = (a, b)
        return np.array(list(seen.values()), dtype=np.int32)

    # -- refinement -------------------------------------------------------
    def refine(self):
        v, c = self.vertices, self.cells
        if self.dim == 1:
            nc = c.shape[0]
            nv = v.shape[0]
            newv = np.zeros((nv+nc, 1))
            newv[:nv] = v
            newc = np.zeros((2*nc, 2), dtype=np.int32)
            for i in range(nc):
                c0, c1 = c[i]
                newv[nv+i, 0] = (v[c0, 0]+v[c1, 0])*0.5
                newc[2*i] = (c0, nv+i)
                newc[2*i+1] = (nv+i, c1)
            return Mesh(newv, newc)
        nc = c.shape[0]
        nv = v.shape[0]
        mid = {}
        order = []
        for i in range(nc):
            c0, c1, c2 = (int(t) for t in c[i])
            for a, b in ((c0, c1), (c0, c2), (c1, c2)):
                key = (min(a, b), max(a, b))
                if key not in mid:
                    mid[key] = nv+len(order)
                    order.append(key)
        newv = np.zeros((nv+len(order), 2))
        newv[:nv] = v
        for key, n in mid.items():
            newv[n] = (v[key[0]]+v[key[1]])*0.5
        newc = np.zeros((4*nc, 3), dtype=np.int32)
        for i in range(nc):
            c0, c1, c2 = (int(t) for t in c[i])
            m01 = mid[(min(c0, c1), max(c0, c1))]
            m02 = mid[(min(c0, c2), max(c0, c2))]
            m12 = mid[(min(c1, c2), max(c1, c2))]
            newc[4*i] = (c0, m01, m02)
            newc[4*i+1] = (c1, m12, m01)
            newc[4*i+2] = (c2, m02, m12)
            newc[4*i+3] = (m01, m12, m02)
        if self.radial:
            for key, n in mid.items():
                r1 = sqrt(v[key[0], 0]**2+v[key[0], 1]**2)
                r2 = sqrt(v[key[1], 0]**2+v[key[1], 1]**2)
                r = 0.5*r1+0.5*r2
                r3 = sqrt(newv[n, 0]**2+newv[n, 1]**2)
                newv[n, 0] *= r/r3
                newv[n, 1] *= r/r3
        return Mesh(newv, newc, self.radial)


def interval(a=-1., b=1., noRef=0):
    m = Mesh(np.array([[a], [b]]), np.array([[0, 1]]))
    for _ in range(noRef):
        m = m.refine()
    return m


def disc(noRef=0, radius=1.):
    pts = [(0., 0.)]
    n = 6
    for i in range(n):
        pts.append((radius*np.cos(i*2*np.pi/n), radius*np.sin(i*2*np.pi/n)))
    cells = [(0, i, i+1) for i in range(1, n)]+[(0, n, 1)]
    m = Mesh(np.array(pts), np.array(cells), radial=True)
    for _ in range(noRef):
        m = m.refine()
    return m


def p1_dofs(mesh):
    """DoFMaps.pyx:157-210 for P1 with the default tag: boundary vertices get
    -1,-2,... in the order of the boundary vertex list, interior vertices are
    numbered by first appearance in cell order.  Only the SIGN of boundary
    entries matters to the assembly (nonlocalAssembly_{SCALAR}.pxi:138-150)."""
    bf = mesh.boundary_facets()
    bverts = []
    seen = set()
    for f in bf:
        for v in f:
            if int(v) not in seen:
                seen.add(int(v))
                bverts.append(int(v))
    num = {}
    for k, v in enumerate(bverts):
        num[v] = -1-k
    dofs = np.zeros(mesh.cells.shape, dtype=np.int32)
    n = 0
    for i in range(mesh.num_cells):
        for k in range(mesh.cells.shape[1]):
            v = int(mesh.cells[i, k])
            if v not in num:
                num[v] = n
                n += 1
            dofs[i, k] = num[v]
    return dofs, n
