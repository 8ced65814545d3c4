"""ORACLE -- test infrastructure only.

CPU restatement of the reference's nonlocal assembly path (see the headers of
nonlocal_oracle.c, tables.py, meshes.py for the reference file:line of every
piece).  Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs
may import this package; the product (pynucleus_b200/) never does.

The restatement is pinned against the reference itself (stub-built copy,
oracle/refbuild) through tests/golden/*.npz.
"""
import ctypes
import os
import subprocess

import numpy as np

from . import tables, meshes  # noqa: F401

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = None


class _Rule(ctypes.Structure):
    _fields_ = [('n', ctypes.c_int), ('bary', ctypes.c_void_p), ('w', ctypes.c_void_p)]


class _Problem(ctypes.Structure):
    _fields_ = [('dim', ctypes.c_int), ('nv', ctypes.c_int), ('nc', ctypes.c_int),
                ('vertices', ctypes.c_void_p), ('cells', ctypes.c_void_p),
                ('vol', ctypes.c_void_p), ('h', ctypes.c_void_p), ('dofs', ctypes.c_void_p),
                ('num_dofs', ctypes.c_int), ('nb', ctypes.c_int), ('bfacets', ctypes.c_void_p),
                ('H0', ctypes.c_double),
                ('s', ctypes.c_double), ('C', ctypes.c_double), ('Cb', ctypes.c_double),
                ('singularity', ctypes.c_double), ('bsingularity', ctypes.c_double),
                ('target_order', ctypes.c_double), ('btarget_order', ctypes.c_double),
                ('qr_face', _Rule), ('qr_edge', _Rule), ('qr_vertex', _Rule),
                ('bqr_edge', _Rule), ('bqr_vertex', _Rule),
                ('max_order', ctypes.c_int),
                ('reg_cell', ctypes.c_void_p), ('reg_facet', ctypes.c_void_p), ('order_num_dofs', ctypes.c_int),
                ('labels', ctypes.c_void_p), ('blabels', ctypes.c_void_p), ('active_class', ctypes.c_int),
                ('pair_class', ctypes.c_ubyte*16), ('bpair_class', ctypes.c_ubyte*16), ('pair_orientation', ctypes.c_int),
                ('tempered', ctypes.c_double), ('smode', ctypes.c_int), ('bmode', ctypes.c_int),
                ('sa', ctypes.c_double), ('ba', ctypes.c_double)]


def build():
    """compile the C restatement (make -C oracle)"""
    subprocess.check_call(['make', '-s', '-C', _HERE], stdout=subprocess.DEVNULL)


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, 'liboracle.so')
        if not os.path.exists(path):
            build()
        _LIB = ctypes.CDLL(path)
        _LIB.orc_dense.restype = ctypes.c_int64
        _LIB.orc_max_order.restype = ctypes.c_int
    return _LIB


def set_threads(n):
    os.environ['OMP_NUM_THREADS'] = str(n)


class Problem:
    """One (mesh, P1 DoFMap, constant-order fractional kernel) assembly problem."""

    def __init__(self, vertices, cells, dofs, num_dofs, s, bfacets=None, target_order=None,
                 hVector=None, volVector=None, hmin=None, diam=None, max_order=None, order_num_dofs=None,
                 s_max=None, labels=None, blabels=None, pair_class=None, active_class=0, bpair_class=None,
                 pair_orientation=0, s_min=None, tempered=0., smooth=None):
        self.vertices = np.ascontiguousarray(vertices, dtype=np.float64)
        self.cells = np.ascontiguousarray(cells, dtype=np.int32)
        self.dofs = np.ascontiguousarray(dofs, dtype=np.int32)
        self.dim = self.vertices.shape[1]
        mesh = meshes.Mesh(self.vertices, self.cells)
        self.mesh = mesh
        self.h = np.ascontiguousarray(mesh.hVector if hVector is None else hVector)
        self.vol = np.ascontiguousarray(mesh.volVector if volVector is None else volVector)
        self.bfacets = np.ascontiguousarray(mesh.boundary_facets() if bfacets is None else bfacets, dtype=np.int32)
        self.num_dofs = int(num_dofs)
        self.s = float(s)
        hmin = float(mesh.hmin if hmin is None else hmin)
        diam = float(mesh.diam if diam is None else diam)
        self.H0 = diam/np.sqrt(8.)
        dim = self.dim
        self.singularity = -dim-2*s
        self.bsingularity = 1.-dim-2*s
        self.C = tables.fractional_scaling(dim, s, tempered=tempered)
        self.Cb = self.C*(1./s)      # phi = 1/s, kernels.py:151-160, kernelsCy.pyx:1990-1995
        if smooth is not None:
            # Gaussian / exponential kernels on the full space: smooth = (C, smode, sa, Cb, bmode, ba); singularity 0 for
            # the kernel and its boundary form (Kernel.__init__, kernelsCy.pyx:657-664); `s` is not used
            self.C, self.Cb = float(smooth[0]), float(smooth[3])
            self.singularity = self.bsingularity = 0.
        # two DoFMaps: the local matrices (orders, getQuadOrder) keep the DoF count of the first map
        self.order_num_dofs = int(order_num_dofs) if order_num_dofs else self.num_dofs
        # variable kernels: the singular quadrature orders follow the largest order s.max (fractionalLaplacian2D.pyx:606)
        sm = self.s if s_max is None else float(s_max)
        smn = sm if s_min is None else float(s_min)
        if smooth is not None:
            self.orders = tables.diag_orders(dim, 0., 0., hmin, self.H0, self.order_num_dofs, target_order)
        else:
            self.orders = tables.diag_orders(dim, -dim-2*sm, 1.-dim-2*sm, hmin, self.H0,
                                             self.order_num_dofs, target_order, min_singularity=-dim-2*smn,
                                             min_boundary_singularity=1.-dim-2*smn)
        self.near = tables.near_rules(dim, self.singularity, self.bsingularity, self.orders)
        self._keep = []
        P = _Problem()
        P.dim = dim
        P.nv, P.nc = self.vertices.shape[0], self.cells.shape[0]
        P.vertices = self.vertices.ctypes.data
        P.cells = self.cells.ctypes.data
        P.vol = self.vol.ctypes.data
        P.h = self.h.ctypes.data
        P.dofs = self.dofs.ctypes.data
        P.num_dofs = self.num_dofs
        P.order_num_dofs = self.order_num_dofs
        if labels is not None:
            self._labels = np.ascontiguousarray(labels, dtype=np.uint8)
            self._blabels = np.ascontiguousarray(blabels, dtype=np.uint8)
            P.labels = self._labels.ctypes.data
            P.blabels = self._blabels.ctypes.data
            P.active_class = int(active_class)
            for i, v in enumerate(np.asarray(pair_class, dtype=np.uint8).ravel()):
                P.pair_class[i] = int(v)
            for i, v in enumerate(np.asarray(pair_class if bpair_class is None else bpair_class, dtype=np.uint8).ravel()):
                P.bpair_class[i] = int(v)
            P.pair_orientation = int(pair_orientation)
        P.tempered = float(tempered)
        if smooth is not None:
            P.smode, P.sa, P.bmode, P.ba = int(smooth[1]), float(smooth[2]), int(smooth[4]), float(smooth[5])
        P.nb = self.bfacets.shape[0]
        P.bfacets = self.bfacets.ctypes.data
        P.H0 = self.H0
        P.s, P.C, P.Cb = self.s, self.C, self.Cb
        P.singularity, P.bsingularity = self.singularity, self.bsingularity
        P.target_order = self.orders['target_order']
        P.btarget_order = self.orders['b_target_order']
        if dim == 2:
            P.qr_face = self._rule(self.near[('interior', -3)])
            P.qr_edge = self._rule(self.near[('interior', -2)])
            P.qr_vertex = self._rule(self.near[('interior', -1)])
            P.bqr_edge = self._rule(self.near[('boundary', -2)])
            P.bqr_vertex = self._rule(self.near[('boundary', -1)])
        else:
            P.qr_face = self._rule(self.near[('interior', -2)])
            P.qr_vertex = self._rule(self.near[('interior', -1)])
            P.bqr_vertex = self._rule(self.near[('boundary', -1)])
        self.P = P
        self._set_regular(2)
        if max_order is None:
            max_order = lib().orc_max_order(ctypes.byref(self.P), 0, P.nc, 1)
        self._set_regular(max_order)

    def _rule(self, bw):
        b = np.ascontiguousarray(bw[0])
        w = np.ascontiguousarray(bw[1])
        self._keep += [b, w]
        return _Rule(b.shape[1], b.ctypes.data, w.ctypes.data)

    def _set_regular(self, max_order):
        cell = (_Rule*(max_order+1))()
        facet = (_Rule*(max_order+1))()
        for p in range(1, max_order+1):
            cell[p] = self._rule(tables.regular_rule(p, self.dim))
            facet[p] = self._rule(tables.regular_rule(p, self.dim-1))
        self._keep += [cell, facet]
        self.P.max_order = max_order
        self.P.reg_cell = ctypes.cast(cell, ctypes.c_void_p)
        self.P.reg_facet = ctypes.cast(facet, ctypes.c_void_p)

    # ------------------------------------------------------------------
    def pairs(self, pairs, with_contrib=True):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32)
        n = pairs.shape[0]
        nvc = self.dim+1
        panels = np.zeros(n, dtype=np.int32)
        perm1 = np.zeros((n, nvc), dtype=np.int32)
        perm2 = np.zeros((n, nvc), dtype=np.int32)
        nloc = (2*nvc)*(2*nvc+1)//2
        contribs = np.zeros((n, nloc)) if with_contrib else None
        lib().orc_pairs(ctypes.byref(self.P), n, pairs.ctypes.data_as(ctypes.c_void_p),
                        panels.ctypes.data_as(ctypes.c_void_p), perm1.ctypes.data_as(ctypes.c_void_p),
                        perm2.ctypes.data_as(ctypes.c_void_p),
                        contribs.ctypes.data_as(ctypes.c_void_p) if with_contrib else None)
        return panels, perm1, perm2, contribs

    def boundary_pairs(self, pairs, with_contrib=True):
        pairs = np.ascontiguousarray(pairs, dtype=np.int32)
        n = pairs.shape[0]
        nvc = self.dim+1
        panels = np.zeros(n, dtype=np.int32)
        contribs = np.zeros((n, nvc*(nvc+1)//2)) if with_contrib else None
        lib().orc_boundary_pairs(ctypes.byref(self.P), n, pairs.ctypes.data_as(ctypes.c_void_p),
                                 panels.ctypes.data_as(ctypes.c_void_p),
                                 contribs.ctypes.data_as(ctypes.c_void_p) if with_contrib else None)
        return panels, contribs

    def histogram(self, start=0, end=None):
        end = self.P.nc if end is None else end
        hist = np.zeros(4+256, dtype=np.int64)
        lib().orc_histogram(ctypes.byref(self.P), start, end, hist.ctypes.data_as(ctypes.c_void_p))
        return {k-3: int(v) for k, v in enumerate(hist) if v}

    def dense(self, zero_exterior=True, start=0, end=None, wrap=0, atomic=False):
        """getDense of the cell slice [start,end) (one MPI rank's share)."""
        end = self.P.nc if end is None else end
        N = self.num_dofs
        A = np.zeros(wrap, dtype=np.float64) if wrap else np.zeros((N, N))
        npairs = lib().orc_dense(ctypes.byref(self.P), start, end, int(zero_exterior),
                                 A.ctypes.data_as(ctypes.c_void_p), ctypes.c_size_t(wrap), int(atomic))
        self.last_npairs = npairs
        return A


def disc_problem(noRef, s=0.75, target_order=0.5):
    m = meshes.disc(noRef)
    dofs, n = meshes.p1_dofs(m)
    return Problem(m.vertices, m.cells, dofs, n, s, target_order=target_order)


def interval_problem(noRef, s=0.25):
    m = meshes.interval(-1., 1., noRef)
    dofs, n = meshes.p1_dofs(m)
    return Problem(m.vertices, m.cells, dofs, n, s)
