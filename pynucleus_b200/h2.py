"""H2 operator: cluster basis (leaf moments, transfer operators), far-field blocks and the three-pass matvec.

Restates, on top of `cluster_tree`:
  * tree_node.enterLeafValues        clusterMethodCy.pyx:1205-1325   V[dof, alpha] = int phi_dof(x) L_alpha(x) dx
  * transferMatrixBuilder.build      clusterMethodCy.pyx:2010-2072   T[alpha_parent, alpha_child] = L^P_alpha(xi^C_beta)
  * upwardPass / downwardPass        clusterMethodCy.pyx:1093-1176
  * H2Matrix.matvec                  clusterMethodCy.pyx:2269-2295   y = Anear x + sum_far V1 K12 V2^T x
  * far-field blocks                 nonlocalBuilder.getFarFieldBlocks (CUDA kernel, pnb_farfield_blocks)

The cluster bases are small dense blocks (m^d <= a few hundred columns); they are built on the host with numpy and
kept as float64 CUDA tensors, the passes run as sparse matrix products on the device.  The near field
(assembleClusters: cluster-union quadrature with surface terms, nonlocalAssembly_{SCALAR}.pxi:1663-2160) is assembled
per near cluster pair by the dense device path on the cluster-union sub-mesh (`assemble_clusters`); `nearFromDense`
(the near pattern filled from the dense operator) is kept for validation only.
"""
import numpy as np

from . import quadrature


def chebyshev(box, m):
    """tensor-factor Chebyshev nodes of a box, [m, dim] (clusterMethodCy.pyx:1259-1262)"""
    eta = np.cos((2.0*np.arange(m, 0, -1)-1.0)/(2.0*m)*np.pi)
    return (box[:, 1]-box[:, 0])[None, :]*0.5*(eta[:, None]+1.0)+box[None, :, 0]


def _multi_index(m, dim):
    """productIterator order: last dimension fastest"""
    if dim == 1:
        return np.arange(m).reshape(-1, 1)
    i, j = np.meshgrid(np.arange(m), np.arange(m), indexing='ij')
    return np.stack((i.ravel(), j.ravel()), axis=1)


def leaf_values(node, mesh, dm, d2c=None):
    """V[local dof, alpha] of a leaf cluster (all cells of the cluster at once)"""
    dim = mesh.dim
    m = node.interpolation_order
    dofs = node.dofs
    bary, w = quadrature.regular(m+2, dim)          # P1: quadOrder = order+2 (Sauter/Schwab p. 428)
    xi = chebyshev(node.box, m)                     # [m, dim]
    diff = xi[:, None, :]-xi[None, :, :]
    diff[np.arange(m), np.arange(m), :] = 1.
    beta = diff.prod(axis=1)                        # [m, dim]
    cells = cells_of_dofs(dm, dofs, d2c)            # cells around the DoFs of the cluster, ascending
    idx = _multi_index(m, dim)
    x = np.einsum('kq,ckd->cqd', bary, mesh.vertices[mesh.cells[cells]])      # [nc, nq, dim]
    d = x[:, :, None, :]-xi[None, None, :, :]                                   # [nc, nq, m, dim]
    # omega[.., l, :] = prod_{l' != l} d[.., l', :] from prefix and suffix products (O(m) instead of O(m^2))
    pre = np.ones_like(d)
    suf = np.ones_like(d)
    for l in range(1, m):
        pre[:, :, l, :] = pre[:, :, l-1, :]*d[:, :, l-1, :]
        suf[:, :, m-1-l, :] = suf[:, :, m-l, :]*d[:, :, m-l, :]
    omega = pre*suf
    omega = np.where(np.abs(d) <= 1e-9, beta[None, None, :, :], omega)
    # L_alpha(x_j) = prod_q omega[j, alpha_q, q] / prod_q beta[alpha_q, q]
    om = np.ones(x.shape[:2]+(idx.shape[0], ))
    be = np.ones(idx.shape[0])
    for q in range(dim):
        om = om*omega[:, :, idx[:, q], q]
        be = be*beta[idx[:, q], q]
    L = om/be[None, None, :]                                                    # [nc, nq, m^dim]
    vol = mesh.volVector[cells]
    V = np.zeros((dofs.shape[0], m**dim))
    sub = dm.dofs[cells]
    pos = np.searchsorted(dofs, np.maximum(sub, 0))
    pos = np.minimum(pos, dofs.shape[0]-1)
    ok = (sub >= 0) & (dofs[pos] == sub)
    for k in range(dim+1):
        contrib = vol[:, None]*np.einsum('cqa,q->ca', L, bary[k]*w)             # [nc, m^dim]
        np.add.at(V, pos[ok[:, k], k], contrib[ok[:, k]])
    return V


def dof_to_cells(dm):
    """CSR dof -> cells around it (ascending)"""
    m = dm.dofs >= 0
    d = dm.dofs[m]
    c = np.nonzero(m)[0]
    order = np.lexsort((c, d))
    d, c = d[order], c[order]
    ptr = np.zeros(dm.num_dofs+1, dtype=np.int64)
    np.add.at(ptr, d+1, 1)
    return np.cumsum(ptr), c


def cells_of_dofs(dm, dofs, d2c=None):
    ptr, cells = dof_to_cells(dm) if d2c is None else d2c
    cnt = ptr[dofs+1]-ptr[dofs]
    start = np.repeat(ptr[dofs]-np.concatenate(([0], np.cumsum(cnt)[:-1])), cnt)
    return np.unique(cells[start+np.arange(cnt.sum())])


def transfer_operator(parent, child):
    dim = parent.dim
    mC, mP = child.interpolation_order, parent.interpolation_order
    xiC, xiP = chebyshev(child.box, mC), chebyshev(parent.box, mP)
    omega = (xiC[:, None, :]-xiP[None, :, :]).prod(axis=1)            # [mC, dim]
    diff = xiP[:, None, :]-xiP[None, :, :]
    diff[np.arange(mP), np.arange(mP), :] = 1.
    beta = diff.prod(axis=1)                                          # [mP, dim]
    I, J = _multi_index(mP, dim), _multi_index(mC, dim)
    T = np.ones((I.shape[0], J.shape[0]))
    for l in range(dim):
        i, j = I[:, l][:, None], J[:, l][None, :]
        far = np.abs(xiP[i, l]-xiC[j, l]) > 1e-8
        with np.errstate(divide='ignore', invalid='ignore'):
            f = omega[j, l]/(xiC[j, l]-xiP[i, l])/beta[i, l]
        T = T*np.where(far, f, 1.)
    return T


def _csr(rows, cols, vals, shape, device):
    """device CSR matrix from COO triplets (duplicates summed)"""
    import warnings
    import torch
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        A = torch.sparse_coo_tensor(torch.as_tensor(np.vstack((rows, cols)), device=device), torch.as_tensor(vals, device=device),
                                    size=shape, dtype=torch.float64, device=device, check_invariants=False)
        return A.coalesce().to_sparse_csr()


class farFieldClusterPair:
    def __init__(self, n1, n2, kernelInterpolant):
        self.n1, self.n2, self.kernelInterpolant = n1, n2, kernelInterpolant

    def apply(self, x, y):
        """y += kernelInterpolant x (farFieldClusterPair.apply, clusterMethodCy.pyx:1968-1974); device tensors"""
        import torch
        if getattr(self, '_K', None) is None or self._K.device != x.device:
            self._K = torch.as_tensor(np.ascontiguousarray(self.kernelInterpolant), device=x.device)
        y += self._K.mv(x)


class H2Matrix:
    """y = Anear x + far field; `Anear` any object with a device matvec (or None), x / y float64 CUDA tensors"""

    def __init__(self, tree, Pfar, Anear, num_dofs, device):
        import torch
        self.tree, self.Pfar, self.Anear = tree, Pfar, Anear
        self.num_rows = self.num_columns = num_dofs
        self.device = device
        self._dev = {}
        tree._device = device
        for n in tree.get_tree_nodes():
            n._dofs_t = torch.as_tensor(n.dofs, device=device) if n.isLeaf else None

    shape = property(lambda self: (self.num_rows, self.num_columns))

    def _t(self, key, array):
        import torch
        if key not in self._dev:
            self._dev[key] = torch.as_tensor(np.ascontiguousarray(array), device=self.device)
        return self._dev[key]

    def farfield_device(self, x):
        """sum over the admissible cluster pairs, three passes on the device"""
        import torch
        y = torch.zeros(self.num_rows, dtype=torch.float64, device=self.device)

        def up(n):
            if n.isLeaf:
                n.coefficientsUp = self._t(('V', n.id), n.value).t().mv(x[n._dofs_t])
            else:
                acc = torch.zeros(n.interpolation_order**n.dim, dtype=torch.float64, device=self.device)
                for c in n.children:
                    up(c)
                    acc += self._t(('T', c.id), c.transferOperator).mv(c.coefficientsUp)
                n.coefficientsUp = acc
            n.coefficientsDown = torch.zeros(n.interpolation_order**n.dim, dtype=torch.float64, device=self.device)
        up(self.tree)
        for lvl in self.Pfar:
            for k, cp in enumerate(self.Pfar[lvl]):
                cp.n1.coefficientsDown += self._t(('K', lvl, k), cp.kernelInterpolant).mv(cp.n2.coefficientsUp)

        def down(n):
            if n.isLeaf:
                y[n._dofs_t] += self._t(('V', n.id), n.value).mv(n.coefficientsDown)
            else:
                for c in n.children:
                    c.coefficientsDown += self._t(('T', c.id), c.transferOperator).t().mv(n.coefficientsDown)
                    down(c)
        down(self.tree)
        return y

    # ---- batched form: the passes as a handful of sparse matrix-vector products --------------------------------
    def compile(self):
        """Stacks the coefficient vectors of all tree nodes into one vector and turns the three passes into sparse
        matrices over it: basis B (coefficients x dofs, leaf moments), one transfer matrix per level, the
        far-field matrix F.  y = B^T (down pass)(F (up pass)(B x)); a dozen SpMVs instead of thousands of tiny
        products."""
        import torch
        nodes = list(self.tree.get_tree_nodes())
        off = {}
        tot = 0
        for n in nodes:
            off[n.id] = tot
            tot += n.interpolation_order**n.dim
        self._ncoef = tot

        def coo(blocks, shape):
            if not blocks:
                return None
            r = np.concatenate([np.repeat(r0+np.arange(B.shape[0]), B.shape[1]) for r0, c0, B in blocks])
            c = np.concatenate([np.tile(c0 if isinstance(c0, np.ndarray) else c0+np.arange(B.shape[1]), B.shape[0]) for r0, c0, B in blocks])
            v = np.concatenate([np.asarray(B).ravel() for r0, c0, B in blocks])
            return _csr(r, c, v, shape, self.device)
        self._B = coo([(off[n.id], n.dofs, n.value.T) for n in nodes if n.isLeaf], (tot, self.num_rows))
        levels = {}
        for n in nodes:
            if n.parent is not None:
                levels.setdefault(n.parent.levelNo, []).append((off[n.parent.id], off[n.id], n.transferOperator))
        self._U = [(lvl, coo(levels[lvl], (tot, tot))) for lvl in sorted(levels, reverse=True)]
        self._Ut = [(lvl, coo([(c0, r0, np.ascontiguousarray(T.T)) for r0, c0, T in levels[lvl]], (tot, tot))) for lvl in sorted(levels)]
        self._F = coo([(off[cp.n1.id], off[cp.n2.id], cp.kernelInterpolant) for lvl in self.Pfar for cp in self.Pfar[lvl]], (tot, tot))
        # B^T: rows are dofs (scattered), columns the coefficient slots of the leaf
        lb = [(n.dofs, off[n.id], n.value) for n in nodes if n.isLeaf]
        r = np.concatenate([np.repeat(d, V.shape[1]) for d, c0, V in lb])
        c = np.concatenate([np.tile(c0+np.arange(V.shape[1]), V.shape[0]) for d, c0, V in lb])
        v = np.concatenate([V.ravel() for d, c0, V in lb])
        self._Bt = _csr(r, c, v, (self.num_rows, tot), self.device)
        self._compiled = True

    def farfield_compiled(self, x):
        import torch
        up = torch.mv(self._B, x)
        for _, U in self._U:            # deepest parents first
            up = up+torch.mv(U, up)
        down = torch.mv(self._F, up)
        for _, Ut in self._Ut:          # root first
            down = down+torch.mv(Ut, down)
        return torch.mv(self._Bt, down)

    # ---- device engine: own kernels for the three passes and the near field (pnb_h2_*, csrc/pnb_h2.cuh) -------------
    def build_engine(self, mesh=None, dm=None, leaf_values_on_device=True):
        """Hands the tree, the transfer operators, the far-field blocks and the near field (CSR) to libpnb200.  With a
        mesh the leaf moments V (enterLeafValues, clusterMethodCy.pyx:1205-1325) are computed by the device kernel and
        written back to `node.value`; otherwise the values already attached to the leaves are uploaded."""
        import ctypes
        from . import _lib
        nodes = list(self.tree.get_tree_nodes())
        index = {n.id: k for k, n in enumerate(nodes)}
        dim = self.tree.dim
        nn = len(nodes)
        i32 = lambda a: np.ascontiguousarray(a, dtype=np.int32)      # noqa: E731
        i64 = lambda a: np.ascontiguousarray(a, dtype=np.int64)      # noqa: E731
        f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)    # noqa: E731
        mm = np.array([n.interpolation_order**dim for n in nodes], dtype=np.int64)
        coef_ptr = i32(np.concatenate(([0], np.cumsum(mm))))
        parent = i32([index[n.parent.id] if n.parent is not None else -1 for n in nodes])
        level = i32([n.levelNo-self.tree.levelNo for n in nodes])
        leaves = [n for n in nodes if n.isLeaf]
        leaf_node = i32([index[n.id] for n in leaves])
        leaf_dof_ptr = i32(np.concatenate(([0], np.cumsum([n.dofs.shape[0] for n in leaves]))))
        leaf_dofs = i32(np.concatenate([n.dofs for n in leaves])) if leaves else i32([])
        if leaf_dofs.shape[0] != np.unique(leaf_dofs).shape[0]:
            raise ValueError('the leaves of the cluster tree must partition the dofs')
        keep = [coef_ptr, parent, level, leaf_node, leaf_dof_ptr, leaf_dofs]
        D = _lib.pnb_h2_desc_t()
        D.dim, D.num_dofs, D.num_nodes = dim, self.num_rows, nn
        D.coef_ptr, D.parent, D.level = coef_ptr.ctypes.data, parent.ctypes.data, level.ctypes.data
        D.num_leaves = len(leaves)
        D.leaf_node, D.leaf_dof_ptr, D.leaf_dofs = leaf_node.ctypes.data, leaf_dof_ptr.ctypes.data, leaf_dofs.ctypes.data
        on_device = leaf_values_on_device and mesh is not None
        if on_device:
            # cells around the dofs of every leaf (ascending) and the positions of their dofs in the leaf
            N = self.num_rows
            dof_leaf = np.full(N, -1, dtype=np.int64)
            dof_pos = np.zeros(N, dtype=np.int64)
            for k, n in enumerate(leaves):
                dof_leaf[n.dofs] = k
                dof_pos[n.dofs] = np.arange(n.dofs.shape[0])
            cd = np.asarray(dm.dofs)
            nv = cd.shape[1]
            cl = np.where(cd >= 0, dof_leaf[np.maximum(cd, 0)], -1)                 # leaf of every cell dof
            key = (cl.astype(np.int64)*mesh.num_cells+np.arange(mesh.num_cells)[:, None])[cl >= 0]
            key = np.unique(key)                                                    # (leaf, cell) ascending
            lc_leaf, lc_cell = key//mesh.num_cells, key % mesh.num_cells
            leaf_cell_ptr = i32(np.concatenate(([0], np.cumsum(np.bincount(lc_leaf, minlength=len(leaves))))))
            pos = np.where(cl[lc_cell] == lc_leaf[:, None], dof_pos[np.maximum(cd[lc_cell], 0)], -1)
            leaf_cells, leaf_cell_pos = i32(lc_cell), i32(pos.reshape(-1, nv))
            boxes = f64(np.stack([n.box for n in leaves])) if leaves else f64(np.zeros((0, dim, 2)))
            orders = i32([n.interpolation_order for n in leaves])
            max_m = int(orders.max()) if leaves else 1
            rule_n = np.zeros(max_m+1, dtype=np.int32)
            bptr, wptr = np.zeros(max_m+1, dtype=np.int64), np.zeros(max_m+1, dtype=np.int64)
            barys, ws = [], []
            for m in sorted(set(orders.tolist())):
                bary, w = quadrature.regular(m+2, dim)          # P1: quadOrder = order+2 (Sauter/Schwab p. 428)
                rule_n[m] = w.shape[0]
                bptr[m] = sum(b.size for b in barys)
                wptr[m] = sum(x.size for x in ws)
                barys.append(f64(bary).ravel())
                ws.append(f64(w))
            rule_bary, rule_w = f64(np.concatenate(barys)), f64(np.concatenate(ws))
            eta_ptr = np.zeros(max_m+2, dtype=np.int32)
            etas = []
            for m in range(max_m+1):
                eta_ptr[m] = sum(e.shape[0] for e in etas)
                etas.append(np.cos((2.0*np.arange(m, 0, -1)-1.0)/(2.0*m)*np.pi) if m > 0 else np.zeros(0))
            eta_ptr[max_m+1] = sum(e.shape[0] for e in etas)
            eta = f64(np.concatenate(etas))
            vertices, cells, vol = f64(mesh.vertices), i32(mesh.cells), f64(mesh.volVector)
            keep += [leaf_cell_ptr, leaf_cells, leaf_cell_pos, boxes, orders, rule_n, bptr, wptr, rule_bary, rule_w, eta_ptr, eta,
                     vertices, cells, vol]
            D.leaf_values = None
            D.leaf_cell_ptr, D.leaf_cells, D.leaf_cell_pos = leaf_cell_ptr.ctypes.data, leaf_cells.ctypes.data, leaf_cell_pos.ctypes.data
            D.leaf_boxes, D.leaf_orders = boxes.ctypes.data, orders.ctypes.data
            D.num_vertices, D.num_cells = mesh.num_vertices, mesh.num_cells
            D.vertices, D.cells, D.vol = vertices.ctypes.data, cells.ctypes.data, vol.ctypes.data
            D.max_m, D.rule_n, D.rule_bary_ptr, D.rule_w_ptr = max_m, rule_n.ctypes.data, bptr.ctypes.data, wptr.ctypes.data
            D.rule_bary, D.rule_w, D.rule_bary_size, D.rule_w_size = rule_bary.ctypes.data, rule_w.ctypes.data, rule_bary.size, rule_w.size
            D.eta, D.eta_ptr = eta.ctypes.data, eta_ptr.ctypes.data
        else:
            V = f64(np.concatenate([np.asarray(n.value).ravel() for n in leaves])) if leaves else f64([])
            keep.append(V)
            D.leaf_values = V.ctypes.data
        tsz = np.array([0 if n.parent is None else n.parent.interpolation_order**dim*n.interpolation_order**dim for n in nodes],
                       dtype=np.int64)
        tstart = np.concatenate(([0], np.cumsum(tsz)))
        transfer_ptr = i64(np.where(tsz > 0, tstart[:-1], -1))
        transfer = f64(np.concatenate([np.asarray(n.transferOperator).ravel() for n in nodes if n.parent is not None])
                       if nn > 1 else [])
        pairs = [cp for lvl in self.Pfar for cp in self.Pfar[lvl]]
        far_n1, far_n2 = i32([index[cp.n1.id] for cp in pairs]), i32([index[cp.n2.id] for cp in pairs])
        fsz = np.array([cp.kernelInterpolant.size for cp in pairs], dtype=np.int64)
        far_ptr = i64(np.concatenate(([0], np.cumsum(fsz))))
        far_blocks = f64(np.concatenate([np.asarray(cp.kernelInterpolant).ravel() for cp in pairs]) if pairs else [])
        keep += [transfer_ptr, transfer, far_n1, far_n2, far_ptr, far_blocks]
        D.transfer_ptr, D.transfer, D.transfer_size = transfer_ptr.ctypes.data, transfer.ctypes.data, transfer.size
        D.num_far, D.far_n1, D.far_n2 = len(pairs), far_n1.ctypes.data, far_n2.ctypes.data
        D.far_ptr, D.far_blocks, D.far_size = far_ptr.ctypes.data, far_blocks.ctypes.data, far_blocks.size
        csr = getattr(self.Anear, '_csr', None)
        if csr is not None:
            crow, col, val = csr.crow_indices(), csr.col_indices(), csr.values()
            keep += [crow, col, val]
            D.near_indptr, D.near_indices, D.near_data = crow.data_ptr(), col.data_ptr(), val.data_ptr()
        handle = ctypes.c_void_p()
        _lib.check(_lib.lib().pnb_h2_create(self.device.index if self.device.index is not None else 0, ctypes.byref(D),
                                            ctypes.byref(handle)))
        self._engine = handle
        self._engine_near = csr is not None
        if on_device:
            nV = int(sum(n.dofs.shape[0]*n.interpolation_order**dim for n in leaves))
            V = np.empty(nV)
            _lib.check(_lib.lib().pnb_h2_leaf_values(handle, V.ctypes.data))
            o = 0
            for n in leaves:
                sz = n.dofs.shape[0]*n.interpolation_order**dim
                n.value = V[o:o+sz].reshape(n.dofs.shape[0], -1)
                o += sz
        return self

    def __del__(self):
        try:
            if getattr(self, '_engine', None):
                from . import _lib
                _lib.lib().pnb_h2_destroy(self._engine)
                self._engine = None
        except Exception:
            pass

    def matvec_device(self, x, y=None):
        import torch
        if getattr(self, '_engine', None):
            from . import _lib
            from .linear_operators import _check_vector
            if y is None:
                y = torch.empty(self.num_rows, dtype=torch.float64, device=self.device)
            _check_vector(x, self.num_columns, self.device, 'x')
            _check_vector(y, self.num_rows, self.device, 'y')
            stream = torch.cuda.current_stream(self.device).cuda_stream
            _lib.check(_lib.lib().pnb_h2_matvec(self._engine, x.data_ptr(), y.data_ptr(), 0 if self._engine_near else 1, stream))
            if not self._engine_near and self.Anear is not None:
                y += self.Anear.matvec_device(x)
            return y
        out = self.farfield_compiled(x) if getattr(self, '_compiled', False) else self.farfield_device(x)
        if self.Anear is not None:
            out += self.Anear.matvec_device(x)
        if y is not None:
            y.copy_(out)
            return y
        return out

    def diagonal_device(self):
        """diagonal of the operator = diagonal of the near field (H2Matrix.diagonal, clusterMethodCy.pyx)"""
        return self.Anear.diagonal_device()

    diagonal = property(lambda self: self.diagonal_device().cpu().numpy())

    def matvec(self, x, y=None):
        import torch
        if isinstance(x, torch.Tensor):
            return self.matvec_device(x, y)
        r = self.matvec_device(torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=self.device)).cpu().numpy()
        if y is None:
            return r
        y[:] = r
        return y

    __call__ = matvec

    def __mul__(self, x):
        return self.matvec(x)

    def dot(self, x):
        return self.matvec(x)

    def __repr__(self):
        nfar = sum(len(v) for v in self.Pfar.values())
        return '<{}x{} H2Matrix {} tree nodes, {} far-field cluster pairs>'.format(self.num_rows, self.num_columns,
                                                                                     len(list(self.tree.get_tree_nodes())), nfar)


class nearFromDense:
    """near-field stand-in: the entries of the dense operator on the near cluster pairs (all other entries zero)"""

    def __init__(self, dense, Pnear):
        import torch
        A = dense.device_data
        mask = torch.zeros_like(A, dtype=torch.bool)
        for n1, n2 in Pnear:
            d1 = torch.as_tensor(n1.dofs, device=A.device)
            d2 = torch.as_tensor(n2.dofs, device=A.device)
            mask[d1[:, None], d2[None, :]] = True
        self.nnz = int(mask.sum())
        self._A = torch.where(mask, A, torch.zeros((), dtype=A.dtype, device=A.device))

    def matvec_device(self, x):
        return self._A.mv(x)


class _SubDoFMap:
    """P1 DoFMap of a cluster pair: the DoFs of the two clusters in local numbering, everything else Dirichlet"""
    polynomialOrder = 1

    def __init__(self, mesh, dofs, num_dofs):
        self.mesh, self.dofs, self.num_dofs = mesh, dofs, num_dofs
        self.dofs_per_element = dofs.shape[1]
        self.dim = mesh.dim


class nearFieldBlocks:
    """Near field as one dense block per near cluster pair, y[dofs(n1)] += B x[dofs(n2)]  (device tensors)."""

    def __init__(self, num_dofs, device):
        self.num_dofs, self.device = num_dofs, device
        self.blocks = []        # (rows tensor, cols tensor, block tensor)
        self.correction = None  # (rows, cols, values) of scattered entries added to the blocks
        self.nnz = 0

    def add(self, rows, cols, block):
        self.blocks.append((np.asarray(rows, dtype=np.int64), np.asarray(cols, dtype=np.int64), block))
        self.nnz += block.numel()

    def compile(self):
        """one CSR matrix on the device"""
        import torch
        # index arrays on the host (two uploads instead of thousands of tiny device operations)
        r = torch.as_tensor(np.concatenate([np.repeat(rr, cc.shape[0]) for rr, cc, B in self.blocks]), device=self.device)
        c = torch.as_tensor(np.concatenate([np.tile(cc, rr.shape[0]) for rr, cc, B in self.blocks]), device=self.device)
        v = torch.cat([B.reshape(-1) for rr, cc, B in self.blocks])
        if self.correction is not None:
            # entries added on top of the blocks (only where the pattern has an entry: DoFs of one cell always are
            # a near pair)
            cr, cc_, cv = (torch.as_tensor(a, device=self.device) for a in self.correction)
            r, c, v = torch.cat((r, cr)), torch.cat((c, cc_)), torch.cat((v, cv))
        import warnings
        with warnings.catch_warnings():
            warnings.simplefilter('ignore')
            coo = torch.sparse_coo_tensor(torch.stack((r, c)), v, size=(self.num_dofs, self.num_dofs),
                                          check_invariants=False).coalesce()
            self._csr = coo.to_sparse_csr()
        idx, val = coo.indices(), coo.values()
        on = idx[0] == idx[1]
        self._diag = torch.zeros(self.num_dofs, dtype=torch.float64, device=self.device)
        self._diag[idx[0][on]] = val[on]

    def diagonal_device(self):
        if getattr(self, '_diag', None) is None:
            self.compile()
        return self._diag

    def matvec_device(self, x):
        import torch
        if getattr(self, '_csr', None) is not None:
            return torch.mv(self._csr, x)
        y = torch.zeros(self.num_dofs, dtype=torch.float64, device=self.device)
        for r, c, B in self.blocks:
            y[torch.as_tensor(r, device=self.device)] += B.mv(x[torch.as_tensor(c, device=self.device)])
        if self.correction is not None:
            cr, cc_, cv = (torch.as_tensor(a, device=self.device) for a in self.correction)
            y.index_add_(0, cr, cv*x[cc_])
        return y

    def toarray(self):
        import torch
        A = torch.zeros((self.num_dofs, self.num_dofs), dtype=torch.float64, device=self.device)
        for r, c, B in self.blocks:
            A[torch.as_tensor(r, device=self.device)[:, None], torch.as_tensor(c, device=self.device)[None, :]] += B
        if self.correction is not None:
            cr, cc_, cv = (torch.as_tensor(a, device=self.device) for a in self.correction)
            A.index_put_((cr, cc_), cv, accumulate=True)
        return A.cpu().numpy()


class _BatchMesh:
    """disjoint union of sub-meshes over the vertices of the parent mesh (cells of block k are contiguous)"""
    manifold_dim = None

    def __init__(self, parent, cells, vol, h, bfacets):
        self.dim = self.manifold_dim = parent.dim
        self.vertices = parent.vertices
        self.num_vertices = parent.num_vertices
        self.cells = cells
        self.num_cells = cells.shape[0]
        self.volVector, self.hVector, self.boundaryFacets = vol, h, bfacets
        self.diam = parent.diam


def _block_boundary_facets(mesh, cells, block_of_cell, nblocks, device=None):
    """boundary facets of every block's sub-mesh (as meshNd.boundaryFacets finds them for one mesh: the edges that
    belong to one cell of the block, oriented as in that cell; 1D: the vertices that belong to one cell), block by block.
    Integer sorting only; runs on `device` when given (millions of edges for a large near field)."""
    import torch
    nv, ncv = mesh.num_vertices, cells.shape[0]
    dev = torch.device('cpu') if device is None else device
    C = torch.as_tensor(np.ascontiguousarray(cells), device=dev).to(torch.int64)
    B = torch.as_tensor(np.ascontiguousarray(block_of_cell), device=dev).to(torch.int64)
    if mesh.dim == 1:
        v = C.reshape(-1)
        Bv = B.repeat_interleave(2)
        key = Bv*nv+v
        _, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
        sel = torch.nonzero(cnt[inv] == 1).reshape(-1)
        # meshNd orders the boundary vertices of a mesh ascending; blocks are contiguous
        sel = sel[torch.argsort(key[sel], stable=True)]
        fac = v[sel].to(torch.int32).reshape(-1, 1)
        fb = Bv[sel]
    else:
        e = torch.cat((C[:, [0, 1]], C[:, [1, 2]], C[:, [2, 0]]))
        Be = B.repeat(3)
        lo, hi = torch.minimum(e[:, 0], e[:, 1]), torch.maximum(e[:, 0], e[:, 1])
        key = (Be*nv+lo)*nv+hi
        _, inv, cnt = torch.unique(key, return_inverse=True, return_counts=True)
        sel = torch.nonzero(cnt[inv] == 1).reshape(-1)        # ascending: edge type major, then cell order
        # per block (blocks are contiguous in the cell order): edge type major, then cell order -- the order
        # meshNd.boundaryFacets produces for the sub-mesh
        typ = sel//ncv
        sel = sel[torch.argsort(Be[sel]*3+typ, stable=True)]
        fac = e[sel].to(torch.int32)
        fb = Be[sel]
    fptr = np.zeros(nblocks+1, dtype=np.int32)
    np.cumsum(torch.bincount(fb, minlength=nblocks).cpu().numpy(), out=fptr[1:])
    return np.ascontiguousarray(fac.cpu().numpy()), fptr


def assemble_clusters(builder, Pnear, entries=False):
    """Near field of the H2 operator (assembleClusters, nonlocalAssembly_{SCALAR}.pxi:1663-1889, constant kernel).

    For a near cluster pair (n1, n2) the reference integrates the bilinear form over D x D, D = the cells around the
    DoFs of n1 and n2, and replaces the rest of the space by a surface integral over the boundary of D
    (:1840-1889); it keeps the entries (i in n1, j in n2).  That is the dense operator of the sub-mesh D with the
    zero-exterior surface terms on its own boundary, with the quadrature parameters of the whole problem.

    All cluster pairs are assembled by ONE batched device problem (round 2; round 1 ran one small problem per pair, 3.1 s
    at 12k DoFs, host-bound): the sub-meshes become the blocks of a disjoint union (pnb_mesh_t.num_blocks) whose cells
    never interact across blocks, every block with its own surface, and the DoF-tile kernels write one dense operator
    per block; the blocks (n1, n2) are gathered from them in one indexing operation.

    entries=True: the getEntry flavour (:1539-1660) -- for the regional operator (zeroExterior=False) the reference
    keeps the patch x patch part only, without surface terms and without the correction below."""
    import torch
    from .assembly import _Problem
    from . import _lib
    if not builder.params.get('near_field_batched', True):
        return _assemble_clusters_one_by_one(builder, Pnear, entries)
    mesh, dm = builder.mesh, builder.dm
    dev_index = builder.problem.device
    dev = torch.device('cuda', dev_index)
    out = nearFieldBlocks(dm.num_dofs, dev)
    d2c = dof_to_cells(dm)
    node_cells = {}

    def cells_of(n):
        if n.id not in node_cells:
            node_cells[n.id] = cells_of_dofs(dm, n.dofs, d2c)
        return node_cells[n.id]
    todo, seen = [], set()
    for n1, n2 in Pnear:
        if (n2.id, n1.id) not in seen and (n1.id, n2.id) not in seen:
            seen.add((n1.id, n2.id))
            todo.append((n1, n2))
    if not todo:
        return out
    align = int(_lib.lib().pnb_block_alignment())
    nb = len(todo)
    cell_lists, sdof_lists, gather = [], [], []
    cptr = np.zeros(nb+1, dtype=np.int64)
    dptr = np.zeros(nb+1, dtype=np.int64)
    out_off = 0
    sizes = []
    for k, (n1, n2) in enumerate(todo):
        d1, d2 = n1.dofs, n2.dofs
        union = np.union1d(d1, d2)
        cells = np.union1d(cells_of(n1), cells_of(n2))       # cellsUnion
        gd = dm.dofs[cells]
        pos = np.minimum(np.searchsorted(union, np.maximum(gd, 0)), union.shape[0]-1)
        inside = (gd >= 0) & (union[pos] == gd)
        n = union.shape[0]
        npad = -(-n//align)*align
        sdof_lists.append(np.where(inside, pos+dptr[k], -1))
        cell_lists.append(cells)
        cptr[k+1] = cptr[k]+cells.shape[0]
        dptr[k+1] = dptr[k]+npad
        r, c = np.searchsorted(union, d1), np.searchsorted(union, d2)
        gather.append((out_off+r[:, None]*npad+c[None, :]).ravel())
        sizes.append((d1.shape[0], d2.shape[0]))
        out_off += npad*npad
    if dptr[-1] >= 2**31 or cptr[-1] >= 2**31:
        raise MemoryError('near field batch too large for 32-bit indices; use params["near_field_batched"] = False')
    cells_all = np.concatenate(cell_lists)
    block_of_cell = np.repeat(np.arange(nb), np.diff(cptr))
    vcells = np.ascontiguousarray(mesh.cells[cells_all], dtype=np.int32)
    fac, fptr = _block_boundary_facets(mesh, vcells, block_of_cell, nb, dev)
    bmesh = _BatchMesh(mesh, vcells, np.ascontiguousarray(mesh.volVector[cells_all]), np.ascontiguousarray(mesh.hVector[cells_all]), fac)
    sdm = _SubDoFMap(bmesh, np.ascontiguousarray(np.concatenate(sdof_lists), dtype=np.int32), int(dptr[-1]))
    surface = 0 if (entries and not builder.zeroExterior) else 1
    prob = _Problem(sdm, builder.kernel, builder.kernelBoundary, builder.orders, dev_index, builder.problem.max_order,
                    order_num_dofs=dm.num_dofs, tables_from=builder.problem, blocks=(cptr, dptr, fptr))
    A = torch.empty(out_off, dtype=torch.float64, device=dev)
    _lib.check(_lib.lib().pnb_dense_assemble(prob.handle, surface, 0, sdm.num_dofs, A.data_ptr(), sdm.num_dofs, 1))
    vals = A[torch.as_tensor(np.concatenate(gather), device=dev)]
    del A, prob
    cache = {}
    o = 0
    for (n1, n2), (a, b) in zip(todo, sizes):
        cache[(n1.id, n2.id)] = vals[o:o+a*b].view(a, b)
        o += a*b
    for n1, n2 in Pnear:
        B = cache.get((n1.id, n2.id))
        if B is None:
            # the local matrices are symmetric: block (n2, n1)^T
            B = cache[(n2.id, n1.id)].t().contiguous()
        out.add(n1.dofs, n2.dofs, B)
    if not builder.zeroExterior and not entries:
        out.correction = _regional_correction(builder)
    return out


def _regional_correction(builder):
    """regional operator: the blocks hold the surface terms around the cluster unions, which stand for the whole
    complement of the union; take the part Omega^c out again (:1889-1912): minus the surface terms of the domain
    boundary, cell by cell, on the entries of the near pattern"""
    import ctypes
    from . import _lib
    mesh, dm = builder.mesh, builder.dm
    nv = mesh.dim+1
    D = np.zeros((mesh.num_cells, nv*(nv+1)//2))
    _lib.check(_lib.lib().pnb_boundary_cell_blocks(builder.problem.handle, D.ctypes.data_as(ctypes.c_void_p)))
    rows, cols, vals = [], [], []
    k = 0
    for p in range(nv):
        for q in range(p, nv):
            ok = (dm.dofs[:, p] >= 0) & (dm.dofs[:, q] >= 0) & (D[:, k] != 0)
            rows.append(dm.dofs[ok, p]); cols.append(dm.dofs[ok, q]); vals.append(-D[ok, k])
            if p != q:
                rows.append(dm.dofs[ok, q]); cols.append(dm.dofs[ok, p]); vals.append(-D[ok, k])
            k += 1
    return (np.concatenate(rows).astype(np.int64), np.concatenate(cols).astype(np.int64), np.concatenate(vals))


def _assemble_clusters_one_by_one(builder, Pnear, entries=False):
    """Near field of the H2 operator (assembleClusters, nonlocalAssembly_{SCALAR}.pxi:1663-1889, constant kernel).

    For a near cluster pair (n1, n2) the reference integrates the bilinear form over D x D, D = the cells around the
    DoFs of n1 and n2, and replaces the rest of the space by a surface integral over the boundary of D
    (:1840-1889); it keeps the entries (i in n1, j in n2).  That is the dense operator of the sub-mesh D with the
    zero-exterior surface terms on its own boundary, with the quadrature parameters of the whole problem, so the
    dense device path assembles it: one small problem per cluster pair, block rows n1 / columns n2 kept.

    entries=True: the getEntry flavour (:1539-1660) -- for the regional operator (zeroExterior=False) the reference
    keeps the patch x patch part only, without surface terms and without the correction below."""
    import torch
    from .assembly import _Problem
    from .linear_operators import Dense_LinearOperator
    from .mesh import meshNd
    from . import _lib
    mesh, dm = builder.mesh, builder.dm
    dev_index = builder.problem.device
    dev = torch.device('cuda', dev_index)
    out = nearFieldBlocks(dm.num_dofs, dev)
    d2c = dof_to_cells(dm)
    node_cells = {}
    cache = {}

    def cells_of(n):
        if n.id not in node_cells:
            node_cells[n.id] = cells_of_dofs(dm, n.dofs, d2c)
        return node_cells[n.id]
    # one small problem per unordered cluster pair; a few host threads keep several of them in flight (the C calls
    # release the GIL and run on per-thread CUDA streams)
    todo, seen = [], set()
    for n1, n2 in Pnear:
        if (n2.id, n1.id) not in seen and (n1.id, n2.id) not in seen:
            seen.add((n1.id, n2.id))
            todo.append((n1, n2))
    for n1, n2 in todo:
        cells_of(n1)
        cells_of(n2)

    surface = 0 if (entries and not builder.zeroExterior) else 1
    template = [builder.problem]      # same kernel, orders and tables: the sub-problems share its table structures

    def work(pair):
        n1, n2 = pair
        d1, d2 = n1.dofs, n2.dofs
        union = np.union1d(d1, d2)
        cells = np.union1d(node_cells[n1.id], node_cells[n2.id])       # cellsUnion
        gd = dm.dofs[cells]
        pos = np.minimum(np.searchsorted(union, np.maximum(gd, 0)), union.shape[0]-1)
        inside = (gd >= 0) & (union[pos] == gd)
        sub = meshNd(mesh.vertices, mesh.cells[cells])
        sdofs = np.ascontiguousarray(np.where(inside, pos, -1), dtype=np.int32)
        sdm = _SubDoFMap(sub, sdofs, union.shape[0])
        prob = _Problem(sdm, builder.kernel, builder.kernelBoundary, builder.orders, dev_index,
                        builder.problem.max_order, order_num_dofs=dm.num_dofs, tables_from=template[0])
        n = union.shape[0]
        # torch work of this thread on its own stream (the legacy default stream would serialise all threads); the C
        # call returns after its kernels have finished, so no further ordering is needed
        import threading
        tl = threading.current_thread()
        if not hasattr(tl, '_pnb_stream'):
            tl._pnb_stream = torch.cuda.Stream(dev)
        with torch.cuda.device(dev), torch.cuda.stream(tl._pnb_stream):
            A = torch.empty((n, n), dtype=torch.float64, device=dev)
            _lib.check(_lib.lib().pnb_dense_assemble(prob.handle, surface, 0, n, A.data_ptr(), A.stride(0), 1))
            r = torch.as_tensor(np.searchsorted(union, d1), device=dev)
            c = torch.as_tensor(np.searchsorted(union, d2), device=dev)
            B = A[r[:, None], c[None, :]].contiguous()
            tl._pnb_stream.synchronize()
        del prob
        return (n1.id, n2.id), B
    nthreads = int(builder.params.get('near_field_threads', 8))
    if nthreads > 1 and len(todo) > 1:
        from concurrent.futures import ThreadPoolExecutor
        with ThreadPoolExecutor(nthreads) as ex:
            cache = dict(ex.map(work, todo))
    else:
        cache = dict(map(work, todo))
    torch.cuda.synchronize(dev)
    for n1, n2 in Pnear:
        B = cache.get((n1.id, n2.id))
        if B is None:
            # the local matrices are symmetric: block (n2, n1)^T
            B = cache[(n2.id, n1.id)].t().contiguous()
        out.add(n1.dofs, n2.dofs, B)
    if not builder.zeroExterior and not entries:
        # regional operator: the blocks above hold the surface terms around the cluster unions, which stand for the
        # whole complement of the union; take the part Omega^c out again (:1889-1912): minus the surface terms of
        # the domain boundary, cell by cell, on the entries of the near pattern
        import ctypes
        nv = mesh.dim+1
        D = np.zeros((mesh.num_cells, nv*(nv+1)//2))
        _lib.check(_lib.lib().pnb_boundary_cell_blocks(builder.problem.handle, D.ctypes.data_as(ctypes.c_void_p)))
        rows, cols, vals = [], [], []
        k = 0
        for p in range(nv):
            for q in range(p, nv):
                ok = (dm.dofs[:, p] >= 0) & (dm.dofs[:, q] >= 0) & (D[:, k] != 0)
                rows.append(dm.dofs[ok, p]); cols.append(dm.dofs[ok, q]); vals.append(-D[ok, k])
                if p != q:
                    rows.append(dm.dofs[ok, q]); cols.append(dm.dofs[ok, p]); vals.append(-D[ok, k])
                k += 1
        out.correction = (np.concatenate(rows).astype(np.int64), np.concatenate(cols).astype(np.int64), np.concatenate(vals))
    return out
