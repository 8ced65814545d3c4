"""Krylov loop on the device for operators assembled by nonlocalBuilder.

Mirrors cg_solver.solve (base/PyNucleus_base/solvers.pyx:364-445: preconditioned CG with the reference's
stopping rule on the preconditioned residual) for device-resident vectors; the operator application is the
hand-written FP64 matvec kernel (single GPU) or its row-block form followed by an NCCL all-gather of the
iterate (DistributedDenseOperator; the reference's distributed operators Bcast/Allreduce instead,
nl/PyNucleus_nl/clusterMethodCy.pyx:3136-3142).  BLAS-1 work (axpy, dot) uses torch.
"""
import numpy as np
import torch


class DistributedDenseOperator:
    """Rows of a dense operator per rank; matvec = local rows of A x, then all-gather of the pieces.

    all_rows[r]: global row indices (ascending) held by rank r -- contiguous blocks (1D) or the rows of the dofs of a
    range of cell groups (2D, nonlocalBuilder.getDenseDistributed).  A_rows: Dense_LinearOperator with
    len(all_rows[rank]) rows and num_dofs columns (None on a rank without rows)."""

    def __init__(self, A_rows, all_rows, rank, num_dofs, process_group=None):
        self.A_rows = A_rows
        self.all_rows = [np.asarray(r, dtype=np.int64) for r in all_rows]
        self.rank = rank
        self.rows = self.all_rows[rank]
        self.num_rows = self.num_columns = num_dofs
        self.group = process_group
        self._send = self._recv = self._perm = None
        self.device = A_rows.device_data.device if A_rows is not None else torch.device('cuda', torch.cuda.current_device())
        # contiguous blocks keep their bounds (1D row blocks)
        r = self.rows
        self.row_begin = int(r[0]) if r.shape[0] else 0
        self.row_end = int(r[-1])+1 if r.shape[0] else 0

    shape = property(lambda self: (self.num_rows, self.num_columns))

    def matvec_device(self, x, y=None):
        import torch.distributed as dist
        if y is None:
            y = torch.empty(self.num_rows, dtype=torch.float64, device=self.device)
        # the pieces differ in length: gather fixed-size slots, then scatter them to the global numbering
        slot = max(r.shape[0] for r in self.all_rows)
        if self._send is None:
            self._send = torch.zeros(slot, dtype=torch.float64, device=self.device)
            self._recv = torch.empty(slot*len(self.all_rows), dtype=torch.float64, device=self.device)
            src = np.concatenate([k*slot+np.arange(r.shape[0]) for k, r in enumerate(self.all_rows)])
            self._src = torch.as_tensor(src, device=self.device)
            self._dst = torch.as_tensor(np.concatenate(self.all_rows), device=self.device)
        n = self.rows.shape[0]
        if self.A_rows is not None and n > 0:
            self.A_rows.matvec_device(x, self._send[:n])
        if len(self.all_rows) > 1:
            dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        else:
            self._recv.copy_(self._send)
        y[self._dst] = self._recv[self._src]
        return y

    def diagonal_device(self):
        import torch.distributed as dist
        d = torch.zeros(self.num_rows, dtype=torch.float64, device=self.device)
        if self.A_rows is not None:
            A = self.A_rows.device_data
            rows = torch.as_tensor(self.rows, device=self.device)
            d[rows] = A[torch.arange(rows.shape[0], device=self.device), rows]
        if len(self.all_rows) > 1:
            dist.all_reduce(d, group=self.group)
        return d


def _device_of(A):
    return A.device_data.device if hasattr(A, 'device_data') else A.device


def _jacobi(A, dev):
    """1/diag(A) on the device (jacobi_solver / invDiagonal, base/PyNucleus_base/solvers.pyx:229-246)"""
    if hasattr(A, 'diagonal_device'):
        d = A.diagonal_device()
    elif hasattr(A, 'device_data'):
        d = torch.diagonal(A.device_data).clone()
    else:
        d = torch.as_tensor(np.ascontiguousarray(A.diagonal, dtype=np.float64)).to(dev)
    return 1.0/d


def _setup(A, b, x0, precond, jacobi):
    host = not isinstance(b, torch.Tensor)
    dev = _device_of(A)
    bt = torch.as_tensor(np.ascontiguousarray(b, dtype=np.float64)).to(dev) if host else b
    if x0 is None:
        x = torch.zeros_like(bt)
    else:
        x = torch.as_tensor(np.ascontiguousarray(x0, dtype=np.float64)).to(dev) if not isinstance(x0, torch.Tensor) else x0.clone()
    if precond is None and jacobi:
        Minv = _jacobi(A, dev)
        precond = lambda v: Minv*v        # noqa: E731
    return host, bt, x, precond


def cg(A, b, x0=None, tol=1e-8, maxiter=1000, jacobi=True, relative=False, use2norm=False, precond=None, fused=True):
    """Preconditioned conjugate gradients on the device, following cg_solver.solve step by step
    (base/PyNucleus_base/solvers.pyx:364-445): stopping rule on sqrt(<r, M^-1 r>) (or the 2-norm with use2norm),
    absolute tolerance or relative to the initial residual (relative=True, :296-301), residual recomputed every 50
    iterations.  b, x0: float64 CUDA tensors or numpy arrays (copied); precond: callable on device vectors (default:
    Jacobi when jacobi=True).  Returns (x, iterations, residuals) with the reference's return value and history."""
    if fused and precond is None and not hasattr(A, 'all_rows') and _device_of(A).type == 'cuda':
        # single GPU, no or diagonal preconditioner: the BLAS-1 of the iteration runs in the library's fused kernels
        return _cg_fused(A, b, x0, tol, maxiter, jacobi, relative, use2norm)
    host, bt, x, precond = _setup(A, b, x0, precond, jacobi)
    r = bt-A.matvec_device(x) if x0 is not None else bt.clone()
    if relative:
        tol = tol*float(torch.linalg.vector_norm(r))
    if precond is None:
        p = r.clone()
        beta_old = torch.dot(r, p)
        crit = float(torch.sqrt(beta_old))
    else:
        p = precond(r)
        beta_old = torch.dot(r, p)
        crit = float(torch.linalg.vector_norm(r)) if use2norm else float(torch.sqrt(torch.abs(beta_old)))
    res = [crit]
    its, k = maxiter, 0
    if crit <= tol:
        its = 0
    else:
        for i in range(maxiter):
            Ap = A.matvec_device(p)
            alpha = beta_old/torch.dot(p, Ap)
            x += alpha*p
            r -= alpha*Ap
            if k == 50:
                r = bt-A.matvec_device(x)
                k = 0
            if precond is None:
                nr = torch.linalg.vector_norm(r)
                crit = float(nr)
                beta = nr*nr
                z = r
            else:
                z = precond(r)
                beta = torch.dot(r, z)
                crit = float(torch.linalg.vector_norm(r)) if use2norm else float(torch.sqrt(torch.abs(beta)))
            res.append(crit)
            if crit <= tol:
                its = i
                break
            p = z+(beta/beta_old)*p
            beta_old = beta
            k += 1
    return (x.cpu().numpy() if host else x), its, res


def _cg_fused(A, b, x0, tol, maxiter, jacobi, relative, use2norm):
    """cg() with the vector updates and inner products of an iteration in three fused library kernels
    (pnb_krylov_cg_update / pnb_krylov_cg_direction, csrc/pnb_krylov.cuh): the scalars stay on the device, the host reads
    <r,z> and <r,r> once per iteration for the stopping rule.  Same recurrences, stopping rule and residual history as the
    loop above."""
    from . import _lib
    L = _lib.lib()
    host, bt, x, _ = _setup(A, b, x0, None, False)
    dev = bt.device
    n = bt.shape[0]
    index = dev.index if dev.index is not None else torch.cuda.current_device()
    Minv = _jacobi(A, dev).contiguous() if jacobi else None
    x = x.contiguous()
    r = (bt-A.matvec_device(x) if x0 is not None else bt.clone()).contiguous()
    if relative:
        tol = tol*float(torch.linalg.vector_norm(r))
    z = Minv*r if Minv is not None else r
    p = z.clone()
    work = torch.zeros(int(L.pnb_krylov_workspace_doubles()), dtype=torch.float64, device=dev)
    stream = lambda: torch.cuda.current_stream(dev).cuda_stream      # noqa: E731

    def dots():
        # work[2] = <r,z>, work[3] = <r,r>
        _lib.check(L.pnb_krylov_dot(index, n, r.data_ptr(), z.data_ptr(), work.data_ptr(), work[2:].data_ptr(), stream()))
        _lib.check(L.pnb_krylov_dot(index, n, r.data_ptr(), r.data_ptr(), work.data_ptr(), work[3:].data_ptr(), stream()))

    def criterion():
        rz, rr = work[2:4].tolist()
        if Minv is None or use2norm:
            return float(np.sqrt(rr))
        return float(np.sqrt(abs(rz)))
    dots()
    work[0] = work[2]
    crit = criterion()
    res = [crit]
    its, k = maxiter, 0
    Ap = torch.empty_like(p)
    if crit <= tol:
        its = 0
    else:
        for i in range(maxiter):
            try:
                Ap = A.matvec_device(p, Ap)
            except TypeError:
                Ap = A.matvec_device(p)
            _lib.check(L.pnb_krylov_cg_update(index, n, p.data_ptr(), Ap.data_ptr(), Minv.data_ptr() if Minv is not None else None,
                                              x.data_ptr(), r.data_ptr(), z.data_ptr(), work.data_ptr(), stream()))
            if k == 50:
                # residual recomputed from the iterate (solvers.pyx:420-424)
                r.copy_(bt-A.matvec_device(x))
                if Minv is not None:
                    torch.mul(Minv, r, out=z)
                dots()
                k = 0
            crit = criterion()
            res.append(crit)
            if crit <= tol:
                its = i
                break
            _lib.check(L.pnb_krylov_cg_direction(index, n, z.data_ptr(), p.data_ptr(), work.data_ptr(), stream()))
            k += 1
    return (x.cpu().numpy() if host else x), its, res


def gmres(A, b, x0=None, tol=1e-8, maxiter=30, restarts=1, jacobi=True, left=True, relative=False, precond=None):
    """Restarted GMRES on the device, following gmres_solver.solve (base/PyNucleus_base/solvers.pyx:504-660): modified
    Gram-Schmidt Arnoldi, Givens rotations, left or right preconditioning, residual history |gamma_i|, the same
    iteration count (sum over the cycles of the last Arnoldi index).  The Krylov basis and the operator stay on the
    device; the (maxiter+1) x maxiter Hessenberg matrix is rotated on the host (one transfer of a column per step).
    Returns (x, iterations, residuals)."""
    host, bt, x, precond = _setup(A, b, x0, precond, jacobi)
    dev, n = bt.device, bt.shape[0]
    do_l = precond is not None and left
    do_r = precond is not None and not left
    if relative:
        tol = tol*float(torch.linalg.vector_norm(bt-A.matvec_device(x) if x0 is not None else bt))
    eps = 1e-15
    Q = torch.empty((maxiter+1, n), dtype=torch.float64, device=dev)
    H = np.ones((maxiter+1, maxiter))
    c, s_, gamma, y = np.zeros(maxiter), np.zeros(maxiter), np.zeros(maxiter+1), np.zeros(maxiter+1)
    res, all_iter, breakout = [], 0, False
    for _ in range(restarts):
        if breakout:
            break
        r = bt-A.matvec_device(x)
        if do_l:
            r = precond(r)
        gamma[0] = float(torch.linalg.vector_norm(r))
        if not res:
            res.append(abs(gamma[0]))
        if abs(gamma[0]) < tol:
            break
        Q[0] = r/gamma[0]
        i = -1
        for i in range(maxiter):
            if do_l:
                r = precond(A.matvec_device(Q[i]))
            elif do_r:
                r = A.matvec_device(precond(Q[i]))
            else:
                r = A.matvec_device(Q[i]).clone()
            h = torch.empty(i+2, dtype=torch.float64, device=dev)
            for j in range(i+1):
                h[j] = torch.dot(Q[j], r)
                r -= h[j]*Q[j]
            h[i+1] = torch.linalg.vector_norm(r)
            H[:i+2, i] = h.cpu().numpy()
            if abs(H[i+1, i]) > eps:
                Q[i+1] = r/H[i+1, i]
            else:
                breakout = True
                break
            for j in range(i):
                rho, sigma = H[j, i], H[j+1, i]
                H[j, i] = c[j]*rho+s_[j]*sigma
                H[j+1, i] = -s_[j]*rho+c[j]*sigma
            beta = np.sqrt(H[i, i]**2+H[i+1, i]**2)
            c[i], s_[i] = H[i, i]/beta, H[i+1, i]/beta
            H[i, i] = beta
            gamma[i+1] = -s_[i]*gamma[i]
            gamma[i] = c[i]*gamma[i]
            res.append(abs(gamma[i+1]))
            if abs(gamma[i+1]) < tol:
                breakout = True
                break
        all_iter += i
        for j in range(i, -1, -1):
            y[j] = (gamma[j]-H[j, j+1:i+1].dot(y[j+1:i+1]))/H[j, j]
        upd = torch.as_tensor(y[:i+1], device=dev) @ Q[:i+1]
        x += precond(upd) if do_r else upd
    return (x.cpu().numpy() if host else x), all_iter, res


def lu(A, b):
    """direct solve with a dense operator on the device (solver 'lu' of the drivers, base/PyNucleus_base/solvers.pyx:80-120:
    LU with partial pivoting).  b: float64 CUDA tensor or numpy array (copied); returns x of the same kind."""
    host = not isinstance(b, torch.Tensor)
    dev = _device_of(A)
    bt = torch.as_tensor(np.ascontiguousarray(b, dtype=np.float64)).to(dev) if host else b
    x = torch.linalg.solve(A.device_data, bt)
    return x.cpu().numpy() if host else x
