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
    """Row block of a dense operator per rank; matvec = local rows of A x, then all-gather."""

    def __init__(self, A_rows, row_begin, row_end, num_dofs, blocks, process_group=None):
        self.A_rows = A_rows
        self.row_begin, self.row_end = row_begin, row_end
        self.num_rows = self.num_columns = num_dofs
        self.blocks = blocks
        self.group = process_group
        self._send = self._recv = None
        self.device = A_rows.device_data.device if A_rows is not None else torch.device('cuda', torch.cuda.current_device())

    shape = property(lambda self: (self.num_rows, self.num_columns))

    def matvec_device(self, x, y=None):
        import torch.distributed as dist
        if y is None:
            y = torch.empty(self.num_rows, dtype=torch.float64, device=self.device)
        # blocks may differ in length (64-row granularity): gather fixed-size slots, then unpack
        slot = max(b-a for a, b in self.blocks)
        if self._send is None:
            self._send = torch.zeros(slot, dtype=torch.float64, device=self.device)
            self._recv = torch.empty(slot*len(self.blocks), dtype=torch.float64, device=self.device)
        n = self.row_end-self.row_begin
        if self.A_rows is not None and n > 0:
            self.A_rows.matvec_device(x, self._send[:n])
        dist.all_gather_into_tensor(self._recv, self._send, group=self.group)
        for r, (a, b) in enumerate(self.blocks):
            y[a:b] = self._recv[r*slot:r*slot+b-a]
        return y

    def diagonal_device(self):
        import torch.distributed as dist
        d = torch.zeros(self.num_rows, dtype=torch.float64, device=self.device)
        if self.A_rows is not None:
            A = self.A_rows.device_data
            idx = torch.arange(self.row_end-self.row_begin, device=self.device)
            d[self.row_begin:self.row_end] = A[idx, idx+self.row_begin]
        dist.all_reduce(d, group=self.group)
        return d


def cg(A, b, x0=None, tol=1e-8, maxiter=1000, jacobi=True):
    """Preconditioned conjugate gradients, x and b float64 CUDA tensors (or numpy arrays, copied).
    Returns (x, iterations, residual norms)."""
    host = not isinstance(b, torch.Tensor)
    dev = A.device_data.device if hasattr(A, 'device_data') else A.device
    bt = torch.as_tensor(np.ascontiguousarray(b, dtype=np.float64)).to(dev) if host else b
    x = torch.zeros_like(bt) if x0 is None else (torch.as_tensor(x0).to(dev) if host else x0.clone())
    if jacobi:
        d = A.diagonal_device() if hasattr(A, 'diagonal_device') else torch.diagonal(A.device_data).clone()
        Minv = 1.0/d
    r = bt-A.matvec_device(x)
    z = Minv*r if jacobi else r
    p = z.clone()
    rz = torch.dot(r, z)
    res = [float(torch.sqrt(torch.abs(rz)))]
    it = 0
    while it < maxiter and res[-1] > tol:
        Ap = A.matvec_device(p)
        alpha = rz/torch.dot(p, Ap)
        x += alpha*p
        r -= alpha*Ap
        z = Minv*r if jacobi else r
        rz_new = torch.dot(r, z)
        p = z+(rz_new/rz)*p
        rz = rz_new
        res.append(float(torch.sqrt(torch.abs(rz))))
        it += 1
    return (x.cpu().numpy() if host else x), it, res
