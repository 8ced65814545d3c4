"""Operators returned by nonlocalBuilder.

Dense_LinearOperator mirrors the surface of
base/PyNucleus_base/DenseLinearOperator_{SCALAR}.pxi:8-93 and
LinearOperator_{SCALAR}.pxi:38-324 that callers of the assembly path use
(shape, num_rows/num_columns, matvec / __call__ / __mul__ / dot, toarray, data,
diagonal, T, getMemorySize).  The entries live in HBM (a torch tensor is the
allocation); ``matvec`` runs the hand-written FP64 kernel through
``pnb_dense_matvec``.  ``data`` copies to host memory on first use.
"""
import numpy as np
import torch

from . import _lib


def _check_vector(v, n, device, name):
    """the matvec kernel reads / writes contiguous float64 device memory through raw pointers"""
    if not isinstance(v, torch.Tensor) or v.dim() != 1 or v.shape[0] != n:
        raise ValueError('{}: need a 1-D tensor of length {}'.format(name, n))
    if v.dtype != torch.float64 or not v.is_cuda or v.device != device:
        raise ValueError('{}: need a float64 tensor on {} (got {} on {})'.format(name, device, v.dtype, v.device))
    if n > 1 and v.stride(0) != 1:
        raise ValueError('{}: need a contiguous tensor (stride {})'.format(name, v.stride(0)))


def check_matrix_out(out, rows, cols, device):
    """`out=` targets of the assembly entry points: the kernels write row-major float64 device memory"""
    if not isinstance(out, torch.Tensor) or out.dim() != 2 or tuple(out.shape) != (rows, cols):
        raise ValueError('out: need a ({}, {}) tensor'.format(rows, cols))
    if out.dtype != torch.float64 or not out.is_cuda or out.device != device:
        raise ValueError('out: need a float64 tensor on {} (got {} on {})'.format(device, out.dtype, out.device))
    if (cols > 1 and out.stride(1) != 1) or (rows > 1 and out.stride(0) < cols):
        raise ValueError('out: need row-major storage (strides {})'.format(tuple(out.stride())))


class Dense_LinearOperator:
    def __init__(self, data, device_index=None):
        if not isinstance(data, torch.Tensor):
            raise TypeError('Dense_LinearOperator holds device memory; use from_numpy() for host arrays')
        if not data.is_cuda or data.dtype != torch.float64 or data.dim() != 2 or data.stride(1) != 1:
            raise ValueError('need a 2-D row-major float64 CUDA tensor')
        self._A = data
        self.device_index = data.device.index if device_index is None else device_index
        self.num_rows, self.num_columns = int(data.shape[0]), int(data.shape[1])
        self._host = None

    @classmethod
    def from_numpy(cls, array, device=0):
        return cls(torch.as_tensor(np.ascontiguousarray(array, dtype=np.float64)).to('cuda:{}'.format(device)))

    shape = property(lambda self: (self.num_rows, self.num_columns))
    device_data = property(lambda self: self._A)

    @property
    def data(self):
        if self._host is None:
            self._host = self._A.cpu().numpy()
        return self._host

    def toarray(self):
        return self.data

    @property
    def diagonal(self):
        return torch.diagonal(self._A).cpu().numpy().copy()

    @property
    def T(self):
        return Dense_LinearOperator(self._A.t().contiguous())

    def isSparse(self):
        return False

    def getMemorySize(self):
        return self._A.numel()*8

    # ---- y = A x ---------------------------------------------------------
    def matvec_device(self, x, y=None):
        """x, y: float64 CUDA tensors on the operator's device"""
        if y is None:
            y = torch.empty(self.num_rows, dtype=torch.float64, device=self._A.device)
        _check_vector(x, self.num_columns, self._A.device, 'x')
        _check_vector(y, self.num_rows, self._A.device, 'y')
        if self.num_rows == 0 or self.num_columns == 0:
            return y.zero_()
        stream = torch.cuda.current_stream(self._A.device).cuda_stream
        _lib.check(_lib.lib().pnb_dense_matvec(self.device_index, self._A.data_ptr(), self.num_rows, self.num_columns,
                                               self._A.stride(0), x.data_ptr(), y.data_ptr(), stream))
        return y

    def matvec(self, x, y=None):
        """PyNucleus calling convention: A(x, y) / A.matvec(x, y) with host vectors"""
        if isinstance(x, torch.Tensor):
            return self.matvec_device(x, y)
        xd = torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64)).to(self._A.device)
        yd = self.matvec_device(xd)
        if y is None:
            return yd.cpu().numpy()
        y[:] = yd.cpu().numpy()
        return y

    __call__ = matvec

    def dot(self, x):
        return self.matvec(x)

    def __mul__(self, x):
        if np.isscalar(x):
            return Dense_LinearOperator(self._A*x)
        return self.matvec(x)

    def __rmul__(self, x):
        if np.isscalar(x):
            return Dense_LinearOperator(self._A*x)
        return NotImplemented

    def __add__(self, other):
        if isinstance(other, Dense_LinearOperator):
            return Dense_LinearOperator(self._A+other._A)
        return NotImplemented

    def toLinearOperator(self):
        from scipy.sparse.linalg import LinearOperator
        return LinearOperator(shape=self.shape, matvec=lambda x: self.matvec(x), dtype=np.float64)

    def __repr__(self):
        return '<{}x{} Dense_LinearOperator on cuda:{}>'.format(self.num_rows, self.num_columns, self.device_index)


class diagonalOperator:
    """Diagonal operator (base/PyNucleus_base/LinearOperator_{SCALAR}.pxi, diagonalOperator): ``data`` is the diagonal.
    Returned by nonlocalBuilder.getDiagonal(); N numbers, kept on the host (Jacobi scaling for the Krylov solvers)."""

    def __init__(self, diagonal):
        self.data = np.ascontiguousarray(diagonal, dtype=np.float64)
        self.num_rows = self.num_columns = int(self.data.shape[0])

    shape = property(lambda self: (self.num_rows, self.num_columns))
    diagonal = property(lambda self: self.data)

    def matvec(self, x, y=None):
        if isinstance(x, torch.Tensor):
            r = torch.as_tensor(self.data, device=x.device)*x
            if y is not None:
                y.copy_(r)
            return r if y is None else y
        r = self.data*np.asarray(x)
        if y is not None:
            y[:] = r
            return y
        return r

    __call__ = matvec
    dot = matvec
    __mul__ = matvec

    def toarray(self):
        return np.diag(self.data)

    def getEntry(self, i, j):
        return float(self.data[i]) if i == j else 0.

    def isSparse(self):
        return True

    def getMemorySize(self):
        return self.data.nbytes


class SSS_LinearOperator:
    """symmetric sparse operator in the reference's SSS layout (base/PyNucleus_base/SSS_LinearOperator_{SCALAR}.pxi):
    CSR arrays (`indptr`, `indices`, `data`, column indices ascending) whose entries below the diagonal carry the
    operator, plus `diagonal`; y = L x + L^T x + diagonal * x.  The arrays live on the device; the product runs as two
    sparse products and a diagonal scaling."""

    def __init__(self, indptr, indices, data, diagonal):
        import torch
        self.indptr, self.indices, self._data, self._diagonal = indptr, indices, data, diagonal
        n = diagonal.shape[0]
        self.shape = (n, n)
        self.num_rows = self.num_columns = n
        self.device = diagonal.device
        self._L = torch.sparse_csr_tensor(indptr.to(torch.int64), indices.to(torch.int64), data, size=(n, n))
        self._LT = self._L.to_sparse_coo().t().to_sparse_csr()

    @property
    def nnz(self):
        return int(self.indices.shape[0])

    @property
    def data(self):
        return self._data.cpu().numpy()

    @property
    def diagonal(self):
        return self._diagonal.cpu().numpy()

    def isSparse(self):
        return True

    def getMemorySize(self):
        return 8*(self.nnz+self.num_rows)+4*(self.nnz+self.num_rows+1)

    def matvec_device(self, x, y=None):
        import torch
        _check_vector(x, self.num_columns, self.device, 'x')
        r = torch.mv(self._L, x)+torch.mv(self._LT, x)+self._diagonal*x
        if y is None:
            return r
        y.copy_(r)
        return y

    def matvec(self, x, y=None):
        import torch
        r = self.matvec_device(torch.as_tensor(np.ascontiguousarray(x, dtype=np.float64), device=self.device)).cpu().numpy()
        if y is None:
            return r
        y[:] = r
        return y

    def dot(self, x):
        return self.matvec(x)

    def __mul__(self, x):
        return self.matvec(x)

    def toarray(self):
        import torch
        L = self._L.to_dense()
        return (L+L.t()+torch.diag(self._diagonal)).cpu().numpy()

    def __repr__(self):
        return '<{}x{} SSS_LinearOperator with {} stored entries>'.format(self.num_rows, self.num_columns, self.nnz)
