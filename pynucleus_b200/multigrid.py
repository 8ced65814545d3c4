"""Geometric multigrid with dense level operators on the device (BASELINE config 3: cg-mg).

Follows multilevelSolver/PyNucleus_multilevelSolver:
  * buildRestriction_{1D,2D}_P1      restriction_1D_P1.pxi, restriction_2D_P1.pxi; P = R^T (restrictionProlongation.pyx:108)
  * jacobiSmoother                   smoothers_{SCALAR}.pxi:75-148  (x += omega D^-1 (b - A x), omega = 2/3, 1 + 1 sweeps)
  * multigrid.solveOnLevel / solve   multigrid_{SCALAR}.pxi:226-262, 280-377 (V / W cycle, LU on the coarsest level)
  * multigridPreconditioner.matvec   multigrid_{SCALAR}.pxi:478-491 (one cycle from a zero initial guess)
The hierarchy of the reference's driver (discretizedProblems.py: one dense operator per uniformly refined mesh) is
assembled level by level with nonlocalBuilder.getDense(); the level operators, smoother diagonals, transfer
operators (sparse) and all vectors stay on the device, the residuals use the FP64 matvec kernel.
"""
import numpy as np
import torch

V, W = 1, 2


def buildRestrictionProlongation(coarse_dm, fine_dm):
    """(R, P) as scipy CSR matrices for P1 on a uniformly refined mesh (sub-cells 2c, 2c+1 in 1D; 4c .. 4c+3 in 2D with
    sub-cell k holding coarse vertex k first, as mesh.refine() and the reference number them)"""
    import scipy.sparse as sp
    cd, fd = coarse_dm.dofs, fine_dm.dofs
    nc = cd.shape[0]
    c = np.arange(nc)
    if coarse_dm.dim == 1:
        # coarse vertex 0: fine (2c, 0) weight 1, (2c, 1) weight 1/2; coarse vertex 1: (2c, 1) 1/2, (2c+1, 1) 1
        trip = [(cd[:, 0], fd[2*c, 0], 1.0), (cd[:, 0], fd[2*c, 1], 0.5),
                (cd[:, 1], fd[2*c, 1], 0.5), (cd[:, 1], fd[2*c+1, 1], 1.0)]
    else:
        s0, s1, s2 = 4*c, 4*c+1, 4*c+2
        trip = [(cd[:, 0], fd[s0, 0], 1.0), (cd[:, 0], fd[s0, 1], 0.5), (cd[:, 0], fd[s0, 2], 0.5),
                (cd[:, 1], fd[s0, 1], 0.5), (cd[:, 1], fd[s1, 0], 1.0), (cd[:, 1], fd[s1, 1], 0.5),
                (cd[:, 2], fd[s0, 2], 0.5), (cd[:, 2], fd[s1, 1], 0.5), (cd[:, 2], fd[s2, 0], 1.0)]
    rows = np.concatenate([r for r, _, _ in trip])
    cols = np.concatenate([f for _, f, _ in trip])
    vals = np.concatenate([np.full(r.shape[0], v) for r, _, v in trip])
    ok = (rows >= 0) & (cols >= 0)
    rows, cols, vals = rows[ok].astype(np.int64), cols[ok].astype(np.int64), vals[ok]
    # entries are set, not added (enterData): a pair met from several cells keeps its single weight
    _, first = np.unique(rows*fine_dm.num_dofs+cols, return_index=True)
    R = sp.coo_matrix((vals[first], (rows[first], cols[first])), shape=(coarse_dm.num_dofs, fine_dm.num_dofs)).tocsr()
    R.sort_indices()
    return R, R.T.tocsr()


def hierarchy(mesh, noRef, kernel, params={}, zeroExterior=True):
    """levels [{'mesh', 'DoFMap', 'A'[, 'R', 'P']}] from `mesh` and its noRef uniform refinements, one dense operator per
    level (the driver's hierarchy, nl/PyNucleus_nl/discretizedProblems.py: 'Assembled matrices on level k')"""
    from .assembly import nonlocalBuilder
    from .dofmap import P1_DoFMap
    levels = []
    for k in range(noRef+1):
        if k > 0:
            mesh = mesh.refine()
        dm = P1_DoFMap(mesh)
        lvl = {'mesh': mesh, 'DoFMap': dm, 'A': nonlocalBuilder(dm, kernel, params, zeroExterior=zeroExterior).getDense()}
        if k > 0:
            lvl['R'], lvl['P'] = buildRestrictionProlongation(levels[-1]['DoFMap'], dm)
        levels.append(lvl)
    return levels


def _to_device_csr(M, dev):
    import warnings
    M = M.tocsr()
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')
        return torch.sparse_csr_tensor(torch.as_tensor(M.indptr.astype(np.int64)), torch.as_tensor(M.indices.astype(np.int64)),
                                           torch.as_tensor(M.data.astype(np.float64)), size=M.shape).to(dev)


class multigrid:
    """multigrid(levels, smoother=('jacobi', {'omega': 2/3})): levels[k]['A'] device operators with matvec_device and
    a diagonal, levels[k]['R'/'P'] scipy sparse (k >= 1)."""

    def __init__(self, levels, smoother=('jacobi', {'omega': 2.0/3.0}), cycle=V):
        if not isinstance(smoother, tuple):
            smoother = (smoother, {})
        if smoother[0] != 'jacobi':
            raise NotImplementedError('smoother {}'.format(smoother[0]))
        prm = {'omega': 2.0/3.0, 'presmoothingSteps': 1, 'postsmoothingSteps': 1}
        prm.update(smoother[1])
        self.omega, self.pre, self.post = prm['omega'], prm['presmoothingSteps'], prm['postsmoothingSteps']
        self.cycle = cycle
        self.levels = levels
        self.A = [lvl['A'] for lvl in levels]
        A = self.A[-1]
        self.device = A.device_data.device if hasattr(A, 'device_data') else A.device
        self.num_rows = A.shape[0]
        self.initialized = False
        self.tolerance, self.maxIter, self.residuals = 1e-5, 50, []

    def setup(self):
        dev = self.device
        self.R = [None]+[_to_device_csr(lvl['R'], dev) for lvl in self.levels[1:]]
        self.P = [None]+[_to_device_csr(lvl['P'], dev) for lvl in self.levels[1:]]
        self.invD = [None]
        for A in self.A[1:]:
            d = A.diagonal_device() if hasattr(A, 'diagonal_device') else torch.diagonal(A.device_data).clone()
            self.invD.append(self.omega/d)
        # coarsest level: LU (solverFactory 'lu', base/PyNucleus_base/solvers.pyx:80-120)
        A0 = self.A[0].device_data
        self._lu = torch.linalg.lu_factor(A0)
        self.initialized = True

    def _smooth(self, lvl, b, x, steps, simple):
        A, invD = self.A[lvl], self.invD[lvl]
        for _ in range(steps):
            r = b.clone() if simple else b-A.matvec_device(x)
            simple = False
            x += invD*r

    def solveOnLevel(self, lvl, b, x, simpleResidual=False):
        if lvl == 0:
            x.copy_(torch.linalg.lu_solve(self._lu[0], self._lu[1], b[:, None])[:, 0])
            return
        A = self.A[lvl]
        self._smooth(lvl, b, x, self.pre, simpleResidual)
        res = b-A.matvec_device(x)
        defect = torch.mv(self.R[lvl], res)
        sol = torch.zeros_like(defect)
        simple = True
        for _ in range(self.cycle):
            self.solveOnLevel(lvl-1, defect, sol, simple)
            simple = False
        x += torch.mv(self.P[lvl], sol)
        self._smooth(lvl, b, x, self.post, False)

    def solve(self, b, x0=None, tol=None, maxiter=None):
        """cycles until the 2-norm of the residual is <= tol; returns (x, iterations, residuals)"""
        if not self.initialized:
            self.setup()
        tol = self.tolerance if tol is None else tol
        maxiter = self.maxIter if maxiter is None else maxiter
        host = not isinstance(b, torch.Tensor)
        bt = torch.as_tensor(np.ascontiguousarray(b, dtype=np.float64)).to(self.device) if host else b
        x = torch.zeros_like(bt) if x0 is None else torch.as_tensor(x0).to(self.device).clone()
        A = self.A[-1]
        simple = x0 is None
        res = [float(torch.linalg.vector_norm(bt if simple else bt-A.matvec_device(x)))]
        it = 0
        while res[-1] > tol and it < maxiter:
            it += 1
            self.solveOnLevel(len(self.A)-1, bt, x, simple)
            simple = False
            res.append(float(torch.linalg.vector_norm(bt-A.matvec_device(x))))
        self.residuals = res
        return (x.cpu().numpy() if host else x), it, res

    def asPreconditioner(self, maxIter=1, cycle=V):
        """callable r -> approximate A^-1 r: maxIter cycles from a zero initial guess"""
        if not self.initialized:
            self.setup()

        def apply(r):
            y = torch.zeros_like(r)
            old, self.cycle = self.cycle, cycle
            simple = True
            for _ in range(maxIter):
                self.solveOnLevel(len(self.A)-1, r, y, simple)
                simple = False
            self.cycle = old
            return y
        return apply
