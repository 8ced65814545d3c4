"""ORACLE INFRASTRUCTURE: finite-horizon fixtures (tests/golden/finite_*.npz) produced by running the REFERENCE ITSELF
(the stub-built copy in oracle/_ref, see build_reference.sh).

    PYTHONPATH=oracle/_ref python oracle/refbuild/make_golden_finite.py

Reference code exercised: ball2_retriangulation (nl/PyNucleus_nl/interactionDomains.pyx:866-965, 406-826), the cut branch
of eval_distant (nonlocalOperator_{SCALAR}.pxi:790-847), the finite-horizon / integrable kernels
(kernelsCy.pyx:75-114, 273-359) and their scalings (kernelNormalization.pyx:70-104, 225-251), getDense
(nonlocalAssembly_{SCALAR}.pxi:1262-1473).  Nothing here is computed by this repository's own code.
"""
import os
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
OUT = os.path.join(ROOT, 'tests', 'golden')

from PyNucleus_fem.mesh import simpleInterval, uniform_disc  # noqa: E402
from PyNucleus_fem.DoFMaps import P1_DoFMap  # noqa: E402
from PyNucleus_fem.functions import constant  # noqa: E402
from PyNucleus_nl.kernels import getFractionalKernel, getIntegrableKernel  # noqa: E402
from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder  # noqa: E402
from PyNucleus_nl.fractionalOrders import constFractionalOrder  # noqa: E402
from PyNucleus_base.myTypes import REAL  # noqa: E402


def make_kernel(dim, ktype, s, horizon):
    if ktype == 'fractional':
        return getFractionalKernel(dim, constFractionalOrder(s), constant(horizon))
    return getIntegrableKernel(dim, ktype, constant(horizon))


def case(dim, noRef, ktype, s, horizon, name, max_cut_pairs=60):
    mesh = simpleInterval(-1, 1) if dim == 1 else uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    kernel = make_kernel(dim, ktype, s, horizon)
    params = {'target_order': 0.5} if dim == 2 else {}
    b = nonlocalBuilder(dm, kernel, params)
    A = np.array(b.getDense().data)
    lm = b.local_matrix
    nc = mesh.num_cells
    cells = np.array(mesh.cells)
    vertices = np.array(mesh.vertices)
    panels = np.zeros((nc, nc), dtype=np.int8)
    relpos = np.full((nc, nc), -1, dtype=np.int8)
    for c1 in range(nc):
        lm.setCell1_py(c1)
        s1 = np.ascontiguousarray(vertices[cells[c1]])
        for c2 in range(c1, nc):
            lm.setCell2_py(c2)
            panels[c1, c2] = lm.getPanelType()
            relpos[c1, c2] = kernel.interaction.getRelativePosition_py(s1, np.ascontiguousarray(vertices[cells[c2]]))
    # local matrices: the pairs cut by the horizon (a sample), plus a few interacting regular and touching pairs
    rng = np.random.RandomState(5)
    cut = [(c1, c2) for c1 in range(nc) for c2 in range(c1, nc) if panels[c1, c2] > 0 and relpos[c1, c2] == 2]
    inter = [(c1, c2) for c1 in range(nc) for c2 in range(c1, nc) if panels[c1, c2] > 0 and relpos[c1, c2] != 2]
    touch = [(c1, c2) for c1 in range(nc) for c2 in range(c1, nc) if -4 < panels[c1, c2] < 0]

    def pick(lst, n):
        if len(lst) <= n:
            return lst
        return [lst[i] for i in rng.choice(len(lst), n, replace=False)]
    pairs = pick(cut, max_cut_pairs)+pick(inter, 15)+pick(touch, 15)
    dpe = dm.dofs_per_element
    nloc = (2*dpe)*(2*dpe+1)//2
    contribs = np.zeros((len(pairs), nloc))
    ppanels = np.zeros(len(pairs), dtype=np.int32)
    contrib = np.zeros((nloc, 1), dtype=REAL)
    for n, (c1, c2) in enumerate(pairs):
        lm.setCell1_py(int(c1))
        lm.setCell2_py(int(c2))
        p = lm.getPanelType()
        ppanels[n] = p
        lm.eval_py(contrib, p)
        contribs[n] = contrib[:, 0]
    out = dict(vertices=vertices, cells=cells, dofs=np.array(dm.dofs), num_dofs=dm.num_dofs,
               hVector=np.array(mesh.hVector), volVector=np.array(mesh.volVector), h=mesh.h, hmin=mesh.hmin, diam=mesh.diam,
               A=A, kernel_type=ktype, s=s, horizon=horizon, scaling=kernel.scalingValue, singularity=kernel.singularityValue,
               target_order=lm.target_order, quad_order_diagonal=lm.quad_order_diagonal,
               panel_matrix=panels, relpos=relpos, pairs=np.array(pairs, dtype=np.int32), panels=ppanels, contribs=contribs,
               num_cut=len(cut))
    if dim == 2:
        out['boundaryEdges'] = np.array(mesh.boundaryEdges)
        out['quad_order_diagonalV'] = lm.quad_order_diagonalV
    else:
        out['boundaryVertices'] = np.array(mesh.boundaryVertices)
    # kernel values on both sides of the horizon
    xs = np.zeros((6, dim))
    ys = np.zeros((6, dim))
    ys[:, 0] = horizon*np.array([0.1, 0.5, 0.999, 1.0, 1.001, 1.5])
    out['kx'] = xs
    out['ky'] = ys
    out['kvals'] = np.array([kernel(xs[i], ys[i]) for i in range(6)])
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, 'N', dm.num_dofs, 'cells', nc, 'cut pairs', len(cut), 'remote', int((panels == -6).sum()), 'h', mesh.h,
          'max order', int(panels.max()))


if __name__ == '__main__':
    case(1, 5, 'fractional', 0.25, 0.5, 'finite_interval_frac0.25_r5')
    case(1, 5, 'fractional', 0.75, 0.3, 'finite_interval_frac0.75_r5')
    case(1, 5, 'constant', 0., 0.4, 'finite_interval_constant_r5')
    case(1, 5, 'inverseDistance', 0., 0.4, 'finite_interval_invdist_r5')
    case(2, 3, 'fractional', 0.75, 0.7, 'finite_disc_frac0.75_r3')
    case(2, 3, 'fractional', 0.25, 0.6, 'finite_disc_frac0.25_r3')
    case(2, 3, 'constant', 0., 0.7, 'finite_disc_constant_r3')
    case(2, 3, 'inverseDistance', 0., 0.7, 'finite_disc_invdist_r3')
    case(2, 4, 'fractional', 0.4, 0.45, 'finite_disc_frac0.4_r4', max_cut_pairs=40)


def sparsified_case(dim, noRef, ktype, s, horizon, name):
    """getDense(trySparsification=True) (nonlocalAssembly_{SCALAR}.pxi:1287-1348): symmetric sparse (SSS) operator whose
    pattern holds every DoF pair of the cell pairs that are not ignored"""
    mesh = simpleInterval(-1, 1) if dim == 1 else uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    kernel = make_kernel(dim, ktype, s, horizon)
    params = {'target_order': 0.5} if dim == 2 else {}
    A = nonlocalBuilder(dm, kernel, params).getDense(trySparsification=True)
    out = dict(vertices=np.array(mesh.vertices), cells=np.array(mesh.cells), dofs=np.array(dm.dofs), num_dofs=dm.num_dofs,
               kernel_type=ktype, s=s, horizon=horizon, target_order=0.5, operator_type=type(A).__name__,
               indptr=np.array(A.indptr), indices=np.array(A.indices), data=np.array(A.data), diagonal=np.array(A.diagonal),
               volume=mesh.volume)
    if dim == 2:
        out['boundaryEdges'] = np.array(mesh.boundaryEdges)
    else:
        out['boundaryVertices'] = np.array(mesh.boundaryVertices)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, type(A).__name__, A.shape, 'nnz', A.nnz)


if __name__ == '__main__' and os.environ.get('SPARSIFIED', '1') == '1':
    sparsified_case(2, 4, 'fractional', 0.4, 0.3, 'sparsified_disc_frac0.4_r4')
    sparsified_case(1, 6, 'constant', 0., 0.25, 'sparsified_interval_constant_r6')


def rows_case(noRef, ktype, s, horizon, name, nrows=16):
    """larger mesh (disc, N = 2977 at noRef 5): sampled rows, the diagonal and products of the reference's operator"""
    mesh = uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    kernel = make_kernel(2, ktype, s, horizon)
    A = np.array(nonlocalBuilder(dm, kernel, {'target_order': 0.5}).getDense().data)
    rng = np.random.default_rng(9)
    rows = np.sort(rng.choice(dm.num_dofs, nrows, replace=False))
    x = rng.standard_normal(dm.num_dofs)
    out = dict(vertices=np.array(mesh.vertices), cells=np.array(mesh.cells), dofs=np.array(dm.dofs), num_dofs=dm.num_dofs,
               boundaryEdges=np.array(mesh.boundaryEdges), kernel_type=ktype, s=s, horizon=horizon, target_order=0.5,
               scaling=kernel.scalingValue, singularity=kernel.singularityValue,
               rows=rows, A_rows=A[rows], diagonal=np.diag(A).copy(), x=x, Ax=A.dot(x), frobenius=np.linalg.norm(A),
               nonzeros=int(np.count_nonzero(A)))
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, dm.num_dofs, 'nonzeros %.1f%%' % (100.*np.count_nonzero(A)/A.size))


if __name__ == '__main__' and os.environ.get('ROWS', '0') == '1':
    rows_case(5, 'fractional', 0.4, 0.3, 'finite_disc_frac0.4_r5_rows')
    rows_case(5, 'constant', 0., 0.25, 'finite_disc_constant_r5_rows')
