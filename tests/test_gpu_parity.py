"""GPU parity tests: the CUDA path (through the C ABI / nonlocalBuilder) against
the reference's goldens and against the oracle on the same inputs.

bit-exact: pair classification (panel type, quadrature order, permutations)
1e-12 relative: local matrices and assembled entries (summation order differs)
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES_2D = ['disc_s0.75_r1', 'disc_s0.75_r2', 'disc_s0.25_r2', 'disc_s0.75_r3']
CASES_1D = ['interval_s0.25_r3', 'interval_s0.25_r6', 'interval_s0.75_r5']
TOL = 1e-12


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name+'.npz'))


def builder_from_golden(g, zeroExterior=True):
    import pynucleus_b200 as pb
    dim = g['vertices'].shape[1]
    bf = g['boundaryEdges'] if dim == 2 else g['boundaryVertices'].reshape(-1, 1)
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=bf)
    dm = pb.P1_DoFMap(mesh)
    assert dm.num_dofs == int(g['num_dofs'])
    assert np.array_equal(np.where(dm.dofs >= 0, dm.dofs, -1), np.where(g['dofs'] >= 0, g['dofs'], -1))
    kernel = pb.getFractionalKernel(dim, float(g['s']))
    params = {'target_order': float(g['target_order'])} if dim == 2 else {}
    return pb.nonlocalBuilder(dm, kernel, params, zeroExterior=zeroExterior)


def entry_err(A, Aref):
    """max entry error relative to max(|Aref_ij|, 1e-2*sqrt(Aref_ii*Aref_jj)).

    Entries are sums of positive and negative pair contributions of size up
    to ~A_ii; the few near-diagonal entries that cancel to < 1e-2 of the
    diagonal scale carry the summation-order rounding of the large terms, so
    they are held to 1e-12 of that scale (i.e. an absolute 1e-14*A_ii),
    everything else to 1e-12 of the entry itself."""
    d = np.sqrt(np.abs(np.diag(Aref)))
    scale = np.maximum(np.abs(Aref), 1e-2*np.outer(d, d))
    return (np.abs(A-Aref)/scale).max()


def relerr_rows(C, Cref):
    scale = np.abs(Cref).max(axis=1, keepdims=True)
    scale[scale == 0] = 1.
    return (np.abs(C-Cref)/scale).max()


@pytest.mark.parametrize('name', CASES_2D+CASES_1D)
def test_classification_bit_exact(golden_dir, name):
    g = load(golden_dir, name)
    b = builder_from_golden(g)
    panel, p1, p2 = b.getPanelTypes(g['pairs'], returnPerms=True)
    assert np.array_equal(panel, g['panels'])
    t = g['panels'] < 0
    assert np.array_equal(p1[t], g['perm1'][t])
    assert np.array_equal(p2[t], g['perm2'][t])
    bpanel = b.getPanelTypes(g['bpairs'], boundary=True)
    assert np.array_equal(bpanel, g['bpanels'])


@pytest.mark.parametrize('name', ['disc_s0.75_r2', 'interval_s0.25_r6'])
def test_all_pairs_classification(golden_dir, name):
    g = load(golden_dir, name)
    b = builder_from_golden(g)
    nc = g['cells'].shape[0]
    iu = np.triu_indices(nc)
    panel = b.getPanelTypes(np.stack(iu, axis=1))
    assert np.array_equal(panel, g['panel_matrix'][iu])
    hist = b.getPanelHistogram()
    ref = {int(k): int(v) for k, v in zip(*np.unique(g['panel_matrix'][iu], return_counts=True))}
    assert hist == ref


@pytest.mark.parametrize('name', CASES_2D+CASES_1D)
def test_local_matrices_vs_reference(golden_dir, name):
    g = load(golden_dir, name)
    b = builder_from_golden(g)
    panel, C = b.getLocalMatrices(g['pairs'])
    assert np.array_equal(panel, g['panels'])
    assert relerr_rows(C, g['contribs']) < TOL
    bpanel, bC = b.getLocalMatrices(g['bpairs'], boundary=True)
    assert np.array_equal(bpanel, g['bpanels'])
    assert relerr_rows(bC, g['bcontribs']) < TOL


@pytest.mark.parametrize('name', CASES_2D)
def test_far_evaluator_vs_reference(golden_dir, name):
    """thread-per-pair factored evaluator on the low-order regular pairs"""
    import pynucleus_b200._lib as L
    g = load(golden_dir, name)
    b = builder_from_golden(g)
    far = (g['panels'] >= 1) & (g['panels'] <= L.lib().pnb_far_max_order())
    if not far.any():
        pytest.skip('no low-order pairs in this fixture')
    panel, C = b.getLocalMatrices(g['pairs'][far], path=1)
    assert np.array_equal(panel, g['panels'][far])
    assert relerr_rows(C, g['contribs'][far]) < TOL


@pytest.mark.parametrize('name', CASES_2D+CASES_1D)
def test_dense_vs_reference(golden_dir, name):
    g = load(golden_dir, name)
    A = builder_from_golden(g).getDense().data
    Aref = g['A']
    assert entry_err(A, Aref) < TOL
    A0 = builder_from_golden(g, zeroExterior=False).getDense().data
    assert entry_err(A0, g['A_interior']) < TOL
    assert np.array_equal(A, A.T)


def test_dense_host_entry_point(golden_dir):
    g = load(golden_dir, 'disc_s0.75_r2')
    b = builder_from_golden(g)
    A = b.getDenseHost()
    assert np.array_equal(A, b.getDense().data)


def test_order_retry_extends_the_tables(golden_dir):
    """tables supplied up to an order below what the mesh needs: the library reports PNB_ERR_ORDER, the builder extends the
    tables to pnb_max_order (the reference grows its rule cache lazily, addQuadRule) and the result is unchanged;
    device, pageable-host and pinned-host outputs"""
    import torch
    g = load(golden_dir, 'disc_s0.75_r3')
    ref = builder_from_golden(g).getDense().data
    b = builder_from_golden(g)
    b.params['max_regular_order'] = 3
    assert b.problem.max_order == 3
    A = b.getDense().data
    assert b.problem.max_order > 3 and np.array_equal(A, ref)
    b2 = builder_from_golden(g)
    b2.params['max_regular_order'] = 3
    assert np.array_equal(b2.getDenseHost(), ref)
    b3 = builder_from_golden(g)
    b3.params['max_regular_order'] = 3
    pinned = torch.empty(ref.shape, dtype=torch.float64).pin_memory()
    assert entry_err(b3.getDenseHost(out=pinned.numpy()), ref) < 1e-14


def test_dense_host_pinned_buffer_overlapped_copy():
    """pinned host buffer: the rows are copied panel by panel while the assembly continues, the entries that receive
    cell-diagonal blocks are written last -- the result is bitwise the device operator"""
    import torch
    import pynucleus_b200 as pb
    mesh = pb.refined(pb.polygon_disc(10), 4)
    dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
    D = b.getDense().data
    host = torch.empty((dm.num_dofs, dm.num_dofs), dtype=torch.float64).pin_memory()
    host.fill_(float('nan'))
    H = b.getDenseHost(out=host.numpy()).copy()
    # panel-wise launches interleave the two unit kernels differently than the device-resident path: same terms, another
    # (fixed) summation order in the entries that both kernels touch
    assert not np.isnan(H).any() and entry_err(H, D) < 1e-14
    assert np.array_equal(H, H.T)
    assert np.array_equal(b.getDenseHost(out=host.numpy()), H)          # bitwise reproducible
    # a wider leading dimension is honoured
    wide = torch.empty((dm.num_dofs, dm.num_dofs+8), dtype=torch.float64).pin_memory()
    wide.fill_(-1.)
    import ctypes
    from pynucleus_b200 import _lib
    _lib.check(_lib.lib().pnb_dense_assemble(b.problem.handle, 1, 0, dm.num_dofs, wide.data_ptr(), dm.num_dofs+8, 0))
    W = wide.numpy()
    assert np.array_equal(W[:, :dm.num_dofs], H) and (W[:, dm.num_dofs:] == -1.).all()


@pytest.mark.parametrize('noRef,s', [(4, 0.75), (5, 0.75), (4, 0.3)])
def test_dense_vs_oracle_disc(noRef, s):
    import oracle
    import pynucleus_b200 as pb
    mesh = pb.refined(pb.uniform_disc(), noRef)
    dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, s), {'target_order': 0.5})
    A = b.getDense().data
    P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, s, bfacets=mesh.boundaryFacets, target_order=0.5)
    Aref = P.dense(True)
    assert entry_err(A, Aref) < TOL
    st = b.getStats()
    assert st['distinct_pairs'] == P.last_npairs
    # bit-exact classification on a random sample of pairs at this size
    rng = np.random.RandomState(noRef)
    pairs = np.sort(rng.randint(0, mesh.num_cells, size=(20000, 2)), axis=1)
    assert np.array_equal(b.getPanelTypes(pairs), P.pairs(pairs, with_contrib=False)[0])


@pytest.mark.parametrize('noRef,s', [(8, 0.25), (10, 0.75)])
def test_dense_vs_oracle_interval(noRef, s):
    import oracle
    import pynucleus_b200 as pb
    mesh = pb.refined(pb.simpleInterval(-1., 1.), noRef)
    dm = pb.P1_DoFMap(mesh)
    A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(1, s), {}).getDense().data
    P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, s, bfacets=mesh.boundaryFacets)
    Aref = P.dense(True)
    assert entry_err(A, Aref) < TOL


def test_nonuniform_mesh_vs_oracle():
    """ragged input: perturbed vertices, shuffled cell order, rotated cell vertex order"""
    import oracle
    import pynucleus_b200 as pb
    rng = np.random.RandomState(7)
    m0 = pb.refined(pb.uniform_disc(), 3)
    v = m0.vertices.copy()
    interior = np.ones(v.shape[0], dtype=bool)
    interior[m0.boundaryVertices] = False
    v[interior] += 0.02*rng.randn(interior.sum(), 2)
    cells = m0.cells[rng.permutation(m0.num_cells)]
    rot = rng.randint(0, 3, size=cells.shape[0])
    cells = np.stack([cells[np.arange(cells.shape[0]), (rot+k) % 3] for k in range(3)], axis=1)
    mesh = pb.meshNd(v, cells)
    dm = pb.P1_DoFMap(mesh)
    A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.6), {'target_order': 0.5}).getDense().data
    P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, 0.6, bfacets=mesh.boundaryFacets, target_order=0.5)
    Aref = P.dense(True)
    assert entry_err(A, Aref) < TOL


def test_deterministic_and_symmetric():
    import pynucleus_b200 as pb
    mesh = pb.refined(pb.uniform_disc(), 5)
    dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
    A1 = b.getDense().data.copy()
    A2 = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5}).getDense().data
    assert np.array_equal(A1, A2)          # bitwise reproducible: no floating point atomics
    assert np.array_equal(A1, A1.T)
    # SPD-ness of the fractional Laplacian with zero exterior (energy norm known answer is a driver-level test)
    w = np.linalg.eigvalsh(A1)
    assert w.min() > 0


def test_matvec():
    import torch
    import pynucleus_b200 as pb
    rng = np.random.RandomState(0)
    for n, m in ((1, 1), (37, 37), (1000, 1003), (4097, 2050)):
        A = rng.randn(n, m)
        x = rng.randn(m)
        op = pb.Dense_LinearOperator.from_numpy(A)
        y = op*x
        assert np.abs(y-A.dot(x)).max() <= 1e-13*np.abs(A).sum(axis=1).max()*max(1., np.abs(x).max())
        yd = op.matvec_device(torch.as_tensor(x).cuda())
        assert np.array_equal(yd.cpu().numpy(), y)


def test_energy_known_answer_interval():
    """tests/test_fracLapl.py:30-58 of the reference: energy of the solution of (-Lap)^s u = 1 on (-1,1)"""
    from scipy.special import gamma
    import pynucleus_b200 as pb
    s = 0.25
    mesh = pb.refined(pb.simpleInterval(-1., 1.), 7)
    dm = pb.P1_DoFMap(mesh)
    A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(1, s), {}).getDense().data
    # P1 load vector for f = 1
    b = np.zeros(dm.num_dofs)
    vol = mesh.volVector
    for k in range(2):
        m = dm.dofs[:, k] >= 0
        np.add.at(b, dm.dofs[m, k], vol[m]/2.)
    u = np.linalg.solve(A, b)
    energy = b.dot(u)
    exact = 2.**(-2.*s)*np.pi/(gamma(0.5+s)*gamma(s+1.5))
    assert abs(energy-exact)/exact < 0.02


@pytest.mark.parametrize('nblocks', [2, 3])
def test_row_blocks_equal_full_operator(nblocks):
    """row-block (multi-GPU) assembly on one GPU: several problem instances, each owning a row block; the
    per-cell diagonal blocks are summed by hand (what the NCCL all-reduce does across ranks)"""
    import torch
    import pynucleus_b200 as pb
    from pynucleus_b200 import _lib
    from pynucleus_b200.assembly import row_partition
    mesh = pb.refined(pb.uniform_disc(), 4)
    dm = pb.P1_DoFMap(mesh)
    N = dm.num_dofs
    kernel = pb.getFractionalKernel(2, 0.75)
    full = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5}).getDense().data
    L = _lib.lib()
    blocks = row_partition(N, nblocks, L.pnb_row_granularity())
    builders = [pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5}) for _ in blocks]
    outs = [torch.empty((b-a, N), dtype=torch.float64, device='cuda') for a, b in blocks]
    Dsum = torch.zeros(mesh.num_cells*6, dtype=torch.float64, device='cuda')
    for bld, (a, b), out in zip(builders, blocks, outs):
        _lib.check(L.pnb_dense_rows_begin(bld.problem.handle, 1, a, b, out.data_ptr(), out.stride(0)))
        D = torch.empty_like(Dsum)
        _lib.check(L.pnb_dense_cell_blocks_copy(bld.problem.handle, D.data_ptr(), 0))
        torch.cuda.synchronize()
        # disjoint supports
        assert int(((D != 0) & (Dsum != 0)).sum()) == 0
        Dsum += D
    for bld, (a, b), out in zip(builders, blocks, outs):
        _lib.check(L.pnb_dense_cell_blocks_copy(bld.problem.handle, Dsum.data_ptr(), 1))
        _lib.check(L.pnb_dense_rows_end(bld.problem.handle, a, b, out.data_ptr(), out.stride(0)))
    A = torch.cat(outs, dim=0).cpu().numpy()
    assert entry_err(A, full) < TOL
    assert np.array_equal(A, A.T)


@pytest.mark.parametrize('name', ['disc_varconst0.75_r2', 'disc_varconst0.4_r3'])
def test_varconst_kernel_vs_reference(golden_dir, name):
    """BASELINE config 4: the reference's variable-order code path with s(x,y) = const"""
    import pynucleus_b200 as pb
    g = load(golden_dir, name)
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'])
    dm = pb.P1_DoFMap(mesh)
    kernel = pb.getFractionalKernel(2, pb.variableConstFractionalOrder(float(g['s'])))
    assert kernel.variable and kernel.symmetric
    A = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5}).getDense().data
    assert entry_err(A, g['A']) < TOL


@pytest.mark.parametrize('name', ['h2_disc_s0.75_r4', 'h2_interval_s0.25_r8'])
def test_farfield_blocks_vs_reference(golden_dir, name):
    """H2 far-field Chebyshev kernel blocks (clusterMethodCy.pyx:2153-2238) against the reference's kernelInterpolant"""
    import pynucleus_b200 as pb
    g = load(golden_dir, name)
    dim = g['vertices'].shape[1]
    bf = g['boundaryEdges'] if dim == 2 else g['boundaryVertices'].reshape(-1, 1)
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=bf)
    dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(dim, float(g['s'])), {'target_order': 0.5} if dim == 2 else {})
    blocks = b.getFarFieldBlocks(g['far_box1'], g['far_box2'], g['far_m1'], g['far_m2'])
    ptr = g['far_ptr']
    worst = 0.
    for k, blk in enumerate(blocks):
        ref = g['far_blocks'][ptr[k]:ptr[k+1]].reshape(blk.shape)
        worst = max(worst, (np.abs(blk-ref)/np.abs(ref)).max())
    assert worst < 1e-14


def _p1_load_vector(mesh, dm, fun=None, rule=None):
    """b_i = int f phi_i (f = 1 when fun is None, exact); rule = (bary[nvc, n], w[n])"""
    nvc = mesh.dim+1
    b = np.zeros(dm.num_dofs)
    bary, w = rule if rule is not None else (None, None)
    for k in range(nvc):
        m = dm.dofs[:, k] >= 0
        if fun is None:
            val = mesh.volVector/nvc
        else:
            pts = np.einsum('qk,ckd->cqd', bary.T, mesh.vertices[mesh.cells])      # cells x nodes x dim
            val = mesh.volVector*np.einsum('q,cq->c', w*bary[k], fun(pts))
        np.add.at(b, dm.dofs[m, k], val[m])
    return b


def _p1_mass(mesh, dm):
    nvc = mesh.dim+1
    M = np.zeros((dm.num_dofs, dm.num_dofs))
    for a in range(nvc):
        for c in range(nvc):
            m = (dm.dofs[:, a] >= 0) & (dm.dofs[:, c] >= 0)
            fac = 2. if a == c else 1.
            np.add.at(M, (dm.dofs[m, a], dm.dofs[m, c]), fac*mesh.volVector[m]/((nvc)*(nvc+1)))
    return M


def test_driver_golden_disc_constant_forcing():
    """Driver-level known answers of the reference (tests/cache_runFractional.py--domaindisc--sconst(0.75)--
    problemconstant--elementP1--solvercg-mg--matrixFormatdense, noRef 5 -> 2977 DoFs): Hs error 0.0603196,
    L2 error 0.00225634, compared with the reference's own tolerance rTol = 3e-2
    (nl/PyNucleus_nl/discretizedProblems.py:225-241).  Error definitions: discretizedProblems.py:77-110,
    exact values nonlocalProblems.py:741-749.  Solved with CG on the device (dense matvec kernel)."""
    from scipy.special import gamma
    import pynucleus_b200 as pb
    s = 0.75
    mesh = pb.refined(pb.uniform_disc(), 5)
    dm = pb.P1_DoFMap(mesh)
    assert dm.num_dofs == 2977
    A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, s), {'target_order': 0.5}).getDense()
    b = _p1_load_vector(mesh, dm)
    u, its, res = pb.cg(A, b, tol=1e-10, maxiter=2000)
    assert res[-1] <= 1e-10
    C = 2.**(-2.*s)*gamma(1.)/gamma(1.+s)/gamma(1.+s)
    exactHs2 = C*np.pi/(s+1)
    Hs_error = np.sqrt(abs(b.dot(u)-exactHs2))
    assert abs(Hs_error-0.060319591944560894) <= 3e-2*0.060319591944560894

    def u_exact(x):
        return C*np.maximum(1.-(x**2).sum(axis=-1), 0.)**s
    # the reference integrates z with its default P1 rule, the edge-midpoint rule (fem/PyNucleus_fem/femCy.pyx:2648-2650,
    # quadrature.pyx:279-282); the L2 number it caches includes that quadrature error, so use the same rule
    midpoints = (np.array([[0.5, 0.0, 0.5], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5]]), np.full(3, 1./3.))
    z = _p1_load_vector(mesh, dm, u_exact, rule=midpoints)
    M = _p1_mass(mesh, dm)
    L2_ex2 = C**2*np.pi/(1+2*s)
    L2_error = np.sqrt(abs(L2_ex2-2*z.dot(u)+u.dot(M.dot(u))))
    assert abs(L2_error-0.002256341047519089) <= 3e-2*0.002256341047519089
    # energy known answer of the reference's tests/test_fracLapl.py:60-77 (disc): 2 pi 2^{-2s} / (Gamma(1+s)^2 2(s+1))
    assert abs(b.dot(u)-2*np.pi*2.**(-2*s)/(gamma(1+s)**2*2*(s+1))) < 0.35*exactHs2


def test_sampled_rows_at_2977_dofs(golden_dir):
    """disc, 5 refinements (N = 2977, the reference's driver-test size): CUDA assembly against sampled rows,
    the diagonal, and seeded products of the reference's own getDense output."""
    import pynucleus_b200 as pb
    g = load(golden_dir, 'disc_s0.75_r5_rows')
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'])
    dm = pb.P1_DoFMap(mesh)
    Aop = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, float(g['s'])), {'target_order': 0.5}).getDense()
    A = Aop.toarray()
    d = np.sqrt(g['diagonal'])
    scale = np.maximum(np.abs(g['A_rows']), 1e-2*np.outer(d[g['rows']], d))
    assert (np.abs(A[g['rows']]-g['A_rows'])/scale).max() < TOL
    assert np.abs(np.diag(A)/g['diagonal']-1).max() < TOL
    assert np.abs(Aop*g['x']-g['Ax']).max() < TOL*np.abs(g['Ax']).max()
    assert np.abs(A.dot(np.ones(A.shape[0]))-g['ones_Ax']).max() < TOL*np.abs(g['diagonal']).max()
    assert abs(np.linalg.norm(A)/float(g['frobenius'])-1) < TOL


def test_two_gpus_row_blocks_matvec_cg():
    """N > 1 on real GPUs (skipped on a one-GPU box): tests/multi_gpu_worker.py under torchrun, NCCL"""
    import subprocess
    import sys
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip('needs 2 GPUs')
    here = os.path.dirname(os.path.abspath(__file__))
    cmd = [sys.executable, '-m', 'torch.distributed.run', '--nnodes=1', '--nproc-per-node', '2',
           '--master-addr', '127.0.0.1', '--master-port', str(29800+os.getpid() % 100),
           os.path.join(here, 'multi_gpu_worker.py')]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0 and 'OK' in out.stdout, out.stdout[-2000:]+out.stderr[-4000:]


@pytest.mark.parametrize('nparts,sides,noRef', [(2, 6, 4), (3, 6, 4), (4, 10, 5)])
def test_distributed_parts_equal_full_operator(nparts, sides, noRef):
    """several GPUs, 2D (pnb_dist_*), emulated on ONE GPU: one problem instance per part, all staging buffers in this
    process (on several GPUs they are peer memory).  Every part evaluates its units into the staging buffers of the row
    owners, the cell-diagonal blocks are summed (what the all-reduce does), every part builds its rows.  Every pair is
    evaluated exactly once over all parts, the rows of the parts partition the dofs, and the rows equal the single-GPU
    operator."""
    import ctypes
    import torch
    import pynucleus_b200 as pb
    from pynucleus_b200 import _lib
    mesh = pb.refined(pb.polygon_disc(sides), noRef)
    dm = pb.P1_DoFMap(mesh)
    N = dm.num_dofs
    kernel = pb.getFractionalKernel(2, 0.75)
    ref = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5})
    full = ref.getDense().data
    distinct = ref.getStats()['distinct_pairs']
    L = _lib.lib()
    builders = [pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5}) for _ in range(nparts)]
    rows, stages = [], []
    for k, bld in enumerate(builders):
        nrows, nstage = ctypes.c_int32(0), ctypes.c_int64(0)
        _lib.check(L.pnb_dist_plan(bld.problem.handle, nparts, k, ctypes.byref(nrows), ctypes.byref(nstage)))
        r = np.empty(nrows.value, dtype=np.int32)
        _lib.check(L.pnb_dist_rows(bld.problem.handle, r.ctypes.data))
        rows.append(r)
        # NaN-filled: a fragment that is read without having been written shows up in the result
        stages.append(torch.full((max(int(nstage.value), 1),), float('nan'), dtype=torch.float64, device='cuda'))
    assert np.array_equal(np.sort(np.concatenate(rows)), np.arange(N))
    ptrs = (ctypes.c_void_p*nparts)(*[t.data_ptr() for t in stages])
    Dsum = torch.zeros(mesh.num_cells*6, dtype=torch.float64, device='cuda')
    for bld in builders:
        _lib.check(L.pnb_dist_eval(bld.problem.handle, 1, ptrs))
        need = ctypes.c_int32(0)
        _lib.check(L.pnb_dist_status(bld.problem.handle, ctypes.byref(need)))
        assert need.value == 0
        D = torch.empty_like(Dsum)
        _lib.check(L.pnb_dense_cell_blocks_copy(bld.problem.handle, D.data_ptr(), 0))
        torch.cuda.synchronize()
        Dsum += D
    A = np.zeros((N, N))
    evaluated = 0
    for bld, r in zip(builders, rows):
        out = torch.empty((r.shape[0], N), dtype=torch.float64, device='cuda')
        _lib.check(L.pnb_dense_cell_blocks_copy(bld.problem.handle, Dsum.data_ptr(), 1))
        _lib.check(L.pnb_dist_apply(bld.problem.handle, 1, out.data_ptr(), out.stride(0)))
        evaluated += bld.getStats()['evaluated_pairs']
        A[r] = out.cpu().numpy()
    assert evaluated == distinct
    assert entry_err(A, full) < TOL
    assert np.abs(A-A.T).max() <= 1e-13*np.abs(A).max()
    # work balance of the plan: no part evaluates more than 1.35 times its share
    shares = [b.getStats()['evaluated_pairs'] for b in builders]
    assert max(shares) < 1.35*distinct/nparts, shares


def test_distributed_operator_single_process():
    """getDenseDistributed without a process group (world size 1): same code path as on several GPUs, own staging buffer"""
    import torch
    import pynucleus_b200 as pb
    mesh = pb.refined(pb.uniform_disc(), 3)
    dm = pb.P1_DoFMap(mesh)
    kernel = pb.getFractionalKernel(2, 0.75)
    full = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5}).getDense()
    b = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5})
    op = b.getDenseDistributed()
    assert np.array_equal(op.rows, np.arange(dm.num_dofs))
    assert entry_err(op.A_rows.data, full.data) < TOL
    x = torch.from_numpy(np.random.default_rng(1).standard_normal(dm.num_dofs)).cuda()
    assert float((op.matvec_device(x)-full.matvec_device(x)).abs().max()) < 1e-12*float(full.matvec_device(x).abs().max())
    b.releaseScratch()


def test_group_path_equals_tile_path():
    """2D: the cell-group kernels (default) against the DoF-tile kernels (params['assembly_path'] = 'tiles'), which share
    only the per-pair evaluators: same operator to rounding, both bitwise symmetric"""
    import pynucleus_b200 as pb
    mesh = pb.refined(pb.polygon_disc(7), 3)
    mesh.vertices[:] = mesh.vertices*np.array([1.3, 0.8])      # anisotropic: several mesh sizes and orders
    mesh = pb.meshNd(mesh.vertices, mesh.cells)
    dm = pb.P1_DoFMap(mesh)
    kernel = pb.getFractionalKernel(2, 0.4)
    A = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5}).getDense().data
    b = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5, 'assembly_path': 'tiles'})
    B = b.getDense().data
    assert b.getStats()['evaluated_pairs'] > b.getStats()['distinct_pairs']      # the tile path ran (halo pairs)
    assert entry_err(A, B) < TOL
    assert np.array_equal(A, A.T) and np.array_equal(B, B.T)


@pytest.mark.parametrize('name', ['disc_dm2_s0.75_r2', 'disc_dm2_s0.25_r3'])
def test_two_dofmaps_vs_reference(golden_dir, name):
    """nonlocalBuilder(dm, kernel, dm2=dm.getComplementDoFMap()).getDense(): the interior x boundary block that the
    drivers use for inhomogeneous Dirichlet data (nonlocalAssembly_{SCALAR}.pxi:1366-1378)"""
    import pynucleus_b200 as pb
    g = load(golden_dir, name)
    # boundary vertices in the reference's order (= order of the negative DoF numbers)
    neg = g['dofs'] < 0
    bv = np.empty(int(g['num_dofs2']), dtype=np.int32)
    bv[-g['dofs'][neg]-1] = g['cells'][neg]
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'], boundaryVertices=bv)
    dm = pb.P1_DoFMap(mesh)
    assert np.array_equal(dm.dofs, g['dofs'])
    dm2 = dm.getComplementDoFMap()
    assert np.array_equal(dm2.dofs, g['dofs2']) and dm2.num_dofs == int(g['num_dofs2'])
    kernel = pb.getFractionalKernel(2, float(g['s']))
    for ze, key in ((True, 'A_bc'), (False, 'A_bc_interior')):
        A = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5}, zeroExterior=ze, dm2=dm2).getDense()
        assert A.shape == g[key].shape
        assert np.abs(A.data-g[key]).max() < TOL*np.abs(g[key]).max()


@pytest.mark.parametrize('name', ['h2_interval_s0.25_r8', 'h2_disc_s0.75_r4'])
def test_h2_operator(golden_dir, name):
    """getH2: far-field part against the reference's H2 matvec (H x - Anear x of its getH2), whole operator against
    the dense one (the near field here is the dense operator on the near cluster pairs)"""
    import scipy.sparse as sp
    import torch
    import pynucleus_b200 as pb
    g = load(golden_dir, name)
    b = builder_from_golden(g)
    H, Pnear, tree = b.getH2(returnNearField=True, returnTree=True)
    assert np.array_equal(np.array([(a.id, c.id) for a, c in Pnear]), g['near_pairs'])
    x = torch.as_tensor(g['x']).cuda()
    N = b.dm.num_dofs
    low = sp.csr_matrix((g['Anear_data'], g['Anear_indices'], g['Anear_indptr']), shape=(N, N))
    ref_far = g['Hx']-(low+low.T+sp.diags(g['Anear_diagonal'])).dot(g['x'])
    yfar = H.farfield_device(x).cpu().numpy()
    assert np.abs(yfar-ref_far).max() < 1e-11*np.abs(ref_far).max()
    # near field: one block per near cluster pair against the reference's SSS matrix (strict lower triangle + diagonal)
    An = H.Anear.toarray()
    ref_near = (low+low.T+sp.diags(g['Anear_diagonal'])).toarray()
    assert np.array_equal(An != 0, ref_near != 0)
    assert np.abs(An-ref_near).max() < TOL*np.abs(ref_near).max()
    # the whole operator against the reference's H2 matvec and against the dense operator
    Hx = H*g['x']
    assert np.abs(Hx-g['Hx']).max() < 1e-11*np.abs(g['Hx']).max()
    A = b.getDense()
    err = np.abs(Hx-A*g['x']).max()/np.abs(A*g['x']).max()
    ref_err = np.abs(g['Hx']-g['Ax']).max()/np.abs(g['Ax']).max()
    assert err < 2*ref_err+1e-12


def test_h2_against_dense_operator_larger_mesh():
    """a size the reference's golden does not cover (N = 2977): H2 matvec against the dense matvec (interpolation error only),
    near-field entries equal to the dense entries there where the near pattern is a full cluster block on the diagonal"""
    import torch
    import pynucleus_b200 as pb
    mesh = pb.refined(pb.uniform_disc(), 5)
    dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
    H, Pnear = b.getH2(returnNearField=True)
    A = b.getDense()
    x = torch.as_tensor(np.sin(np.arange(dm.num_dofs)*0.37)+0.1).cuda()
    y, yd = H.matvec_device(x), A.matvec_device(x)
    assert float((y-yd).abs().max()) < 1e-4*float(yd.abs().max())
    assert H.Anear.nnz < 0.3*dm.num_dofs**2 and sum(len(v) for v in H.Pfar.values()) > 100


@pytest.mark.parametrize('name', ['disc_leftright_r2', 'disc_leftright_r3'])
def test_piecewise_variable_order_vs_reference(golden_dir, name):
    """getDense with leftRightFractionalOrder: per cell pair the order of its label class, assembled as a
    sum of constant-order passes restricted to the pairs of each class"""
    import pynucleus_b200 as pb
    g = load(golden_dir, name)
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'])
    dm = pb.P1_DoFMap(mesh)
    s = pb.leftRightFractionalOrder(float(g['sll']), float(g['srr']), float(g['slr']), float(g['slr']), float(g['interface']))
    kernel = pb.getFractionalKernel(2, s)
    assert kernel.variable and kernel.symmetric
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        A = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5}, zeroExterior=ze).getDense().data
        assert entry_err(A, g[key]) < TOL
        assert np.array_equal(A, A.T)


def test_piecewise_variable_order_larger_mesh_vs_oracle():
    """a mesh with uniformly far units on both sides of the interface (all kernels, label-uniform and mixed groups)"""
    import oracle
    import pynucleus_b200 as pb
    mesh = pb.refined(pb.polygon_disc(7), 4)
    dm = pb.P1_DoFMap(mesh)
    s = pb.leftRightFractionalOrder(0.3, 0.7, 0.45, 0.45, 0.1)
    A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, s), {'target_order': 0.5}).getDense().data
    svals, pc = s.classes()
    labels = s.labels(mesh.vertices[mesh.cells].mean(axis=1))
    blabels = s.labels(mesh.vertices[mesh.boundaryFacets].mean(axis=1))
    ref = 0.
    for k, sv in enumerate(svals):
        P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, sv, bfacets=mesh.boundaryFacets, target_order=0.5,
                           s_max=max(svals), labels=labels, blabels=blabels, pair_class=pc, active_class=k, max_order=40)
        ref = ref+P.dense(True)
    assert entry_err(A, ref) < TOL


@pytest.mark.parametrize('noRef', [0, 1])
def test_tiny_meshes_vs_oracle(noRef):
    """one and seven DoFs: a single cell group, every pair touching or near"""
    import oracle
    import pynucleus_b200 as pb
    mesh = pb.refined(pb.uniform_disc(), noRef)
    dm = pb.P1_DoFMap(mesh)
    A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5}).getDense().data
    P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, 0.75, bfacets=mesh.boundaryFacets, target_order=0.5)
    ref = P.dense(True)
    assert A.shape == ref.shape == (dm.num_dofs, dm.num_dofs)
    assert np.abs(A-ref).max() < TOL*np.abs(ref).max()


@pytest.mark.parametrize('name', ['entry_disc_s0.75_r3', 'entry_interval_s0.25_r6'])
def test_single_entries_and_diagonal_vs_reference(golden_dir, name):
    """getEntry / getDiagonal (nonlocalAssembly_{SCALAR}.pxi:1539-1660, 2269-2289) against the reference's values"""
    import pynucleus_b200 as pb
    g = load(golden_dir, name)
    dim = g['vertices'].shape[1]
    bf = g['boundaryEdges'] if dim == 2 else g['boundaryVertices'].reshape(-1, 1)
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=bf)
    dm = pb.P1_DoFMap(mesh)
    assert np.array_equal(np.where(dm.dofs >= 0, dm.dofs, -1), np.where(g['dofs'] >= 0, g['dofs'], -1))
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(dim, float(g['s'])), {'target_order': float(g['target_order'])})
    scale = np.abs(g['diagonal']).max()
    vals = b.getEntries(g['IJ'])
    assert (np.abs(vals-g['entries'])/np.maximum(np.abs(g['entries']), 1e-2*scale)).max() < TOL
    I, J = (int(v) for v in g['IJ'][-1])
    assert abs(b.getEntry(I, J)-g['entries'][-1]) < TOL*max(abs(g['entries'][-1]), 1e-2*scale)
    D = b.getDiagonal()
    assert D.shape == (dm.num_dofs, dm.num_dofs)
    assert np.abs(D.data/g['diagonal']-1).max() < TOL
    x = np.arange(1., dm.num_dofs+1)
    assert np.array_equal(D*x, D.data*x)


def test_krylov_loops_on_device(golden_dir):
    """device CG / GMRES on the assembled operators: residual histories, iteration counts and solutions
    of the reference's cg_solver / gmres_solver (base/PyNucleus_base/solvers.pyx:329-660); then the same loops on the
    H2 operator"""
    import torch
    import pynucleus_b200 as pb
    g = load(golden_dir, 'solvers_disc_s0.75_r3')
    b = builder_from_golden(load(golden_dir, 'disc_s0.75_r3'))
    A = b.getDense()
    rhs = torch.as_tensor(g['b']).cuda()
    for tag, jac in (('', True), ('_noprec', False)):
        x, its, res = pb.cg(A, rhs, tol=1e-10, maxiter=200, jacobi=jac)
        assert its == int(g['cg_iterations'+tag]) and len(res) == len(g['cg_residuals'+tag])
        assert np.abs(np.array(res)/g['cg_residuals'+tag]-1).max() < 1e-6
        assert np.abs(x.cpu().numpy()-g['cg_x'+tag]).max() < 1e-10*np.abs(g['cg_x'+tag]).max()
        for left in (True, False):
            key = 'gmres_'+('left' if left else 'right')+tag
            x, its, res = pb.gmres(A, rhs, tol=1e-10, maxiter=12, restarts=20, jacobi=jac, left=left)
            assert its == int(g[key+'_iterations']) and len(res) == len(g[key+'_residuals'])
            assert np.abs(np.array(res)/g[key+'_residuals']-1).max() < 1e-6
            assert np.abs(x.cpu().numpy()-g[key+'_x']).max() < 1e-10*np.abs(g[key+'_x']).max()
    # H2 operator of the r=4 mesh: Jacobi from the near-field diagonal; solution close to the dense solve
    g4 = load(golden_dir, 'h2_disc_s0.75_r4')
    b4 = builder_from_golden(g4)
    H, A4 = b4.getH2(), b4.getDense()
    assert np.abs(H.diagonal/A4.diagonal-1).max() < 1e-3
    rhs = torch.as_tensor(g4['x']).cuda()
    xd, _, _ = pb.cg(A4, rhs, tol=1e-10, maxiter=500)
    xh, its, res = pb.cg(H, rhs, tol=1e-10, maxiter=500)
    xg, _, _ = pb.gmres(H, rhs, tol=1e-10, maxiter=30, restarts=20)
    assert float((xh-xd).abs().max()) < 1e-3*float(xd.abs().max())
    assert float((xh-xg).abs().max()) < 1e-7*float(xd.abs().max())


@pytest.mark.parametrize('name', ['h2_regional_interval_s0.25_r8', 'h2_regional_disc_s0.75_r4'])
def test_h2_regional_operator(golden_dir, name):
    """zeroExterior=False: the near field carries the surface terms around the cluster unions minus those of the domain
    boundary (assembleClusters, nonlocalAssembly_{SCALAR}.pxi:1840-1912); getEntry keeps the patch part only"""
    import scipy.sparse as sp
    g = load(golden_dir, name)
    b = builder_from_golden(g, zeroExterior=False)
    H, Pnear = b.getH2(returnNearField=True)
    assert np.array_equal(np.array([(a.id, c.id) for a, c in Pnear]), g['near_pairs'])
    N = b.dm.num_dofs
    low = sp.csr_matrix((g['Anear_data'], g['Anear_indices'], g['Anear_indptr']), shape=(N, N))
    ref_near = (low+low.T+sp.diags(g['Anear_diagonal'])).toarray()
    An = H.Anear.toarray()
    assert np.array_equal(An != 0, ref_near != 0)
    assert entry_err(An, ref_near) < TOL
    Hx = H*g['x']
    assert np.abs(Hx-g['Hx']).max() < 1e-11*np.abs(g['Hx']).max()
    A = b.getDense()
    assert np.abs(A*g['x']-g['Ax']).max() < TOL*np.abs(g['Ax']).max()
    vals = b.getEntries(g['IJ'])
    scale = np.abs(g['Anear_diagonal']).max()
    assert (np.abs(vals-g['entries'])/np.maximum(np.abs(g['entries']), 1e-2*scale)).max() < TOL


def test_cg_mg_driver_config_vs_reference(golden_dir):
    """BASELINE config 3 (runFractional: disc, s = varconst(0.75), P1, dense, solver cg-mg) at 4 refinements: the
    hierarchy of dense level operators assembled on the device, V-cycle multigrid (Jacobi 2/3, LU on the coarsest
    level) and multigrid-preconditioned CG against the reference's driver run: level diagonals 1e-12, iteration
    counts equal, residual histories 1e-6, solution 1e-8, Hs error of the driver"""
    import torch
    from scipy.special import gamma
    import pynucleus_b200 as pb
    g = load(golden_dir, 'mg_disc_varconst0.75_r4')
    kernel = pb.getFractionalKernel(2, pb.variableConstFractionalOrder(0.75))
    levels = pb.hierarchy(pb.uniform_disc(), int(g['noRef']), kernel, {'target_order': 0.5})
    assert [lvl['A'].shape[0] for lvl in levels] == list(g['level_num_dofs'])
    for k, lvl in enumerate(levels):
        assert np.abs(lvl['A'].diagonal/g['diag%d' % k]-1).max() < TOL
    mg = pb.multigrid(levels)
    b = torch.as_tensor(g['b']).cuda()
    x, its, res = mg.solve(b, tol=1e-8, maxiter=60)
    assert its == int(g['mg_iterations']) and len(res) == len(g['mg_residuals'])
    assert np.abs(np.array(res)/g['mg_residuals']-1).max() < 1e-6
    assert np.abs(x.cpu().numpy()-g['mg_x']).max() < 1e-9*np.abs(g['mg_x']).max()
    x, its, res = pb.cg(levels[-1]['A'], b, tol=float(g['tol']), maxiter=100, precond=mg.asPreconditioner())
    assert its == int(g['cgmg_iterations']) and len(res) == len(g['cgmg_residuals'])
    assert np.abs(np.array(res)/g['cgmg_residuals']-1).max() < 1e-6
    u = x.cpu().numpy()
    assert np.abs(u-g['u']).max() < 1e-8*np.abs(g['u']).max()
    s = 0.75
    C = 2.**(-2.*s)*gamma(1.)/gamma(1.+s)/gamma(1.+s)
    Hs_error = np.sqrt(abs(g['b'].dot(u)-C*np.pi/(s+1)))
    assert abs(Hs_error/float(g['Hs_error'])-1) < 1e-6
    # the right-hand side of the driver is the P1 load vector of f = 1
    assert np.abs(_p1_load_vector(levels[-1]['mesh'], levels[-1]['DoFMap'])-g['b']).max() < 1e-14


def _driver_errors(domain, s, fmt, noRef, params, solver, variable=False):
    """Hs and L2 errors of the 'constant' problem (f = 1, u = C (1-|x|^2)^s) as the reference's drivers report them
    (nl/PyNucleus_nl/discretizedProblems.py:77-110, exact values nonlocalProblems.py:741-749)"""
    from scipy.special import gamma
    import pynucleus_b200 as pb
    dim = 1 if domain == 'interval' else 2
    mesh = pb.refined(pb.simpleInterval(-1, 1) if dim == 1 else pb.uniform_disc(), noRef)
    dm = pb.P1_DoFMap(mesh)
    order = pb.constantNonSymFractionalOrder(s) if variable == 'nonsym' else (pb.variableConstFractionalOrder(s) if variable else s)
    builder = pb.nonlocalBuilder(dm, pb.getFractionalKernel(dim, order), params)
    A = builder.getH2() if fmt == 'H2' else builder.getDense()
    b = _p1_load_vector(mesh, dm)
    if solver == 'lu':
        u = pb.lu(A, b)
    elif solver == 'gmres':
        u = pb.gmres(A, b, tol=1e-12, maxiter=40, restarts=100)[0]
    else:
        u = pb.cg(A, b, tol=1e-12, maxiter=3000)[0]
    C = 2.**(-2.*s)*gamma(dim/2.)/gamma(dim/2.+s)/gamma(1.+s)
    if dim == 1:
        Hs_ex2 = C*np.sqrt(np.pi)*gamma(s+1)/gamma(s+1.5)
        L2_ex2 = C**2*np.sqrt(np.pi)*gamma(2*s+1)/gamma(2*s+1.5)
        # default rule of assembleRHS for P1 in 1D: Gauss1D(order=3) (fem/PyNucleus_fem/femCy.pyx:2640, quadrature.pyx:303-316)
        t, w = np.polynomial.legendre.leggauss(2)
        rule = (np.stack(((t+1)/2, 1-(t+1)/2)), w/2)
    else:
        Hs_ex2 = C*np.pi/(s+1)
        L2_ex2 = C**2*np.pi/(1+2*s)
        rule = (np.array([[0.5, 0.0, 0.5], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5]]), np.full(3, 1./3.))

    def u_exact(x):
        return C*np.maximum(1.-(x**2).sum(axis=-1), 0.)**s
    z = _p1_load_vector(mesh, dm, u_exact, rule=rule)
    M = _p1_mass(mesh, dm)
    return np.sqrt(abs(b.dot(u)-Hs_ex2)), np.sqrt(abs(L2_ex2-2*z.dot(u)+u.dot(M.dot(u))))


# (domain, s, format, refinements, params, solver, variable) -> cached (Hs error, L2 error) of the reference's
# tests/cache_runFractional.py--domain*--s*--problemconstant--elementP1--solver*--matrixFormat*, and the relative
# tolerances held here.  The 'interval' domain of the drivers starts from two cells (127 DoFs after the default
# refinements).  1D dense runs are pinned end to end (1e-9 / 1e-6: the L2 number also carries the quadrature of the
# exact solution); the 2D dense run carries the substituted regular triangle rules (~1e-5); H2 runs are
# held to the reference's own tolerance rTol = 3e-2 (its cluster parameters come from the driver defaults).
DRIVER_CASES = [
    (('interval', 0.25, 'dense', 7, {}, 'cg', False), (0.09611243700804001, 0.026655318974538753), (1e-9, 1e-6)),
    (('interval', 0.75, 'dense', 7, {}, 'lu', False), (0.04184296289342096, 0.0014584869810690354), (1e-9, 1e-6)),
    (('interval', 0.75, 'dense', 7, {}, 'cg', True), (0.041842962898268554, 0.0014584869817160686), (1e-9, 1e-6)),
    (('interval', 0.25, 'H2', 7, {}, 'cg', False), (0.0961124909768421, 0.026655322403497637), (3e-2, 3e-2)),
    (('interval', 0.75, 'H2', 7, {}, 'gmres', False), (0.041849732677658555, 0.001458788789368659), (3e-2, 3e-2)),
    (('disc', 0.25, 'dense', 5, {'target_order': 0.5}, 'cg', False), (0.1839933908571473, 0.057885119791182965), (1e-4, 1e-4)),
    (('disc', 0.75, 'H2', 5, {'target_order': 0.5}, 'cg', False), (0.059725648882225826, 0.0022274080583107514), (3e-2, 3e-2)),
    # --s constantNonSym(s): the reference's unsymmetric code path on a constant order (both orientations of every cell
    # pair); cached under ...--sconstantNonSym(*)--problemconstant--elementP1--solvergmres-jacobi--matrixFormatdense.  The
    # driver stops gmres-jacobi at a relative residual of 1e-6, which shows in the 7th digit of the Hs error for s = 0.75
    # (the systems here are solved to 1e-12); the operators themselves are pinned entry by entry (nonsym_*_constant*).
    (('interval', 0.25, 'dense', 7, {}, 'gmres', 'nonsym'), (0.09611243700814974, 0.0266553185536795), (1e-9, 1e-6)),
    (('interval', 0.75, 'dense', 7, {}, 'gmres', 'nonsym'), (0.04184297664965481, 0.0014584875781664202), (1e-6, 1e-5)),
    (('disc', 0.25, 'dense', 5, {'target_order': 0.5}, 'gmres', 'nonsym'), (0.18399339204392906, 0.05788512423832981), (1e-4, 1e-4)),
]


@pytest.mark.parametrize('case,ref,tol', DRIVER_CASES, ids=['-'.join(str(v) for v in c[0][:4])+'-'+c[0][5]+('-'+str(c[0][6]) if c[0][6] else '') for c in DRIVER_CASES])
def test_driver_known_answers(case, ref, tol):
    """BASELINE configs 0 / 1 / 3 at the sizes of the reference's own cached driver tests"""
    Hs, L2 = _driver_errors(*case)
    assert abs(Hs/ref[0]-1) < tol[0]
    assert abs(L2/ref[1]-1) < tol[1]


def test_mesh_without_unknowns():
    """a one-cell interval has no interior DoF: empty operator of shape (0, 0)"""
    import pynucleus_b200 as pb
    mesh = pb.simpleInterval(-1, 1)
    dm = pb.P1_DoFMap(mesh)
    assert dm.num_dofs == 0
    A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(1, 0.25), {}).getDense()
    assert A.shape == (0, 0) and A.data.shape == (0, 0)
    assert A.matvec(np.zeros(0)).shape == (0, )


@pytest.mark.parametrize('name', ['nonsym_disc_leftright_r2', 'nonsym_disc_leftright_r3', 'nonsym_interval_leftright_r5',
                                  'nonsym_disc_layers_r3', 'disc_layers_sym_r2', 'nonsym_interval_innerouter_r5',
                                  'nonsym_disc_constant0.75_r2', 'nonsym_disc_constant0.25_r3', 'nonsym_interval_constant0.75_r5'])
def test_unsymmetric_piecewise_order_vs_reference(golden_dir, name):
    """Unsymmetric piecewise constant orders s(x,y) != s(y,x) (the reference switches to
    fractionalLaplacian{1,2}D_nonsym and visits both orientations of every cell pair,
    nonlocalAssembly_{SCALAR}.pxi:1412-1428): two half-weight passes per order over the two orientations; leftRight,
    layers (3 blocks) and innerOuter orders, 1D and 2D, with and without the surface terms"""
    import pynucleus_b200 as pb
    from test_oracle_golden import order_from_fixture
    g = load(golden_dir, name)
    dim = g['vertices'].shape[1]
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'] if dim == 2 else g['boundaryVertices'])
    dm = pb.P1_DoFMap(mesh)
    s = order_from_fixture(g)
    kernel = pb.getFractionalKernel(dim, s)
    assert kernel.variable and kernel.symmetric == bool(g['symmetric'])
    params = {'target_order': 0.5} if dim == 2 else {}
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = pb.nonlocalBuilder(dm, kernel, params, zeroExterior=ze)
        assert b.orders.quad_order_diagonal == int(g['quad_order_diagonal'])
        A = b.getDense().data
        assert entry_err(A, g[key]) < TOL


@pytest.mark.parametrize('name', ['h2_interval_s0.25_r8', 'h2_disc_s0.75_r4'])
def test_h2_device_engine(golden_dir, name):
    """The library's own H2 kernels (pnb_h2_*: leaf moments, upward / far-field / downward passes, CSR near field;
    clusterMethodCy.pyx:1093-1325, 2269-2295): leaf moments against the host restatement, far field and whole product
    against the reference's H2 matvec and against the node-by-node torch recursion, bitwise reproducible"""
    import scipy.sparse as sp
    import torch
    from pynucleus_b200 import h2, _lib
    g = load(golden_dir, name)
    b = builder_from_golden(g)
    H = b.getH2()
    assert getattr(H, '_engine', None), 'the device engine must be the path that runs'
    d2c = h2.dof_to_cells(b.dm)
    for n in H.tree.get_tree_nodes():
        if n.isLeaf:
            V = h2.leaf_values(n, b.mesh, b.dm, d2c)
            assert np.abs(V-n.value).max() < 1e-13*np.abs(V).max()
    x = torch.as_tensor(g['x']).cuda()
    N = b.dm.num_dofs
    yfar = torch.empty(N, dtype=torch.float64, device='cuda')
    _lib.check(_lib.lib().pnb_h2_matvec(H._engine, x.data_ptr(), yfar.data_ptr(), 1, torch.cuda.current_stream().cuda_stream))
    low = sp.csr_matrix((g['Anear_data'], g['Anear_indices'], g['Anear_indptr']), shape=(N, N))
    ref_far = g['Hx']-(low+low.T+sp.diags(g['Anear_diagonal'])).dot(g['x'])
    assert np.abs(yfar.cpu().numpy()-ref_far).max() < 1e-11*np.abs(ref_far).max()
    ytorch = H.farfield_device(x)
    assert float((yfar-ytorch).abs().max()) < 1e-13*float(ytorch.abs().max())
    y1, y2 = H.matvec_device(x), H.matvec_device(x)
    assert torch.equal(y1, y2)
    assert np.abs(y1.cpu().numpy()-g['Hx']).max() < 1e-11*np.abs(g['Hx']).max()
    # the library-SpMV formulation gives the same product
    H.compile()
    yc = H.farfield_compiled(x)+H.Anear.matvec_device(x)
    assert float((y1-yc).abs().max()) < 1e-12*float(yc.abs().max())


@pytest.mark.parametrize('dim,s,errBnd', [(1, 0.3, 1e-4), (1, 0.7, 1e-2), (2, 0.3, 1.2e-4), (2, 0.7, 1e-2)])
def test_h2_like_the_reference_test(dim, s, errBnd):
    """tests/test_fracLapl.py:141-237 of the reference (`h2` / `testH2`), statement for statement and with its parameters
    (orders, error bounds, refinements, eta, maxLevels, the DoFMap tag) against this package's objects: near field against
    the dense operator outside the admissible blocks, the far field through the tree's own upwardPass_py /
    resetCoefficientsDown_py / clusterPair.apply / downwardPass_py, and the whole product.  The 2D mesh is the 10-gon fan
    instead of `circle(10)` (meshpy is not available offline)."""
    import pynucleus_b200 as pb
    from pynucleus_b200.h2 import H2Matrix
    if dim == 1:
        mesh, refinements, eta, maxLevels = pb.simpleInterval(-1, 1), 6, 1, None
    else:
        fan = pb.polygon_disc(10)
        mesh, refinements, eta, maxLevels = pb.meshNd(fan.vertices, fan.cells), 3, 3, 4
    for _ in range(refinements):
        mesh = mesh.refine()
    # tag=-1 (no Dirichlet vertices) for s < 0.5, the whole boundary otherwise
    dm = pb.P1_DoFMap(mesh, tag=np.zeros(mesh.num_vertices, dtype=bool) if s < 0.5 else None)
    params = {'eta': eta, 'maxLevels': maxLevels}
    builder = pb.nonlocalBuilder(dm, pb.getFractionalKernel(dim, s), params=params, zeroExterior=True)
    A_d = np.array(builder.getDense().data)
    A_h2 = builder.getH2()
    assert isinstance(A_h2, H2Matrix)
    n = A_d.shape[0]
    Afar = np.zeros((n, n))
    for level in A_h2.Pfar:
        for c in A_h2.Pfar[level]:
            Afar[np.ix_(list(c.n1.dofs.toSet()), list(c.n2.dofs.toSet()))] = A_d[np.ix_(list(c.n1.dofs.toSet()), list(c.n2.dofs.toSet()))]
    Anear = A_d-Afar
    errNear = np.absolute(Anear-A_h2.Anear.toarray()).max()
    x = np.ones((A_d.shape[0]))
    y_d = np.dot(Afar, x)
    y_h2 = np.zeros_like(y_d)
    assert len(A_h2.Pfar) > 0
    A_h2.tree.upwardPass_py(x)
    A_h2.tree.resetCoefficientsDown_py()
    for level in A_h2.Pfar:
        for clusterPair in A_h2.Pfar[level]:
            n1, n2 = clusterPair.n1, clusterPair.n2
            clusterPair.apply(n2.coefficientsUp, n1.coefficientsDown)
    A_h2.tree.downwardPass_py(y_h2)
    errFar = np.absolute(y_d-y_h2).max()
    y_d = np.dot(A_d, x)
    y_h2 = A_h2*x
    errAll = np.absolute(y_d-y_h2).max()
    # the reference asserts errNear < errBnd, errFar < errBnd, errAll < errBnd (tests/test_fracLapl.py:191-200)
    assert errNear < errBnd and errFar < errBnd and errAll < errBnd, (errNear, errFar, errAll)


def test_unsymmetric_order_larger_mesh_vs_oracle():
    """an unsymmetric leftRight order on a mesh with uniformly far units on both sides of the interface (cell-group path for
    the pairs with both orientations of an order, filtered DoF-tile passes for the touching pairs, half-weight passes across
    the interface) against the oracle's two-orientation restatement (pinned to the reference by
    test_unsymmetric_piecewise_order_matches_reference)"""
    import oracle
    import pynucleus_b200 as pb
    mesh = pb.refined(pb.polygon_disc(7), 4)
    dm = pb.P1_DoFMap(mesh)
    s = pb.leftRightFractionalOrder(0.3, 0.7, 0.55, 0.4, 0.1)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, s), {'target_order': 0.5})
    A = b.getDense().data
    assert len(b._classes['passes']) == 12
    labels = s.labels(mesh.vertices[mesh.cells].mean(axis=1))
    blabels = s.labels(mesh.vertices[mesh.boundaryFacets].mean(axis=1))
    sV = s.blockOrders()
    vals = sorted(set(sV.ravel().tolist()))
    pc = np.array([[vals.index(sV[i, j]) for j in range(2)] for i in range(2)])
    ref = 0.
    for k, sv in enumerate(vals):
        for M, orientation in ((pc == k, 0), ((pc == k).T, 1)):
            P4, B4 = np.zeros((4, 4), dtype=np.uint8), np.zeros((4, 4), dtype=np.uint8)
            P4[:2, :2] = M
            B4[:2, :2] = pc == k
            P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, sv, bfacets=mesh.boundaryFacets, target_order=0.5,
                               s_max=max(vals), s_min=min(vals), labels=labels, blabels=blabels, pair_class=P4, bpair_class=B4,
                               pair_orientation=orientation, active_class=1, max_order=40)
            ref = ref+0.5*P.dense(True)
    assert entry_err(A, ref) < TOL


def test_fused_cg_kernels(golden_dir):
    """the fused BLAS-1 kernels of the CG loop (pnb_krylov_*, csrc/pnb_krylov.cuh) against the reference's histories (above,
    through the default path of pb.cg) and against the step-by-step torch formulation: same iteration counts, histories to
    1e-9, bitwise reproducible; the 50-iteration residual refresh is exercised on the larger operator"""
    import torch
    import pynucleus_b200 as pb
    from pynucleus_b200 import _lib
    # the dot product kernel alone
    L = _lib.lib()
    rng = np.random.default_rng(5)
    a, c = rng.standard_normal(100003), rng.standard_normal(100003)
    at, ct = torch.as_tensor(a).cuda(), torch.as_tensor(c).cuda()
    work = torch.zeros(int(L.pnb_krylov_workspace_doubles()), dtype=torch.float64, device='cuda')
    out = torch.zeros(1, dtype=torch.float64, device='cuda')
    _lib.check(L.pnb_krylov_dot(0, a.shape[0], at.data_ptr(), ct.data_ptr(), work.data_ptr(), out.data_ptr(),
                                torch.cuda.current_stream().cuda_stream))
    assert abs(float(out)-a.dot(c)) < 1e-12*np.abs(a*c).sum()
    for noRef, maxiter in ((3, 200), (5, 400)):
        mesh = pb.refined(pb.uniform_disc(), noRef)
        dm = pb.P1_DoFMap(mesh)
        A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5}).getDense()
        rhs = torch.ones(dm.num_dofs, dtype=torch.float64, device='cuda')
        for jac in (True, False):
            for kw in ({}, {'use2norm': True}, {'relative': True}):
                x1, i1, r1 = pb.cg(A, rhs, tol=1e-12, maxiter=maxiter, jacobi=jac, **kw)
                x0, i0, r0 = pb.cg(A, rhs, tol=1e-12, maxiter=maxiter, jacobi=jac, fused=False, **kw)
                assert i1 == i0 and len(r1) == len(r0)
                # histories agree to rounding until the residual reaches the noise floor of the recurrences (~1e-13 of its start)
                assert (np.abs(np.array(r1)-np.array(r0)) < 1e-9*np.array(r0)+1e-12*r0[0]).all()
                assert float((x1-x0).abs().max()) < 1e-9*float(x0.abs().max())
                x2, i2, r2 = pb.cg(A, rhs, tol=1e-12, maxiter=maxiter, jacobi=jac, **kw)
                assert torch.equal(x1, x2) and r1 == r2
        if noRef == 5:
            assert i1 > 50      # the residual refresh of iteration 50 ran


def test_piecewise_order_with_a_single_block_in_the_mesh():
    """an interface outside the domain: every cell carries the same label, the passes of the other orders are empty, and
    the operator equals the constant-order one (symmetric order) / the constantNonSym one (unsymmetric order)"""
    import pynucleus_b200 as pb
    mesh = pb.refined(pb.uniform_disc(), 3)
    dm = pb.P1_DoFMap(mesh)
    params = {'target_order': 0.5}
    A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.3), params).getDense().data
    # singular quadrature orders follow s.max = 0.3 in both
    As = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, pb.leftRightFractionalOrder(0.3, 0.2, 0.25, 0.25, interface=5.)), params).getDense().data
    assert entry_err(As, A) < TOL
    An = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, pb.leftRightFractionalOrder(0.3, 0.2, 0.25, 0.1, interface=5.)), params).getDense().data
    Ac = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, pb.constantNonSymFractionalOrder(0.3)), params).getDense().data
    assert entry_err(An, Ac) < TOL
    assert 1e-9 < entry_err(Ac, A) < 1e-5      # the two orientations of the singular rules differ by their quadrature error


def _varorder_from_fixture(pb, g):
    kind = str(g['kind'])
    if kind == 'smoothedLeftRight':
        return pb.smoothedLeftRightFractionalOrder(float(g['sl']), float(g['sr']), r=float(g['r']), interface=float(g['interface']))
    assert kind == 'linearLeftRight'
    return pb.linearLeftRightFractionalOrder(float(g['sl']), float(g['sr']), r=float(g['r']), interface=float(g['interface']))


@pytest.mark.parametrize('name', ['varorder_interval_smoothed_r5', 'varorder_interval_smoothed_r6', 'varorder_interval_linear_r5',
                                  'varorder_disc_smoothed_r2', 'varorder_disc_smoothed_r3', 'varorder_p2_interval_smoothed_r4',
                                  'varorder_p2_disc_smoothed_r1', 'varorder_p0_disc_smoothed_r2'])
def test_order_varying_inside_cells_vs_reference(golden_dir, name):
    """Orders that vary inside a cell, s(x,y) = sFun(x) (SURVEY 8 a12 updateAndEvalFractional, a13 variable scaling, a14
    singleVariableUnsymmetricFractionalOrder; the driver's --s twoDomainNonSym): the reference's unsymmetric local matrices
    over both orientations, order / scaling / kernel per quadrature node, a singular rule per pair singularity
    (csrc/pnb_varorder.cuh) against operators assembled by the reference itself (make_golden_varorder.py)"""
    import pynucleus_b200 as pb
    g = load(golden_dir, name)
    dim = g['vertices'].shape[1]
    element = str(g['element']) if 'element' in g.files else 'P1'
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'] if dim == 2 else g['boundaryVertices'])
    dm = {'P0': pb.P0_DoFMap, 'P1': pb.P1_DoFMap, 'P2': pb.P2_DoFMap}[element](mesh)
    # the unknowns are numbered like the reference's (the numbering of the boundary dofs, all negative, does not matter)
    assert dm.num_dofs == int(g['num_dofs']) and np.array_equal(np.where(dm.dofs >= 0, dm.dofs, -1), np.where(g['dofs'] >= 0, g['dofs'], -1))
    kernel = pb.getFractionalKernel(dim, _varorder_from_fixture(pb, g))
    assert kernel.variable and not kernel.symmetric and not kernel.piecewise
    params = {'target_order': 0.5} if dim == 2 else {}
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = pb.nonlocalBuilder(dm, kernel, params, zeroExterior=ze)
        assert b.orders.quad_order_diagonal == int(g['quad_order_diagonal'])
        assert b.orders.bquad_order_diagonal == int(g['bquad_order_diagonal'])
        A = b.getDense().data
        assert entry_err(A, g[key]) < TOL
        # bitwise reproducible (one writer per entry, fixed summation order)
        assert np.array_equal(A, b.getDense().data)
    with pytest.raises(NotImplementedError):
        b.getH2()


def test_order_varying_inside_cells_vs_oracle():
    """the same path on meshes / orders without a reference fixture, against the numpy restatement oracle/varorder.py
    (itself pinned to the reference's fixtures in test_oracle_golden.py): a linear ramp across the disc, an interface off
    the mesh lines, a small max_regular_order so that the table retry runs"""
    import pynucleus_b200 as pb
    from oracle import varorder
    mesh = pb.refined(pb.uniform_disc(), 2)
    dm = pb.P1_DoFMap(mesh)
    order = pb.linearLeftRightFractionalOrder(0.35, 0.65, r=0.4, interface=0.13)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, order), {'target_order': 0.5, 'max_regular_order': 3})
    A = b.getDense().data
    Aref = varorder.dense(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, varorder.linearStep(0.35, 0.65, 0.4, 0.13),
                          mesh.boundaryFacets, hmin=mesh.hmin, diam=mesh.diam)
    assert entry_err(A, Aref) < TOL
    mesh = pb.refined(pb.simpleInterval(-1, 1), 6)
    dm = pb.P1_DoFMap(mesh)
    order = pb.smoothedLeftRightFractionalOrder(0.2, 0.8, r=0.33, interface=-0.21)
    A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(1, order), {}).getDense().data
    Aref = varorder.dense(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, varorder.smoothStep(0.2, 0.8, 0.33, -0.21),
                          mesh.boundaryFacets, hmin=mesh.hmin, diam=mesh.diam)
    assert entry_err(A, Aref) < TOL


# cached tests/cache_runFractional.py--domain*--stwoDomainNonSym(0.25,0.75)--problemknownSolution--elementP1--solver*--matrixFormatdense
# of the reference: (L2 error, L2 error interpolated, Linf error interpolated) and the relative tolerance held here
VARORDER_DRIVER_CASES = [
    ('interval', 7, (0.0020560901451394443, 0.001265060713568335, 0.003599161364716205), 1e-7),
    # the disc file is in the reference's cache but no longer in its test list (tests/test_drivers_intFracLapl.py runs the
    # order on the interval and the square only), i.e. it may predate the current code: held as a sanity bound (measured
    # deviation 2.0e-3 in the interpolated L2 error); the 2D operator itself is pinned by rows of the reference's own
    # operator at 721 and 2 977 DoFs (test_order_varying_inside_cells_rows_vs_reference)
    ('disc', 5, (0.005965596537366911, 0.003240255585173516, 0.011192410553490267), 5e-2),
]


@pytest.mark.parametrize('domain,noRef,ref,tol', VARORDER_DRIVER_CASES, ids=[c[0] for c in VARORDER_DRIVER_CASES])
def test_driver_known_answer_two_domain_nonsym(domain, noRef, ref, tol):
    """the reference's cached driver run `--s twoDomainNonSym(0.25,0.75) --problem knownSolution` (u = (1-|x|^2)^0.7, forcing
    from the order at the point, nonlocalProblems.py:710-726, 783-799): operator from pnb_dense_assemble_varorder, direct
    solve, the errors as the driver reports them (discretizedProblems.py:77-110)"""
    from scipy.special import hyp2f1, gamma as Gamma
    import pynucleus_b200 as pb
    from pynucleus_b200 import quadrature
    dim = 1 if domain == 'interval' else 2
    mesh = pb.refined(pb.simpleInterval(-1, 1) if dim == 1 else pb.uniform_disc(), noRef)
    dm = pb.P1_DoFMap(mesh)
    order = pb.smoothedLeftRightFractionalOrder(0.25, 0.75)
    A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(dim, order), {'target_order': 0.5} if dim == 2 else {}).getDense()
    beta = 0.7

    def forcing(pts):
        s = order.evaluate(pts)
        r2 = (pts**2).sum(axis=-1)
        if dim == 1:
            return 2**(2*s)*Gamma(s+0.5)*Gamma(beta+1.)/np.sqrt(np.pi)/Gamma(beta+1.-s)*hyp2f1(s+0.5, -beta+s, 0.5, r2)
        return 2**(2*s)*Gamma(s+1.0)*Gamma(beta+1.)/Gamma(beta+1.-s)*hyp2f1(s+1.0, -beta+s, 1.0, r2)

    def u_exact(pts):
        return np.maximum(1.-(pts**2).sum(axis=-1), 0.)**beta
    # load vector with simplexXiaoGimbutas(3, dim) (discretizedProblems.py:561), the error integrals with the default rule
    b = _p1_load_vector(mesh, dm, forcing, rule=quadrature.regular(3, dim))
    u = pb.lu(A, b)
    xs = np.zeros((dm.num_dofs, dim))
    for k in range(dim+1):
        m = dm.dofs[:, k] >= 0
        xs[dm.dofs[m, k]] = mesh.vertices[mesh.cells[m, k]]
    e = u-u_exact(xs)
    M = _p1_mass(mesh, dm)
    if dim == 1:
        L2_ex2 = np.sqrt(np.pi)*Gamma(1+2*beta)/Gamma(1.5+2*beta)
        t, w = np.polynomial.legendre.leggauss(2)
        rule = (np.stack(((t+1)/2, 1-(t+1)/2)), w/2)
    else:
        L2_ex2 = np.pi/(1+2*beta)
        rule = (np.array([[0.5, 0.0, 0.5], [0.5, 0.5, 0.0], [0.0, 0.5, 0.5]]), np.full(3, 1./3.))
    z = _p1_load_vector(mesh, dm, u_exact, rule=rule)
    L2 = np.sqrt(abs(L2_ex2-2*z.dot(u)+u.dot(M.dot(u))))
    L2i = np.sqrt(e.dot(M.dot(e)))
    Linf = np.abs(e).max()
    print('\nknown answer', domain, dm.num_dofs, 'L2 %.12g L2i %.12g Linf %.12g' % (L2, L2i, Linf), 'cached', ref)
    assert abs(L2i/ref[1]-1) < tol and abs(Linf/ref[2]-1) < tol
    assert abs(L2/ref[0]-1) < max(tol, 1e-6)


TEMPERED_CASES = ['tempered_interval_s0.75_l2_r5', 'tempered_interval_s0.25_l0.5_r6', 'tempered_disc_s0.75_l2_r2',
                  'tempered_disc_s0.25_l1_r3', 'tempered_p2_interval_s0.75_l1.5_r4', 'tempered_p2_disc_s0.75_l1.5_r1']


@pytest.mark.parametrize('name', TEMPERED_CASES)
def test_tempered_kernel_vs_reference(golden_dir, name):
    """tempered fractional kernels C |x-y|^(-d-2s) exp(-lambda |x-y|) (SURVEY 8 a12 temperedFracKernelInfinite*, a13 the
    tempered scaling constant) on the row-owner kernel (P1 and P2), against operators assembled by the reference itself
    (make_golden_tempered.py); the surface terms are the untempered power law, as in the reference"""
    import pynucleus_b200 as pb
    g = load(golden_dir, name)
    dim = g['vertices'].shape[1]
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'] if dim == 2 else g['boundaryVertices'])
    dm = pb.P2_DoFMap(mesh) if str(g['element']) == 'P2' else pb.P1_DoFMap(mesh)
    assert dm.num_dofs == int(g['num_dofs'])
    kernel = pb.getFractionalKernel(dim, float(g['s']), tempered=float(g['tempered']))
    assert abs(kernel.scalingValue/float(g['scaling'])-1) < 1e-14
    params = {'target_order': 0.5} if dim == 2 else {}
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = pb.nonlocalBuilder(dm, kernel, params, zeroExterior=ze)
        assert b.orders.quad_order_diagonal == int(g['quad_order_diagonal'])
        A = b.getDense().data
        assert entry_err(A, g[key]) < TOL
        # every row has its own warp and summation order: symmetric to rounding, bitwise reproducible
        assert np.abs(A-A.T).max() < 1e-13*np.abs(A).max()
        assert np.array_equal(A, b.getDense().data)
    with pytest.raises(NotImplementedError):
        b.getH2()
    with pytest.raises(NotImplementedError):
        b.getEntry(0, 0)


def test_tempered_kernel_larger_mesh_vs_oracle():
    """tempered kernel on a mesh without a fixture against the C oracle (pinned to the fixtures in test_oracle_golden.py);
    lambda -> 0 recovers the production path's operator up to the scaling constant"""
    import pynucleus_b200 as pb
    import oracle
    mesh = pb.refined(pb.uniform_disc(), 4)
    dm = pb.P1_DoFMap(mesh)
    s, lam = 0.6, 3.0
    A = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, s, tempered=lam), {'target_order': 0.5}).getDense().data
    P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, s, bfacets=mesh.boundaryFacets, target_order=0.5,
                       tempered=lam)
    assert entry_err(A, P.dense(True)) < TOL
    k0, kt = pb.getFractionalKernel(2, s), pb.getFractionalKernel(2, s, tempered=1e-300)
    A0 = pb.nonlocalBuilder(dm, k0, {'target_order': 0.5}).getDense().data
    At = pb.nonlocalBuilder(dm, kt, {'target_order': 0.5}).getDense().data
    assert entry_err(At*(k0.scalingValue/kt.scalingValue), A0) < TOL


@pytest.mark.parametrize('name', ['varorder_disc_smoothed_r4_rows', 'varorder_disc_smoothed_r5_rows'])
def test_order_varying_inside_cells_rows_vs_reference(golden_dir, name):
    """the 2D operator of an order that varies inside the cells at 721 and 2 977 DoFs against sampled rows, the diagonal and
    the products A x, A^T x of the reference's own operator (make_golden_varorder_rows.py)"""
    import pynucleus_b200 as pb
    g = load(golden_dir, name)
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'])
    dm = pb.P1_DoFMap(mesh)
    assert dm.num_dofs == int(g['num_dofs'])
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, _varorder_from_fixture(pb, g)), {'target_order': 0.5})
    assert b.orders.quad_order_diagonal == int(g['quad_order_diagonal'])
    assert b.orders.bquad_order_diagonal == int(g['bquad_order_diagonal'])
    A = b.getDense().data
    rows = g['rows']
    d = np.sqrt(np.abs(g['diag']))
    scale = np.maximum(np.abs(g['A_rows']), 1e-2*np.outer(d[rows], d))
    assert (np.abs(A[rows]-g['A_rows'])/scale).max() < TOL
    assert np.abs(np.diag(A)/g['diag']-1).max() < TOL
    x = g['x']
    assert np.abs(A.dot(x)-g['Ax']).max() < TOL*np.abs(g['Ax']).max()
    assert np.abs(A.T.dot(x)-g['ATx']).max() < TOL*np.abs(g['ATx']).max()


SMOOTH_CASES = ['gaussian_interval_v0.1_r5', 'gaussian_interval_v0.02_r6', 'exponential_interval_a8_r5', 'exponential_interval_a2.5_r6',
                'gaussian_disc_v0.1_r2', 'gaussian_disc_v0.3_r3', 'gaussian_p2_interval_v0.1_r4']


@pytest.mark.parametrize('name', SMOOTH_CASES)
def test_gaussian_and_exponential_kernels_vs_reference(golden_dir, name):
    """Gaussian and exponential kernels on the full space (SURVEY 8 a12 gaussianKernel* / exponentialKernel* with their
    boundary forms, a13 their constantIntegrableScaling; the kernels of the reference's driver tests `--kernelType gaussian /
    exponential --interaction fullSpace`): pnb_dense_assemble_element_smooth against operators assembled by the reference
    itself (make_golden_smooth.py), with and without the surface terms"""
    import pynucleus_b200 as pb
    g = load(golden_dir, name)
    dim = g['vertices'].shape[1]
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'] if dim == 2 else g['boundaryVertices'])
    dm = pb.P2_DoFMap(mesh) if str(g['element']) == 'P2' else pb.P1_DoFMap(mesh)
    assert dm.num_dofs == int(g['num_dofs'])
    kw = {'variance': float(g['variance'])} if str(g['kernelType']) == 'gaussian' else {'exponentialRate': float(g['exponentialRate'])}
    kernel = pb.getIntegrableKernel(dim, str(g['kernelType']), np.inf, interaction='fullSpace', **kw)
    params = {'target_order': 0.5} if dim == 2 else {}
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = pb.nonlocalBuilder(dm, kernel, params, zeroExterior=ze)
        A = b.getDense().data
        assert entry_err(A, g[key]) < TOL
        assert np.array_equal(A, b.getDense().data)
    with pytest.raises(NotImplementedError):
        b.getH2()


def test_gaussian_kernel_larger_mesh_vs_oracle(golden_dir):
    """Gaussian kernel on a finer disc (705 DoFs) than the fixtures, against the C oracle (pinned to the fixtures)"""
    import pynucleus_b200 as pb
    import oracle
    from math import pi
    mesh = pb.refined(pb.uniform_disc(), 4)
    dm = pb.P1_DoFMap(mesh)
    var = 0.05
    kernel = pb.getIntegrableKernel(2, 'gaussian', np.inf, variance=var)
    A = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5}).getDense().data
    C, a = kernel.scalingValue, 0.5/var**2
    assert abs(C-1./(2*pi*var)/2) < 1e-15
    P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, -1., bfacets=mesh.boundaryFacets, target_order=0.5,
                       smooth=(C, 2, a, C/a, 4, a))
    assert entry_err(A, P.dense(True)) < TOL


@pytest.mark.parametrize('kt', ['gaussian', 'exponential'])
def test_driver_known_answers_gaussian_and_exponential(kt):
    """the reference's cached driver runs `runNonlocal.py --domain interval --kernelType gaussian --gaussianVariance 0.1` /
    `--kernelType exponential --exponentialRate 8` with `--interaction fullSpace --horizon inf` (tests/test_drivers_intFracLapl.py:42-43;
    511 unknowns): operator from pnb_dense_assemble_element_smooth, the errors as the driver reports them"""
    import pynucleus_b200 as pb
    from test_oracle_golden import SMOOTH_DRIVER_CASES, smooth_driver_setup, smooth_driver_errors
    mesh = pb.refined(pb.simpleInterval(-1, 1), 9)
    dm = pb.P1_DoFMap(mesh)
    kw, _, f, u_ex = smooth_driver_setup(kt)
    kernel = pb.getIntegrableKernel(1, kt, np.inf, interaction='fullSpace', **kw)
    A = pb.nonlocalBuilder(dm, kernel, {}).getDense().data
    L2i, Linf = smooth_driver_errors(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, A, f, u_ex)
    assert abs(L2i/SMOOTH_DRIVER_CASES[kt][0]-1) < 1e-6 and abs(Linf/SMOOTH_DRIVER_CASES[kt][1]-1) < 1e-12


@pytest.mark.parametrize('name', ['varorder_fe_interval_r5', 'varorder_fe_disc_r2', 'varorder_fe_disc_r3'])
def test_order_given_by_a_fe_function_vs_reference(golden_dir, name):
    """feFractionalOrder (the order is a P1 finite element function, evaluated per quadrature node from the
    barycentric coordinates of the node's cell; every cell has a pair singularity of its own, hence as many sets of singular
    tables as cells) against operators assembled by the reference itself"""
    import pynucleus_b200 as pb
    from test_oracle_golden import fe_order_vertex_values
    g = load(golden_dir, name)
    dim = g['vertices'].shape[1]
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'] if dim == 2 else g['boundaryVertices'])
    dm = pb.P1_DoFMap(mesh)
    # the order lives on a map without boundary dofs (NO_BOUNDARY in the reference): every vertex carries a value
    dms = pb.P1_DoFMap(mesh, tag=np.zeros(mesh.num_vertices, dtype=bool))
    vs = fe_order_vertex_values(g)
    u = np.zeros(dms.num_dofs)
    for k in range(dim+1):
        u[dms.dofs[:, k]] = vs[mesh.cells[:, k]]
    order = pb.feFractionalOrder(dms, u, float(g['smin']), float(g['smax']))
    assert np.array_equal(order.vertexValues(mesh), vs)
    kernel = pb.getFractionalKernel(dim, order)
    assert kernel.variable and not kernel.symmetric and not kernel.piecewise
    params = {'target_order': 0.5} if dim == 2 else {}
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        b = pb.nonlocalBuilder(dm, kernel, params, zeroExterior=ze)
        assert b.orders.quad_order_diagonal == int(g['quad_order_diagonal'])
        assert b.orders.bquad_order_diagonal == int(g['bquad_order_diagonal'])
        A = b.getDense().data
        assert entry_err(A, g[key]) < TOL


@pytest.mark.parametrize('name', ['tempered_disc_s0.75_l2_r5_rows', 'gaussian_disc_v0.05_r5_rows'])
def test_smooth_factor_kernels_rows_vs_reference_at_2977_dofs(golden_dir, name):
    """a tempered and a Gaussian kernel at 2 977 DoFs against sampled rows, the diagonal and the product A x of the reference's
    own operator (make_golden_smooth_rows.py)"""
    import pynucleus_b200 as pb
    g = load(golden_dir, name)
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'])
    dm = pb.P1_DoFMap(mesh)
    assert dm.num_dofs == int(g['num_dofs'])
    if name.startswith('tempered'):
        kernel = pb.getFractionalKernel(2, float(g['s']), tempered=float(g['tempered']))
    else:
        kernel = pb.getIntegrableKernel(2, 'gaussian', np.inf, variance=float(g['variance']))
    assert abs(kernel.scalingValue/float(g['scaling'])-1) < 1e-14
    b = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5})
    assert b.orders.quad_order_diagonal == int(g['quad_order_diagonal'])
    assert b.orders.bquad_order_diagonal == int(g['bquad_order_diagonal'])
    A = b.getDense().data
    rows = g['rows']
    d = np.sqrt(np.abs(g['diag']))
    scale = np.maximum(np.abs(g['A_rows']), 1e-2*np.outer(d[rows], d))
    assert (np.abs(A[rows]-g['A_rows'])/scale).max() < TOL
    assert np.abs(np.diag(A)/g['diag']-1).max() < TOL
    assert np.abs(A.dot(g['x'])-g['Ax']).max() < TOL*np.abs(g['Ax']).max()
