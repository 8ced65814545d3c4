"""Host-side logic of the row-block (multi-GPU) path on CPU with gloo, world_size 2: the row partition and the
exchange of the per-cell diagonal blocks (sum over ranks of arrays with disjoint supports)."""
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pynucleus_b200.assembly import row_partition, exchange_cell_blocks
    N, nc = 721, 1536
    blocks = row_partition(N, world, 64)
    a, b = blocks[rank]
    # every rank "owns" the cells whose index falls into its share; supports are disjoint
    mine = np.zeros(nc*6)
    own = np.arange(nc) % world == rank
    mine.reshape(nc, 6)[own] = np.arange(nc*6).reshape(nc, 6)[own]+1.

    def fill(buf):
        buf.copy_(torch.from_numpy(mine))
    D = exchange_cell_blocks(nc*6, torch.device('cpu'), None, fill)
    ok = np.array_equal(D.numpy(), np.arange(nc*6)+1.)
    q.put((rank, (a, b), ok))
    dist.destroy_process_group()


def test_partition_and_exchange_gloo():
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 29500+os.getpid() % 1000
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(r[2] for r in res)
    (a0, b0), (a1, b1) = res[0][1], res[1][1]
    assert a0 == 0 and b0 == a1 and b1 == 721 and b0 % 64 == 0


def test_row_partition_properties():
    from pynucleus_b200.assembly import row_partition
    for N in (1, 37, 64, 65, 721, 20161, 105_000):
        for world in (1, 2, 3, 4, 8):
            blocks = row_partition(N, world, 64)
            assert len(blocks) == world and blocks[0][0] == 0 and blocks[-1][1] == N
            for (a, b), (c, d) in zip(blocks[:-1], blocks[1:]):
                assert b == c and a <= b
            assert all(a % 64 == 0 for a, _ in blocks)


class _HostRows:
    """stands in for the device row block (the CUDA matvec cannot run here): same interface, torch CPU"""

    def __init__(self, rows):
        self.device_data = rows

    def matvec_device(self, x, y):
        torch.mv(self.device_data, x, out=y)
        return y


def _cg_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    from pynucleus_b200.solvers import DistributedDenseOperator, cg
    N = 203
    rng = np.random.default_rng(5)
    B = rng.standard_normal((N, N))
    A = torch.from_numpy(B.dot(B.T)+N*np.eye(N))
    # row SETS like the 2D distributed assembly hands out (rows of the dofs of a range of cell groups): interleaved runs
    all_rows = [np.nonzero((np.arange(N)//7) % world == r)[0] for r in range(world)]
    mine = torch.from_numpy(all_rows[rank])
    op = DistributedDenseOperator(_HostRows(A[mine].contiguous()), all_rows, rank, N)
    x = torch.from_numpy(rng.standard_normal(N))
    y = op.matvec_device(x)
    ok_mv = torch.allclose(y, A.mv(x), rtol=1e-13, atol=1e-11)
    ok_diag = torch.equal(op.diagonal_device(), torch.diagonal(A))
    rhs = torch.from_numpy(rng.standard_normal(N))
    u, its, res = cg(op, rhs, tol=1e-12, maxiter=500)
    ok_cg = float((A.mv(u)-rhs).abs().max()) < 1e-9
    from pynucleus_b200.solvers import gmres
    for left in (True, False):
        v, gits, gres = gmres(op, rhs, tol=1e-11, maxiter=20, restarts=30, left=left)
        ok_cg = ok_cg and float((A.mv(v)-rhs).abs().max()) < 1e-8 and float((u-v).abs().max()) < 1e-9
    q.put((rank, bool(ok_mv), bool(ok_diag), bool(ok_cg), its))
    dist.destroy_process_group()


def test_distributed_matvec_and_cg_gloo():
    """row-block operator: local rows + all-gather of the product, and the device CG loop on top of it"""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 30500+os.getpid() % 1000
    procs = [ctx.Process(target=_cg_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] and r[2] and r[3] for r in res), res
    assert res[0][4] == res[1][4]


class _HostOperator:
    """dense operator on the CPU with the interface the Krylov loops use (the CUDA matvec cannot run here)"""

    def __init__(self, A):
        self.device_data = torch.from_numpy(np.ascontiguousarray(A))

    def matvec_device(self, x, y=None):
        return torch.mv(self.device_data, x)


def test_krylov_loops_follow_the_reference(golden_dir):
    """cg / gmres (solvers.py) against the residual histories, iteration counts and solutions of the reference's
    cg_solver / gmres_solver (base/PyNucleus_base/solvers.pyx:329-660) on the reference's own operator"""
    from pynucleus_b200.solvers import cg, gmres
    g = np.load(os.path.join(golden_dir, 'solvers_disc_s0.75_r3.npz'))
    A = _HostOperator(np.load(os.path.join(golden_dir, 'disc_s0.75_r3.npz'))['A'])
    b = torch.from_numpy(g['b'])
    for tag, jac in (('', True), ('_noprec', False)):
        x, its, res = cg(A, b, tol=1e-10, maxiter=200, jacobi=jac)
        assert its == int(g['cg_iterations'+tag]) and len(res) == len(g['cg_residuals'+tag])
        assert np.abs(np.array(res)/g['cg_residuals'+tag]-1).max() < 1e-6
        assert np.abs(x.numpy()-g['cg_x'+tag]).max() < 1e-10*np.abs(g['cg_x'+tag]).max()
        for left in (True, False):
            key = 'gmres_'+('left' if left else 'right')+tag
            x, its, res = gmres(A, b, tol=1e-10, maxiter=12, restarts=20, jacobi=jac, left=left)
            assert its == int(g[key+'_iterations']) and len(res) == len(g[key+'_residuals'])
            assert np.abs(np.array(res)/g[key+'_residuals']-1).max() < 1e-6
            assert np.abs(x.numpy()-g[key+'_x']).max() < 1e-10*np.abs(g[key+'_x']).max()


def test_multigrid_follows_the_reference(golden_dir):
    """multigrid / cg-mg host logic (pynucleus_b200/multigrid.py) on CPU stand-ins of the level operators (assembled by
    the oracle) against the reference's driver run disc / varconst(0.75) / P1 / dense / cg-mg at 3 refinements:
    level diagonals, V-cycle residual history, preconditioned CG history and solution"""
    sys.path.insert(0, ROOT)
    import oracle
    import pynucleus_b200 as pb
    from pynucleus_b200.solvers import cg
    g = np.load(os.path.join(golden_dir, 'mg_disc_varconst0.75_r3.npz'))
    mesh = pb.uniform_disc()
    levels = []
    for k in range(int(g['num_levels'])):
        if k > 0:
            mesh = mesh.refine()
        dm = pb.P1_DoFMap(mesh)
        P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, 0.75, bfacets=mesh.boundaryFacets, target_order=0.5)
        A = P.dense(True)
        assert np.abs(np.diag(A)/g['diag%d' % k]-1).max() < 1e-12
        lvl = {'A': _HostOperator(A), 'DoFMap': dm}
        lvl['A'].shape = A.shape
        if k > 0:
            lvl['R'], lvl['P'] = pb.buildRestrictionProlongation(levels[-1]['DoFMap'], dm)
        levels.append(lvl)
    mg = pb.multigrid(levels)
    b = torch.from_numpy(g['b'])
    x, its, res = mg.solve(b, tol=1e-8, maxiter=60)
    assert its == int(g['mg_iterations']) and len(res) == len(g['mg_residuals'])
    assert np.abs(np.array(res)/g['mg_residuals']-1).max() < 1e-6
    assert np.abs(x.numpy()-g['mg_x']).max() < 1e-9*np.abs(g['mg_x']).max()
    x, its, res = cg(levels[-1]['A'], b, tol=1e-8, maxiter=100, precond=mg.asPreconditioner())
    assert its == int(g['cg2_iterations']) == int(g['cgmg_iterations'])
    assert np.abs(np.array(res)/g['cg2_residuals']-1).max() < 1e-6
    assert np.abs(x.numpy()-g['u']).max() < 1e-8*np.abs(g['u']).max()


def _row_parts_worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    import pynucleus_b200 as pb
    from pynucleus_b200.assembly import element_row_parts
    from pynucleus_b200.solvers import DistributedDenseOperator, gmres
    # the rows of the row-owner kernels (here: a P2 map) as every rank lists them for all ranks, without a device
    dm = pb.P2_DoFMap(pb.refined(pb.uniform_disc(), 2))
    N = dm.num_dofs
    all_rows = element_row_parts(dm, world)
    rng = np.random.default_rng(9)
    A = torch.from_numpy(rng.standard_normal((N, N))+N*np.eye(N))       # unsymmetric, like the operators of these kernels
    mine = torch.from_numpy(all_rows[rank].astype(np.int64))
    op = DistributedDenseOperator(_HostRows(A[mine].contiguous()), all_rows, rank, N)
    x = torch.from_numpy(rng.standard_normal(N))
    ok_mv = torch.allclose(op.matvec_device(x), A.mv(x), rtol=1e-13, atol=1e-11)
    rhs = torch.from_numpy(rng.standard_normal(N))
    v, its, res = gmres(op, rhs, tol=1e-11, maxiter=30, restarts=20)
    ok_solve = float((A.mv(v)-rhs).abs().max()) < 1e-8
    q.put((rank, [r.tolist() for r in all_rows], bool(ok_mv), bool(ok_solve)))
    dist.destroy_process_group()


def test_row_parts_of_the_row_owner_kernels_gloo():
    """N > 1 for the row-owner kernels on the host: both ranks list the same rows for all parts (pnb_element_rows_host), the parts
    partition the rows with long and short rows shared out evenly, and the distributed operator over these interleaved row sets
    multiplies and solves (GMRES: the operators of these kernels need not be symmetric) like the full matrix"""
    world = 2
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = 31500+os.getpid() % 1000
    procs = [ctx.Process(target=_row_parts_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=180) for _ in range(world))
    for p in procs:
        p.join(timeout=60)
    assert all(r[2] and r[3] for r in res), [r[2:] for r in res]
    assert res[0][1] == res[1][1]
    rows0, rows1 = (np.array(r) for r in res[0][1])
    N = rows0.shape[0]+rows1.shape[0]
    assert np.array_equal(np.sort(np.concatenate((rows0, rows1))), np.arange(N)) and abs(rows0.shape[0]-rows1.shape[0]) <= 1
    assert (np.diff(rows0) > 0).all() and (np.diff(rows1) > 0).all()


def test_element_row_parts_properties():
    import pynucleus_b200 as pb
    from pynucleus_b200.assembly import element_row_parts
    mesh = pb.refined(pb.uniform_disc(), 2)
    for dm in (pb.P1_DoFMap(mesh), pb.P2_DoFMap(mesh), pb.P0_DoFMap(mesh)):
        cells_around = np.bincount(dm.dofs[dm.dofs >= 0].ravel(), minlength=dm.num_dofs)
        for nparts in (1, 2, 3, 8, dm.num_dofs+5):
            parts = element_row_parts(dm, nparts)
            assert len(parts) == nparts
            assert np.array_equal(np.sort(np.concatenate(parts)), np.arange(dm.num_dofs))
            sizes = [p.shape[0] for p in parts]
            assert max(sizes)-min(sizes) <= 1
            if nparts in (2, 3):
                # rows are dealt in the order of their length: the work (cells around the dofs) is shared out evenly
                work = [cells_around[p].sum() for p in parts]
                assert max(work)-min(work) <= cells_around.max()
