"""ORACLE INFRASTRUCTURE: golden data of the REFERENCE ITSELF (stub-built copy in oracle/_ref) at the benchmarked
size, where the full matrix (3.25 GB at disc20k) cannot be committed.

    python oracle/refbuild/make_golden_big.py disc20k 32 64        # ~5 min on 8 cores, <= 8 x 3.3 GB of RAM

The reference's getDense is run once per rank slice of its own cell partition (fake communicator of `nslices` ranks,
nonlocalAssembly_{SCALAR}.pxi:1280-1285; the Allreduce at :1449-1450 is replaced by summing the linear functionals
below in the parent).  Kept per workload: `nrows` sampled rows, the diagonal, and A X for four seeded vectors
(ones, linspace, two Gaussian) -- all linear in A, so the sum over the slices is exact up to rounding.
`--check` compares the sliced sum against a single-rank run (small workloads only).
"""
import os
import sys
import time

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
sys.path.insert(0, ROOT)
OUT = os.path.join(ROOT, 'tests', 'golden')
S_ORDER = 0.75
TARGET_ORDER = 0.5


def probes(N, nrows):
    """sampled rows and probe vectors: the same seeded choice bench.py and the tests make"""
    import bench
    return bench.golden_probes(N, nrows)


def worker(args):
    workload, rank, size, nrows = args
    sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    from mpi4py import MPI
    from PyNucleus_fem.mesh import mesh2d
    from PyNucleus_fem.DoFMaps import P1_DoFMap
    from PyNucleus_nl.kernels import getFractionalKernel
    from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder
    from PyNucleus_nl.fractionalOrders import constFractionalOrder
    import bench
    mesh0, dm0 = bench.make_mesh(workload)
    mesh = mesh2d(mesh0.vertices.copy(), mesh0.cells.copy())
    dm = P1_DoFMap(mesh)
    d1, d0 = np.array(dm.dofs), np.asarray(dm0.dofs)
    assert np.array_equal(d1 < 0, d0 < 0) and np.array_equal(d1[d1 >= 0], d0[d0 >= 0])      # boundary DoFs are numbered differently
    kernel = getFractionalKernel(2, constFractionalOrder(S_ORDER), np.inf)
    comm = MPI.fakeComm(rank, size) if size > 1 else None
    b = nonlocalBuilder(dm, kernel, {'target_order': TARGET_ORDER}, comm=comm)
    t = time.time()
    A = np.array(b.getDense().data, copy=False)
    t = time.time()-t
    rows, X = probes(dm.num_dofs, nrows)
    return A[rows].copy(), np.diag(A).copy(), A.dot(X), t


def main():
    import multiprocessing as mp
    workload = sys.argv[1]
    size = int(sys.argv[2])
    nrows = int(sys.argv[3])
    check = '--check' in sys.argv
    ctx = mp.get_context('spawn')
    t0 = time.time()
    with ctx.Pool(min(size, len(os.sched_getaffinity(0))), maxtasksperchild=1) as pool:
        # heaviest slices (lowest cell ranges) first
        res = pool.map(worker, [(workload, r, size, nrows) for r in range(size)], chunksize=1)
    A_rows = sum(r[0] for r in res)
    diag = sum(r[1] for r in res)
    AX = sum(r[2] for r in res)
    print('slices done in %.1f s (cpu %.1f s)' % (time.time()-t0, sum(r[3] for r in res)))
    if check:
        ref = worker((workload, 0, 1, nrows))
        for a, b_, name in ((A_rows, ref[0], 'rows'), (diag, ref[1], 'diag'), (AX, ref[2], 'AX')):
            print('check', name, np.abs(a-b_).max()/np.abs(b_).max())
    import bench
    mesh, dm = bench.make_mesh(workload)
    rows, X = probes(dm.num_dofs, nrows)
    np.savez_compressed(os.path.join(OUT, workload+'_rows.npz'), workload=workload, s=S_ORDER, target_order=TARGET_ORDER,
                        num_dofs=dm.num_dofs, num_cells=mesh.num_cells, rows=rows, A_rows=A_rows, diagonal=diag, AX=AX,
                        nslices=size, seed=20161)
    print('wrote', workload+'_rows.npz', A_rows.shape)


if __name__ == '__main__':
    main()
