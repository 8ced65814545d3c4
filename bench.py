#!/usr/bin/env python
"""Benchmark of the nonlocal dense assembly hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repository (CUDA)
    python bench.py --impl reference --gpus N --steps K ...    # the reference's CPU assembly on the host cores

A step = one full getDense() of the 2D fractional Laplacian (s=0.75, P1,
infinite horizon, zero exterior) on a synthetic disc mesh.  One JSON line is
printed by rank 0.  See DESIGN.md (Measurement) for how every field is formed.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

METRIC = 'matrix entries assembled/sec (2D frac. Laplacian P1, FP64)'
UNIT = 'entries/s'
S_ORDER = 0.75
TARGET_ORDER = 0.5

# workload: name -> (polygon sides, refinements).  N = 1 + nc/2 - sides*2^r/2
WORKLOADS = {
    'disc20k': (10, 6),    # 40 960 cells, 20 161 DoFs, 3.25 GB   (BASELINE.json configs[1])
    'disc12k': (6, 6),     # hexagon r=6: 24 576 cells, 12 097 DoFs
    'disc49k': (6, 7),     # hexagon r=7: 98 304 cells, 48 769 DoFs, 19 GB
    'disc3k': (6, 5),      # hexagon r=5: 2 977 DoFs
    'disc105k': (13, 7),   # 212 992 cells, ~105.7k DoFs, 89 GB
}


def make_mesh(workload):
    import pynucleus_b200 as pb
    sides, noRef = WORKLOADS[workload]
    mesh = pb.refined(pb.polygon_disc(sides), noRef)
    dm = pb.P1_DoFMap(mesh)
    return mesh, dm


def golden_probes(N, nrows, seed=20161):
    """sampled rows and probe vectors (ones, linspace, two Gaussian) of the size-N goldens in tests/golden/*_rows.npz
    (oracle/refbuild/make_golden_big.py evaluates the reference's getDense on exactly these)"""
    rng = np.random.default_rng(seed)
    rows = np.sort(rng.choice(N, min(nrows, N), replace=False))
    X = np.stack((np.ones(N), np.linspace(-1., 1., N), rng.standard_normal(N), rng.standard_normal(N)), axis=1)
    return rows, X


# --------------------------------------------------------------------------
# clocks
# --------------------------------------------------------------------------
class ClockSampler:
    Q = ('index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,'
         'clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,'
         'clocks_event_reasons.sw_power_cap')

    def __init__(self, gpu=0):
        self.gpu = gpu
        self.rows = []
        self.proc = None

    def start(self):
        try:
            self.proc = subprocess.Popen(['nvidia-smi', '-i', str(self.gpu), '--query-gpu='+self.Q,
                                          '--format=csv,noheader,nounits', '-lms', '100'],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.thread = threading.Thread(target=self._read, daemon=True)
            self.thread.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append(line.strip().split(', '))

    def stop(self):
        if self.proc is None:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons = [], [], set()
        for r in self.rows:
            try:
                sm.append(float(r[1]))
                smax.append(float(r[2]))
                for name, v in zip(('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap'), r[5:9]):
                    if v.strip() == 'Active':
                        reasons.add(name)
            except Exception:
                pass
        return {'sm_mhz': float(np.median(sm)) if sm else None, 'sm_max_mhz': max(smax) if smax else None,
                'reasons': sorted(reasons), 'samples': len(sm)}


# --------------------------------------------------------------------------
# algorithmic work
# --------------------------------------------------------------------------
def algorithmic_flops(hist, builder):
    """sum over the panel histogram of nq * (F_map + 7 + 3*ne) FP64 flops; pow counted separately"""
    from pynucleus_b200 import quadrature
    sing = builder.problem.singular
    flops = 0.
    pows = 0.
    for panel, count in hist.items():
        if panel >= 1:
            n = quadrature.regular(panel, 2)[1].shape[0]
            nq, per = n*n, 70.
        elif panel == -3:
            nq, per = sing['identical'][1].shape[0], 24+7+3*6
        elif panel == -2:
            nq, per = sing['edge'][1].shape[0], 24+7+3*10
        elif panel == -1:
            nq, per = sing['vertex'][1].shape[0], 24+7+3*15
        else:
            continue
        flops += count*nq*per
        pows += count*nq
    return flops, pows


# --------------------------------------------------------------------------
# CPU reference arm: the reference's own getDense on the host cores
# --------------------------------------------------------------------------
def _ref_worker(args):
    workload, rank, size = args
    sys.path.insert(0, os.path.join(ROOT, 'oracle', '_ref'))
    os.environ.setdefault('OMP_NUM_THREADS', '1')
    import numpy as np
    from mpi4py import MPI
    from PyNucleus_fem.mesh import mesh2d
    from PyNucleus_fem.DoFMaps import P1_DoFMap
    from PyNucleus_nl.kernels import getFractionalKernel
    from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder
    from PyNucleus_nl.fractionalOrders import constFractionalOrder
    mesh0, _ = make_mesh(workload)
    mesh = mesh2d(mesh0.vertices.copy(), mesh0.cells.copy())
    dm = P1_DoFMap(mesh)
    kernel = getFractionalKernel(2, constFractionalOrder(S_ORDER), np.inf)
    comm = MPI.fakeComm(rank, size)
    b = nonlocalBuilder(dm, kernel, {'target_order': TARGET_ORDER}, comm=comm)
    nc = mesh.num_cells
    start, end = int(np.ceil(nc*rank/size)), int(np.ceil(nc*(rank+1)/size))
    pairs = sum(nc-c for c in range(start, end))
    t = time.time()
    b.getDense()
    t = time.time()-t
    return pairs, t, dm.num_dofs, nc


def _port_worker(args):
    """the same slice with the CPU restatement (oracle/) -- only used where the stub-built reference is absent"""
    workload, rank, size = args
    import numpy as np
    import oracle
    mesh, dm = make_mesh(workload)
    P = oracle.Problem(mesh.vertices, mesh.cells, dm.dofs, dm.num_dofs, S_ORDER, bfacets=mesh.boundaryFacets,
                       target_order=TARGET_ORDER)
    nc = mesh.num_cells
    start, end = int(np.ceil(nc*rank/size)), int(np.ceil(nc*(rank+1)/size))
    pairs = sum(nc-c for c in range(start, end))
    t = time.time()
    P.dense(True, start, end, wrap=1 << 22)       # entries folded into a 32 MB buffer: the timing needs no N x N matrix
    t = time.time()-t
    return pairs, t, dm.num_dofs, nc


def reference_throughput(workload, target_seconds=15., cores=None):
    """entries/s of the reference's Cython getDense on `cores` host processes, each assembling one rank slice of
    a `size`-rank cell partition of the SAME mesh (nonlocalAssembly_{SCALAR}.pxi:1280-1285)"""
    import multiprocessing as mp
    have_ref = os.path.isdir(os.path.join(ROOT, 'oracle', '_ref', 'PyNucleus_nl')) and not os.environ.get('PNB_BENCH_FORCE_PORT')
    worker, kind = (_ref_worker, 'reference') if have_ref else (_port_worker, 'port')
    cores = cores or len(os.sched_getaffinity(0))
    sides, noRef = WORKLOADS[workload]
    nc = sides*4**noRef
    per_pair = 2.5e-6     # s per cell pair of the reference on one core (measured: ~2.2 us here)
    size = max(cores, int(np.ceil(nc*(nc+1)/2*per_pair/target_seconds)))
    ranks = [int(i*size/cores) for i in range(cores)]
    ctx = mp.get_context('spawn')
    t0 = time.time()
    with ctx.Pool(cores) as pool:
        res = pool.map(worker, [(workload, r, size) for r in ranks])
    wall = time.time()-t0
    pairs = sum(r[0] for r in res)
    tmax = max(r[1] for r in res)
    N, nc = res[0][2], res[0][3]
    entries_per_pair = N*N/(nc*(nc+1)/2.)
    return {'value': pairs*entries_per_pair/tmax, 'unit': UNIT, 'cores': cores, 'kind': kind,
            'sample': '{} of {} rank slices (ranks i*{}//{}) of the reference cell partition of the same mesh '
                      '({} of {} cell pairs); slowest slice {:.1f} s, wall {:.1f} s incl. import/mesh'.format(
                          cores, size, size, cores, pairs, nc*(nc+1)//2, tmax, wall),
            'seconds': tmax}


def run_reference(args):
    rank = int(os.environ.get('RANK', '0'))
    if rank != 0:
        return
    vals = []
    last = None
    for _ in range(args.warmup_ref+args.steps_ref):
        last = reference_throughput(args.workload, target_seconds=args.ref_seconds)
        if last is None:
            emit({'impl': 'reference', 'unavailable': 'oracle/_ref (stub-built reference) is missing'})
            return
        vals.append(last)
    vals = vals[args.warmup_ref:]
    v = float(np.mean([x['value'] for x in vals]))
    mesh, dm = make_mesh(args.workload)
    N = dm.num_dofs
    line = {'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': UNIT, 'n_gpus': args.gpus,
            'steps': args.steps_ref, 'warmup': args.warmup_ref, 'ms_per_step': 1e3*N*N/v, 'higher_is_better': True,
            'scaling': 'strong', 'vs_baseline': None, 'dtype': 'f64', 'data': 'synthetic',
            'config': config_block(args.workload, N, mesh.num_cells, max(1, args.gpus)),
            'note': ('CPU: reference Cython getDense (stub-built, oracle/_ref)' if last['kind'] == 'reference' else
                     'CPU: C restatement of the reference algorithm (oracle/; oracle/_ref is absent)')
                    + '; ms_per_step extrapolated from the sampled slices to the full matrix; `config` describes the workload and is '
                      'the CUDA arm\'s (l2 / parallelism refer to that arm)',
            'cpu_baseline': dict(last, value=v),
            'e2e': {'value': v, 'unit': UNIT, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}
    emit(line)


# --------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------
def config_block(workload, N, nc, world):
    """the `config` of the JSON line: identical in both arms at the same number of GPUs (the driver compares them)"""
    return {'workload': workload_string(workload, N, nc),
            'l2': 'output ({:.2f} GB/step) exceeds L2; 512 MB flush written between steps (untimed)'.format(N*N*8/1e9),
            'parallelism': 'single GPU' if world == 1 else
                           'rows owned by cell groups x{} (peer-memory fragments, no matrix collective)'.format(world)}


def workload_string(workload, N, nc):
    """the same description in both arms (the driver compares the config of the two lines)"""
    return ('{}: 2D disc ({}-gon fan, {} radial refinements), s={}, P1, infinite horizon, zero exterior, dense, N={} ({} cells), '
            'target_order={}'.format(workload, WORKLOADS[workload][0], WORKLOADS[workload][1], S_ORDER, N, nc, TARGET_ORDER))


def parity_block(workload, A_rows, rows, dev, world):
    """max relative entry error of this run's matrix against the REFERENCE's own getDense on the same mesh
    (tests/golden/<workload>_rows.npz: sampled rows, diagonal, A X for four seeded vectors; generated by
    oracle/refbuild/make_golden_big.py with the stub-built reference).  Every rank checks the rows it holds; the maxima
    are reduced over the ranks.  None when there is no golden for the workload."""
    import torch
    import torch.distributed as dist
    path = os.path.join(ROOT, 'tests', 'golden', workload+'_rows.npz')
    if not os.path.exists(path):
        return None
    g = np.load(path)
    N = int(g['num_dofs'])
    grows, X = golden_probes(N, g['rows'].shape[0])
    assert np.array_equal(grows, g['rows'])
    rows = np.asarray(rows, dtype=np.int64)
    pos = {int(r): k for k, r in enumerate(rows)}
    mine = [(k, pos[int(r)]) for k, r in enumerate(grows) if int(r) in pos]
    d = np.sqrt(np.abs(g['diagonal']))
    err_rows = 0.
    if mine:
        gi = np.array([m[0] for m in mine])
        li = torch.as_tensor(np.array([m[1] for m in mine]), device=dev)
        mineA = A_rows[li].cpu().numpy()
        scale = np.maximum(np.abs(g['A_rows'][gi]), 1e-2*np.outer(d[grows[gi]], d))
        err_rows = float((np.abs(mineA-g['A_rows'][gi])/scale).max())
    err_diag = err_ax = 0.
    if rows.shape[0]:
        rt = torch.as_tensor(rows, device=dev)
        diag = A_rows[torch.arange(rows.shape[0], device=dev), rt].cpu().numpy()
        err_diag = float(np.abs(diag/g['diagonal'][rows]-1).max())
        Xd = torch.as_tensor(X, device=dev)
        AX = (A_rows @ Xd).cpu().numpy()
        bound = (A_rows.abs() @ Xd.abs()).cpu().numpy()
        err_ax = float((np.abs(AX-g['AX'][rows])/bound).max())
    errs = torch.tensor([err_rows, err_diag, err_ax, float(len(mine))], dtype=torch.float64, device=dev)
    if world > 1:
        mx = errs.clone()
        dist.all_reduce(mx, op=dist.ReduceOp.MAX)
        cnt = errs[3:].clone()
        dist.all_reduce(cnt, op=dist.ReduceOp.SUM)
        errs = torch.cat((mx[:3], cnt))
    e = errs.cpu().tolist()
    return {'max_rel_err': max(e[:3]), 'rows': e[0], 'diagonal': e[1], 'products': e[2], 'rows_checked': int(e[3]),
            'golden': 'tests/golden/{}_rows.npz (reference getDense, stub-built, {} slices)'.format(workload, int(g['nslices'])),
            'bar': 1e-12}


# --------------------------------------------------------------------------
# CUDA arm
# --------------------------------------------------------------------------
def assemble_steps(builder, world, group, steps, warmup, flush, dev):
    """warm-up + timed steps of one workload; returns (ms per step list, per-kernel ms, launches, operator, A)"""
    import torch
    import torch.distributed as dist
    state = {'A': None, 'op': None}

    def step():
        if world == 1:
            if state['A'] is None:
                N = builder.dm.num_dofs
                state['A'] = torch.empty((N, N), dtype=torch.float64, device=dev)
            state['op'] = builder.getDense(out=state['A'])
        else:
            state['op'] = builder.getDenseDistributed(process_group=group, out=state['A'])
            if state['A'] is None and state['op'].A_rows is not None:
                state['A'] = state['op'].A_rows.device_data

    cold0 = time.perf_counter()
    step()
    torch.cuda.synchronize()
    cold_ms = (time.perf_counter()-cold0)*1e3
    for _ in range(max(warmup-1, 0)):
        step()
    torch.cuda.synchronize()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    tile_ms, launches = [], 0
    kms = {'f2': [], 'near': [], 'mix': [], 'symmetrize': []}
    for k in range(steps):
        flush.zero_()
        if world > 1:
            dist.barrier()
        ev[k][0].record()
        step()
        ev[k][1].record()
        torch.cuda.synchronize()
        st = builder.getStats()
        tile_ms.append(st['ms_tiles'])
        launches += st['launches']
        for key in kms:
            kms[key].append(st['ms_'+key])
    ms = [a.elapsed_time(b) for a, b in ev]
    if world > 1:
        # device time of a step = max over ranks
        t = torch.tensor(ms, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = t.cpu().tolist()
    return ms, tile_ms, kms, launches, cold_ms, state, step


def run_cuda(args):
    import torch
    import torch.distributed as dist
    import pynucleus_b200 as pb
    from pynucleus_b200 import _lib

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local_rank = int(os.environ.get('LOCAL_RANK', '0'))
    if not torch.cuda.is_available():
        raise RuntimeError('bench.py needs a CUDA device (no CPU fallback)')
    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group('nccl', device_id=torch.device('cuda', local_rank))
    dev = torch.device('cuda', local_rank)

    mesh, dm = make_mesh(args.workload)
    N = dm.num_dofs
    kernel = pb.getFractionalKernel(2, S_ORDER)
    params = {'target_order': TARGET_ORDER, 'device': local_rank}
    builder = pb.nonlocalBuilder(dm, kernel, params)

    flush = torch.empty(64*1024*1024, dtype=torch.float64, device=dev)   # 512 MB > L2
    pk = np.zeros(1)
    _lib.check(_lib.lib().pnb_fp64_peak(local_rank, pk.ctypes.data_as(_lib.c_double_p)))
    peak = float(pk[0])

    # clocks are sampled from the warm-up on: nvidia-smi needs a few hundred ms for its first line, longer than the
    # whole timed region of the short multi-GPU runs
    sampler = ClockSampler(local_rank)
    sampler.start()
    ms, tile_ms, kms, launches, cold_ms, state, step = assemble_steps(builder, world, None, args.steps, args.warmup, flush, dev)
    # a timed region shorter than the sampling period: keep the same load up (untimed, same count on every rank)
    # until the sampler has seen it
    busy_ms = (args.warmup+args.steps)*float(np.mean(ms))
    if busy_ms < 1500.:
        for _ in range(int(np.ceil((1500.-busy_ms)/max(float(np.mean(ms)), 1.)))):
            step()
        torch.cuda.synchronize()
    clocks = sampler.stop()
    ms_step = float(np.mean(ms))
    st = builder.getStats()
    value = N*float(N)/(ms_step*1e-3)
    op = state['op']
    A = state['A']
    rows = np.arange(N) if world == 1 else op.rows
    nrows = int(rows.shape[0])

    # correctness of THIS run's matrix at the benchmarked size, on every rank
    parity = parity_block(args.workload, A, rows, dev, world)

    # second kernel of the path: y = A x on the assembled rows (HBM-bound, reads the matrix once)
    matvec = None
    Aop = xv = yv = None
    if nrows > 0:
        Aop = pb.Dense_LinearOperator(A[:nrows], local_rank)
        xv = torch.ones(N, dtype=torch.float64, device=dev)
        yv = torch.empty(nrows, dtype=torch.float64, device=dev)
        for _ in range(3):
            Aop.matvec_device(xv, yv)
        mv = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(10)]
        for a, b in mv:
            flush.zero_()
            a.record()
            Aop.matvec_device(xv, yv)
            b.record()
        torch.cuda.synchronize()
        mv_ms = float(np.median([a.elapsed_time(b) for a, b in mv]))
        hbm_peak = None
        try:
            with open(os.path.join(ROOT, 'MEASURED_PEAKS.json')) as f:
                hbm_peak = float(json.load(f)['hbm_gbs'])
        except Exception:
            pass
        gbs = (nrows*N*8+N*8+nrows*8)/(mv_ms*1e-3)/1e9
        matvec = {'kernel': 'matvec_kernel', 'bound': 'hbm', 'ms': mv_ms, 'achieved': gbs, 'unit': 'GB/s',
                  'peak': hbm_peak, 'frac': gbs/hbm_peak if hbm_peak else None,
                  'note': 'algorithmic bytes = 8*(rows*N + N + rows); L2 flushed between launches; peak = MEASURED_PEAKS.json hbm_gbs'}

    # roofline: algorithmic FP64 flops / device time measured with CUDA events inside the C call.
    # Headline object = the dominant kernel; the other pair kernels and their sum are listed beside it.
    hist = builder.getPanelHistogram()
    share = st['evaluated_pairs']/max(st['distinct_pairs'], 1) if world > 1 else 1.   # several GPUs: this rank's share
    flops_all, pows_all = algorithmic_flops(hist, builder)
    near_hist = {k: v for k, v in hist.items() if k < 0 or k > 5}
    flops_near, pows_near = algorithmic_flops(near_hist, builder)
    flops_f2, pows_f2 = st['f2_pairs']*9*70., st['f2_pairs']*9.
    flops_mix = (flops_all-flops_near)*share-flops_f2
    t_all = float(np.mean(tile_ms))*1e-3
    t = {key: float(np.mean(v))*1e-3 for key, v in kms.items()}
    per_kernel = {
        'gmix_kernel': {'ms': t['mix']*1e3, 'tflops': flops_mix/max(t['mix'], 1e-9)/1e12},
        'gnear_eval_kernel': {'ms': t['near']*1e3, 'tflops': flops_near*share/max(t['near'], 1e-9)/1e12},
        'gf2_kernel': {'ms': t['f2']*1e3, 'tflops': flops_f2/max(t['f2'], 1e-9)/1e12},
        'all pair kernels': {'ms': t_all*1e3, 'tflops': flops_all*share/t_all/1e12}}
    for v in per_kernel.values():
        v['frac'] = v['tflops']/peak if peak else None
    dominant = max(('gmix_kernel', 'gnear_eval_kernel', 'gf2_kernel'), key=lambda k: per_kernel[k]['ms'])
    achieved = per_kernel[dominant]['tflops']
    # dram bytes per launch of the dominant kernel: from the committed ncu capture of the same kernel on the same mesh
    # (not measurable inside a timed run); None when the summary is absent or for another workload
    traffic, traffic_src = None, 'no ncu summary for this workload'
    try:
        if args.workload == 'disc20k' and world == 1:
            with open(os.path.join(ROOT, 'profiles', 'r2_final_ncu_summary.json')) as f:
                summ = json.load(f)
            traffic = float(summ[dominant]['dram_bytes'])
            traffic_src = 'profiles/r2_final_ncu_summary.json ({})'.format(summ.get('_source', 'ncu --set full'))
    except Exception:
        pass
    roofline = {'bound': 'fp64', 'kernel': dominant, 'achieved': achieved, 'peak': peak, 'unit': 'TFLOP/s',
                'frac': achieved/peak if peak else None,
                'traffic': traffic,
                'kernels': per_kernel,
                'frac_all_pair_kernels': per_kernel['all pair kernels']['tflops']/peak if peak else None,
                'note': 'algorithmic flops (SURVEY 8d: 70/node-pair regular, 49/61/76 singular; pow excluded) of the cell pairs '
                        'the kernel evaluates / its device time (CUDA events around the launch inside the C call, averaged '
                        'over the timed steps); peak = DFMA microbenchmark run in this process (MEASURED_PEAKS.json has no '
                        'FP64 figure); traffic = dram__bytes_read.sum + dram__bytes_write.sum of one launch of the dominant kernel, '
                        'taken from ' + traffic_src + ' (algorithmic output of the whole assembly: 8 N^2 bytes); pow evaluations/s over all pair kernels = {:.3e}; evaluated/distinct '
                        'pairs {:.3f}; pair-kernel share of step {:.3f}'.format(
                            pows_all*share/t_all, st['evaluated_pairs']/max(st['distinct_pairs'], 1), t_all*1e3/ms_step)}

    # end to end through the host-buffer entry point: a NEW builder every step (mesh / DoFMap / table upload, schedules,
    # near pair list; several GPUs: plan, staging buffers, peer-memory handshake), assembly, copy of the rows to pinned
    # host memory
    del op
    state['op'] = None
    builder.releaseScratch()
    if world > 1:
        torch.cuda.empty_cache()
    host = torch.empty((max(nrows, 1), N), dtype=torch.float64).pin_memory()
    hA = host.numpy()
    e2e_ms = []
    h2d = (mesh.vertices.nbytes+mesh.cells.nbytes+dm.dofs.nbytes+mesh.volVector.nbytes+mesh.hVector.nbytes
           + mesh.boundaryFacets.nbytes)
    checksum = None
    for k in range(max(1, min(args.steps, 3))+1):
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        t0 = time.perf_counter()
        b2 = pb.nonlocalBuilder(dm, kernel, params)      # uploads mesh, DoFMap, tables (host -> device)
        b2.problem
        t1 = time.perf_counter()
        if world == 1:
            b2.getDenseHost(out=hA)                       # assembly + device -> host copy inside the C call
            if rank == 0 and os.environ.get('PNB_BENCH_VERBOSE'):
                print('e2e: setup %.1f ms, assemble+copy %.1f ms' % ((t1-t0)*1e3, (time.perf_counter()-t1)*1e3), file=sys.stderr)
        else:
            o2 = b2.getDenseDistributed(out=A)
            if nrows:
                host[:nrows].copy_(A[:nrows])
            torch.cuda.synchronize()
            del o2
        if rank == 0:
            checksum = float(hA[0, 0])
        dt = (time.perf_counter()-t0)*1e3
        if world > 1:
            tt = torch.tensor([dt], dtype=torch.float64, device=dev)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            dt = float(tt.item())
        e2e_ms.append(dt)
        b2.releaseScratch()
        del b2
    e2e_ms = e2e_ms[1:]
    e2e_value = N*float(N)/(float(np.mean(e2e_ms))*1e-3)

    # BASELINE configs[4] (>= 100k DoFs, the north-star size) in the same line: a few steps of disc105k
    north = None
    if args.north_star and args.workload != 'disc105k':
        del host, hA, A
        Aop = xv = yv = None       # noqa: F841  (views of the disc20k matrix)
        state['A'] = None
        builder = None
        torch.cuda.empty_cache()
        _lib.lib().pnb_release_cached_memory()
        try:
            north = north_star_leg(args, world, rank, local_rank, dev, flush, peak)
        except Exception as e:       # e.g. out of memory on a small part count: reported, not fatal
            north = {'workload': 'disc105k', 'error': '{}: {}'.format(type(e).__name__, str(e)[:300])}

    # the rows next to the hot path (SURVEY 8f / verdict rows): H2 product on the library's kernels, P2 elements, an
    # unsymmetric piecewise order -- small sizes, a few seconds, one GPU only
    widening = None
    if rank == 0 and world == 1 and not args.no_widening:
        try:
            widening = widening_leg(dev)
        except Exception as e:
            widening = {'error': '{}: {}'.format(type(e).__name__, str(e)[:300])}

    cpu = None
    if not args.no_cpu_baseline and rank == 0 and world == 1:
        cpu = reference_throughput(args.workload, target_seconds=args.ref_seconds)
    line = {'metric': METRIC, 'value': value, 'unit': UNIT, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
            'ms_per_step': ms_step, 'higher_is_better': True, 'scaling': 'strong', 'vs_baseline': None,
            'dtype': 'f64', 'data': 'synthetic',
            'config': config_block(args.workload, N, mesh.num_cells, world),
            'clocks': clocks,
            'e2e': {'value': e2e_value, 'unit': UNIT, 'h2d_bytes_per_step': int(h2d)*world, 'd2h_bytes_per_step': int(N*N*8),
                    'ms_per_step': float(np.mean(e2e_ms)), 'checksum_A00': checksum},
            'cold_ms_first_assembly': cold_ms,
            'gpu_launches': int(launches),
            'parity': parity,
            'roofline': roofline,
            'matvec': matvec,
            'north_star': north,
            'widening': widening,
            'cpu_baseline': cpu,
            'phases_ms': {'tiles': st['ms_tiles'], 'boundary': st['ms_boundary'], 'reduce_scatter': st['ms_reduce_scatter']},
            'pairs': {'distinct': st['distinct_pairs'], 'evaluated': st['evaluated_pairs']}}
    if rank == 0:
        emit(line)
    if world > 1:
        pb.release_staging_pool()
        dist.destroy_process_group()


def _event_ms(fn, n, warm=3):
    import torch
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(n):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1)/n


def widening_leg(dev):
    """H2 operator (N = 12 097): assembly time, product on the library's own kernels against the dense product; P2 elements
    (2 977 dofs) and an unsymmetric leftRight order (2 977 dofs): dense assembly time and symmetry of the result"""
    import torch
    import pynucleus_b200 as pb
    out = {}
    params = {'target_order': TARGET_ORDER, 'device': dev.index}
    mesh = pb.refined(pb.uniform_disc(), 6)
    dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, S_ORDER), params)
    A = b.getDense()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    H = b.getH2()
    torch.cuda.synchronize()
    t_h2 = time.perf_counter()-t0
    x = torch.as_tensor(np.sin(np.arange(dm.num_dofs)*0.37)+0.1, device=dev)
    y = torch.empty_like(x)
    yh, yd = H.matvec_device(x).clone(), A.matvec_device(x).clone()
    out['h2'] = {'workload': 'disc r=6, N={}, s={}'.format(dm.num_dofs, S_ORDER), 'getH2_s': t_h2,
                 'near_nnz_frac': H.Anear.nnz/float(dm.num_dofs)**2, 'far_pairs': sum(len(v) for v in H.Pfar.values()),
                 'matvec_ms': _event_ms(lambda: H.matvec_device(x, y), 50), 'dense_matvec_ms': _event_ms(lambda: A.matvec_device(x, y), 50),
                 'rel_diff_vs_dense_product': float((yh-yd).abs().max()/yd.abs().max()),
                 'kernels': 'pnb_h2_matvec (csr_matvec, h2_leaf_up, h2_transfer_up, h2_far, h2_far_sum, h2_transfer_down, h2_leaf_down)'}
    del H, A, b
    mesh = pb.refined(pb.uniform_disc(), 4)
    dm2 = pb.P2_DoFMap(mesh)
    b2 = pb.nonlocalBuilder(dm2, pb.getFractionalKernel(2, S_ORDER), params)
    A2 = b2.getDense()
    ms = _event_ms(lambda: b2.getDense(out=A2.device_data), 3, warm=1)
    d = A2.device_data
    out['p2'] = {'workload': 'disc r=4, P2, {} dofs ({} cells), s={}'.format(dm2.num_dofs, mesh.num_cells, S_ORDER), 'ms_per_assembly': ms,
                 'entries_per_s': dm2.num_dofs**2/(ms*1e-3), 'symmetry_rel': float((d-d.t()).abs().max()/d.abs().max()),
                 'min_diagonal': float(torch.diagonal(d).min()), 'kernel': 'elem_rows_kernel<2,2>'}
    del A2, b2
    mesh = pb.refined(pb.uniform_disc(), 5)
    dm = pb.P1_DoFMap(mesh)
    s = pb.leftRightFractionalOrder(0.25, 0.75, 0.6, 0.4)
    b3 = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, s), params)
    A3 = b3.getDense()
    ms = _event_ms(lambda: b3.getDense(out=A3.device_data), 3, warm=1)
    d = A3.device_data
    out['unsymmetric_order'] = {'workload': 'disc r=5, N={}, leftRight(0.25, 0.75, 0.6, 0.4)'.format(dm.num_dofs), 'ms_per_assembly': ms,
                                'passes': len(b3._classes['passes']), 'symmetry_rel': float((d-d.t()).abs().max()/d.abs().max()),
                                'min_diagonal': float(torch.diagonal(d).min())}
    del A3, b3
    # BASELINE configs[2] flavour at size: finite horizon (l2 ball), fractional and constant kernels, dense operator
    mesh = pb.refined(pb.uniform_disc(), 6)
    dm = pb.P1_DoFMap(mesh)
    fin = {}
    for name, k in (('fractional s=0.4', pb.getFractionalKernel(2, 0.4, 0.15)), ('constant', pb.getIntegrableKernel(2, 'constant', 0.15))):
        b4 = pb.nonlocalBuilder(dm, k, params)
        A4 = b4.getDense()
        ms = _event_ms(lambda: b4.getDense(out=A4.device_data), 2, warm=1)
        d = A4.device_data
        fin[name] = {'ms_per_assembly': ms, 'nonzero_frac': float((d != 0).sum())/d.numel(), 'symmetric': bool(torch.equal(d, d.t())),
                     'min_diagonal': float(torch.diagonal(d).min())}
        del A4, b4
    out['finite_horizon'] = {'workload': 'disc r=6, N={}, horizon 0.15 (l2 ball), P1, dense'.format(dm.num_dofs), **fin}
    # kernels outside the power-table path, on the row-owner kernels: an order that varies inside the cells (the driver's
    # twoDomainNonSym: order, scaling constant and a general power per quadrature node) and a tempered kernel
    mesh = pb.refined(pb.uniform_disc(), 5)
    dm = pb.P1_DoFMap(mesh)
    for name, k, kern in (('order_inside_cells', pb.getFractionalKernel(2, pb.smoothedLeftRightFractionalOrder(0.25, 0.75)), 'varorder_rows_kernel<2,1>'),
                          ('tempered', pb.getFractionalKernel(2, S_ORDER, tempered=2.), 'elem_rows_kernel<2,1>'),
                          ('gaussian', pb.getIntegrableKernel(2, 'gaussian', np.inf, variance=0.1), 'elem_rows_kernel<2,1>')):
        try:
            b5 = pb.nonlocalBuilder(dm, k, params)
            A5 = b5.getDense()
            ms = _event_ms(lambda: b5.getDense(out=A5.device_data), 2, warm=1)
            d = A5.device_data
            out[name] = {'workload': 'disc r=5, N={}, {}'.format(dm.num_dofs, k), 'ms_per_assembly': ms,
                         'entries_per_s': dm.num_dofs**2/(ms*1e-3), 'asymmetry_rel': float((d-d.t()).abs().max()/d.abs().max()),
                         'min_diagonal': float(torch.diagonal(d).min()), 'kernel': kern}
            del A5, b5
        except Exception as e:
            out[name] = {'error': '{}: {}'.format(type(e).__name__, str(e)[:200])}
    return out


def north_star_leg(args, world, rank, local_rank, dev, flush, peak):
    """disc105k (BASELINE configs[4]: 105 665 DoFs, 212 992 cells, 89 GB of entries) on the same GPUs: device time per
    assembly, entries/s, structural checks of the result (no golden at this size: symmetric products, positive
    diagonal, every pair once)"""
    import torch
    import torch.distributed as dist
    import pynucleus_b200 as pb
    mesh, dm = make_mesh('disc105k')
    N = dm.num_dofs
    kernel = pb.getFractionalKernel(2, S_ORDER)
    builder = pb.nonlocalBuilder(dm, kernel, {'target_order': TARGET_ORDER, 'device': local_rank})
    steps = 2
    ms, tile_ms, kms, launches, cold_ms, state, step = assemble_steps(builder, world, None, steps, 1, flush, dev)
    st = builder.getStats()
    op, A = state['op'], state['A']
    rows = np.arange(N) if world == 1 else op.rows
    nrows = int(rows.shape[0])
    # structural checks: x^T A y == y^T A x (symmetry through two distributed products), positive diagonal, pairs once
    rng = np.random.default_rng(11)
    x = torch.as_tensor(rng.standard_normal(N), device=dev)
    y = torch.as_tensor(rng.standard_normal(N), device=dev)
    Aop = op if world > 1 else pb.Dense_LinearOperator(A, local_rank)
    Ax, Ay = Aop.matvec_device(x).clone(), Aop.matvec_device(y).clone()
    sym = abs(float(torch.dot(y, Ax)-torch.dot(x, Ay)))/max(abs(float(torch.dot(y, Ax))), 1e-300)
    dmin = float(Aop.diagonal_device().min()) if world > 1 else float(torch.diagonal(A).min())
    ev = torch.tensor([float(st['evaluated_pairs'])], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(ev)
    ms_step = float(np.mean(ms))
    t_all = float(np.mean(tile_ms))*1e-3
    out = {'workload': workload_string('disc105k', N, mesh.num_cells), 'steps': steps, 'warmup': 1, 'ms_per_step': ms_step,
           'value': N*float(N)/(ms_step*1e-3), 'unit': UNIT, 'rows_this_rank': nrows,
           'kernels_ms_rank0': {k: float(np.mean(v)) for k, v in kms.items()}, 'pair_kernels_ms_rank0': t_all*1e3,
           'checks': {'symmetry_rel': sym, 'min_diagonal': dmin, 'pairs_evaluated': int(ev.item()),
                      'pairs_distinct': int(st['distinct_pairs'])},
           'cold_ms_first_assembly': cold_ms}
    del op, A
    state['op'] = state['A'] = None
    builder.releaseScratch()
    return out


_REAL_STDOUT = None


def emit(line):
    """the one JSON line, on the real stdout"""
    data = (json.dumps(line)+'\n').encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    # libraries (NCCL: "NCCL version ..." with NCCL_DEBUG set) write to stdout; the contract is ONE JSON line there,
    # so everything else is sent to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--impl', default='cuda', choices=['cuda', 'reference'])
    ap.add_argument('--workload', default='disc20k', choices=sorted(WORKLOADS))
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--ref-seconds', type=float, default=15.)
    ap.add_argument('--north-star', dest='north_star', action='store_true', default=True,
                    help='also time a few assemblies of disc105k (BASELINE configs[4]) and report them under north_star')
    ap.add_argument('--no-north-star', dest='north_star', action='store_false')
    ap.add_argument('--no-widening', dest='no_widening', action='store_true', default=False,
                    help='skip the H2 / P2 / unsymmetric-order legs (a few seconds on one GPU)')
    args = ap.parse_args()
    # the reference arm runs the same W + K steps; every step is a bounded sample of the workload (rank slices of the
    # reference's own cell partition), sized so that the whole run takes about two minutes of CPU work plus process start-up
    args.steps_ref = max(1, args.steps)
    args.warmup_ref = max(0, args.warmup)
    if args.impl == 'reference':
        args.ref_seconds = min(args.ref_seconds, 120./(args.steps_ref+args.warmup_ref))
        run_reference(args)
    else:
        run_cuda(args)


if __name__ == '__main__':
    main()
