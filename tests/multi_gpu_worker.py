"""Launched by test_gpu_parity.test_two_gpus with torchrun (one process per GPU, NCCL): row-block assembly,
distributed matvec and CG against the single-GPU operator.  Prints OK on rank 0."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.abspath(os.path.join(os.path.dirname(__file__), '..')))
import pynucleus_b200 as pb  # noqa: E402


def main():
    rank = int(os.environ['RANK'])
    local = int(os.environ.get('LOCAL_RANK', rank))
    torch.cuda.set_device(local)
    dist.init_process_group('nccl', device_id=torch.device('cuda', local))
    mesh = pb.refined(pb.uniform_disc(), 4)
    dm = pb.P1_DoFMap(mesh)
    kernel = pb.getFractionalKernel(2, 0.75)
    full = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5, 'device': local}).getDense()
    op = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5, 'device': local}).getDenseDistributed()
    A = full.device_data
    mine = op.A_rows.device_data
    rows = torch.as_tensor(op.rows, device='cuda')
    # every rank evaluates a share of the cell pairs and the blocks reach the row owners through NVLink peer stores: same
    # terms as on one GPU, different summation order -> 1e-12 relative to the entry / diagonal scale
    d = torch.sqrt(torch.diagonal(A))
    scale = torch.maximum(A[rows].abs(), 1e-2*torch.outer(d[rows], d))
    err = float(((mine-A[rows]).abs()/scale).max())
    assert err < 1e-12, 'rows differ from the single-GPU operator: %g' % err
    counts = [None]*dist.get_world_size()
    dist.all_gather_object(counts, int(rows.shape[0]))
    assert sum(counts) == dm.num_dofs
    x = torch.from_numpy(np.random.default_rng(3).standard_normal(dm.num_dofs)).cuda()
    y = op.matvec_device(x)
    y1 = full.matvec_device(x)
    assert float((y-y1).abs().max()) < 1e-12*float(y1.abs().max()), 'distributed matvec differs'
    b = torch.full((dm.num_dofs,), 1e-3, dtype=torch.float64, device='cuda')
    u, its, res = pb.cg(op, b, tol=1e-12)
    u1, its1, res1 = pb.cg(full, b, tol=1e-12)
    assert abs(its-its1) <= 1 and float((u-u1).abs().max()) < 1e-9*float(u1.abs().max()), 'CG solutions differ'
    assert float((full.matvec_device(u)-b).abs().max()) < 1e-11
    # tables too small for some pairs (max_regular_order below the orders the mesh needs): only the ranks that meet such a
    # pair see the error, the retry with larger tables is decided by all ranks together (no rank re-enters a collective alone)
    small = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5, 'device': local, 'max_regular_order': 4})
    op2 = small.getDenseDistributed()
    assert small.problem.max_order > 4
    err2 = float(((op2.A_rows.device_data-A[rows]).abs()/scale).max())
    assert err2 < 1e-12, 'rows differ after the collective table retry: %g' % err2
    # row-owner kernels (here: P2 elements, an order that varies inside the cells): rows dealt to the ranks, every row complete
    # on its owner, nothing exchanged during the assembly; rows bitwise equal to the single-GPU operator, GMRES on the
    # unsymmetric distributed operator
    mesh3 = pb.refined(pb.uniform_disc(), 3)
    for dm3, k3 in ((pb.P2_DoFMap(mesh3), kernel),
                    (pb.P1_DoFMap(mesh3), pb.getFractionalKernel(2, pb.smoothedLeftRightFractionalOrder(0.25, 0.75, r=0.3)))):
        b3 = pb.nonlocalBuilder(dm3, k3, {'target_order': 0.5, 'device': local})
        full3 = b3.getDense()
        op3 = b3.getDenseDistributed()
        r3 = torch.as_tensor(op3.rows, device='cuda')
        assert torch.equal(op3.A_rows.device_data, full3.device_data[r3]), 'row-owner rows differ from the single-GPU operator'
        dist.all_gather_object(counts, int(r3.shape[0]))
        assert sum(counts) == dm3.num_dofs
        x3 = torch.from_numpy(np.random.default_rng(5).standard_normal(dm3.num_dofs)).cuda()
        y3, y31 = op3.matvec_device(x3), full3.matvec_device(x3)
        assert float((y3-y31).abs().max()) < 1e-12*float(y31.abs().max()), 'distributed matvec (row-owner kernels) differs'
    dist.barrier()
    if rank == 0:
        print('OK', dm.num_dofs, its)
    pb.release_staging_pool()
    dist.destroy_process_group()


if __name__ == '__main__':
    main()
