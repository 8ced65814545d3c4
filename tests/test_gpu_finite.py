"""GPU parity tests, finite horizon: the CUDA path against fixtures produced by the reference
itself (oracle/refbuild/make_golden_finite.py): fractional, constant and inverse-distance kernels on the l2 ball, 1D and 2D.

bit-exact: pair classification incl. REMOTE pairs (IGNORED) and quadrature orders
1e-12 relative: local matrices (regular, singular and horizon-cut pairs) and assembled entries
"""
import os

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

CASES_1D = ['finite_interval_frac0.25_r5', 'finite_interval_frac0.75_r5', 'finite_interval_constant_r5', 'finite_interval_invdist_r5']
CASES_2D = ['finite_disc_frac0.75_r3', 'finite_disc_frac0.25_r3', 'finite_disc_constant_r3', 'finite_disc_invdist_r3']
TOL = 1e-12


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name+'.npz'))


def kernel_from_golden(g):
    import pynucleus_b200 as pb
    dim = g['vertices'].shape[1]
    kt = str(g['kernel_type'])
    if kt == 'fractional':
        return pb.getFractionalKernel(dim, float(g['s']), float(g['horizon']))
    return pb.getIntegrableKernel(dim, kt, float(g['horizon']))


def builder_from_golden(g):
    import pynucleus_b200 as pb
    dim = g['vertices'].shape[1]
    bf = g['boundaryEdges'] if dim == 2 else g['boundaryVertices'].reshape(-1, 1)
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=bf)
    dm = pb.P1_DoFMap(mesh)
    assert dm.num_dofs == int(g['num_dofs'])
    params = {'target_order': float(g['target_order'])} if dim == 2 else {}
    b = pb.nonlocalBuilder(dm, kernel_from_golden(g), params)
    assert not b.zeroExterior      # nonlocalAssembly_{SCALAR}.pxi:918-921
    return b


def relerr_rows(C, Cref):
    scale = np.abs(Cref).max(axis=1, keepdims=True)
    scale[scale == 0] = 1.
    return (np.abs(C-Cref)/scale).max()


def entry_err(A, Aref):
    d = np.sqrt(np.abs(np.diag(Aref)))
    scale = np.maximum(np.abs(Aref), 1e-2*np.outer(d, d))
    return (np.abs(A-Aref)/scale).max()


@pytest.mark.parametrize('name', CASES_1D+CASES_2D)
def test_kernel_values_and_scaling(golden_dir, name):
    g = load(golden_dir, name)
    k = kernel_from_golden(g)
    assert abs(k.scalingValue-float(g['scaling'])) <= 1e-15*abs(float(g['scaling']))
    assert k.singularityValue == float(g['singularity'])
    vals = np.array([k(x, y) for x, y in zip(g['kx'], g['ky'])])
    assert np.allclose(vals, g['kvals'], rtol=1e-14, atol=0.)
    assert (vals[-2:] == 0.).all() and (vals[:4] != 0.).all()


@pytest.mark.parametrize('name', CASES_1D+CASES_2D+['finite_disc_frac0.4_r4'])
def test_classification_bit_exact(golden_dir, name):
    g = load(golden_dir, name)
    b = builder_from_golden(g)
    nc = g['cells'].shape[0]
    iu = np.triu_indices(nc)
    panel = b.getPanelTypes(np.stack(iu, axis=1))
    assert np.array_equal(panel, g['panel_matrix'][iu])
    assert (panel == -6).sum() > 0
    hist = b.getPanelHistogram()
    ref = {int(k): int(v) for k, v in zip(*np.unique(g['panel_matrix'][iu], return_counts=True)) if k != -6}
    assert hist == ref


@pytest.mark.parametrize('name', CASES_1D+CASES_2D+['finite_disc_frac0.4_r4'])
def test_local_matrices_vs_reference(golden_dir, name):
    g = load(golden_dir, name)
    b = builder_from_golden(g)
    panel, C = b.getLocalMatrices(g['pairs'])
    assert np.array_equal(panel, g['panels'])
    rel = g['relpos'][g['pairs'][:, 0], g['pairs'][:, 1]]
    assert (rel == 2).sum() >= 20          # pairs cut by the horizon are in the sample
    assert relerr_rows(C, g['contribs']) < TOL


@pytest.mark.parametrize('name', CASES_1D+CASES_2D+['finite_disc_frac0.4_r4'])
def test_dense_vs_reference(golden_dir, name):
    g = load(golden_dir, name)
    b = builder_from_golden(g)
    A = b.getDense().data
    assert entry_err(A, g['A']) < TOL
    assert np.abs(A-A.T).max() <= 1e-15*np.abs(A).max()
    assert b.getStats()['evaluated_pairs'] > 0
