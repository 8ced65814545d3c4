"""The oracle (oracle/) against the REFERENCE's own outputs.

tests/golden/*.npz were produced by the stub-built reference
(oracle/refbuild/make_golden.py): panel types, vertex permutations, local
matrices and assembled dense matrices.  These tests pin the C/numpy
restatement before anything else is compared against it.
"""
import os

import numpy as np
import pytest

import oracle
from oracle import meshes, tables

CASES_2D = ['disc_s0.75_r1', 'disc_s0.75_r2', 'disc_s0.25_r2', 'disc_s0.75_r3']
CASES_1D = ['interval_s0.25_r3', 'interval_s0.25_r6', 'interval_s0.75_r5']


def load(golden_dir, name):
    return np.load(os.path.join(golden_dir, name+'.npz'))


def problem_from_golden(g):
    dim = g['vertices'].shape[1]
    to = float(g['target_order']) if dim == 2 else None
    bf = g['boundaryEdges'] if dim == 2 else g['boundaryVertices'].reshape(-1, 1)
    return oracle.Problem(g['vertices'], g['cells'], g['dofs'], int(g['num_dofs']), float(g['s']),
                          bfacets=bf, target_order=to, hVector=g['hVector'], volVector=g['volVector'],
                          hmin=float(g['hmin']), diam=float(g['diam']))


@pytest.mark.parametrize('r', [0, 4])
def test_disc_mesh_matches_reference(golden_dir, r):
    g = load(golden_dir, 'disc_mesh_r%d' % r)
    m = meshes.disc(r)
    assert np.array_equal(m.cells, g['cells'])
    assert np.abs(m.vertices-g['vertices']).max() < 1e-15
    dofs, n = meshes.p1_dofs(m)
    assert n == int(g['num_dofs'])
    assert np.array_equal(np.where(dofs >= 0, dofs, -1), np.where(g['dofs'] >= 0, g['dofs'], -1))
    assert set(map(tuple, m.boundary_facets())) == set(map(tuple, g['boundaryEdges']))
    assert np.allclose(m.hVector, g['hVector'], rtol=1e-14)
    assert np.allclose(m.volVector, g['volVector'], rtol=1e-13)


def test_interval_mesh_matches_reference(golden_dir):
    g = load(golden_dir, 'interval_s0.25_r6')
    m = meshes.interval(-1., 1., 6)
    assert np.array_equal(m.cells, g['cells'])
    assert np.array_equal(m.vertices, g['vertices'])
    dofs, n = meshes.p1_dofs(m)
    assert n == 63
    assert np.array_equal(np.where(dofs >= 0, dofs, -1), np.where(g['dofs'] >= 0, g['dofs'], -1))


@pytest.mark.parametrize('name', CASES_2D)
def test_singular_tables_2d(golden_dir, name):
    g = load(golden_dir, name)
    s = float(g['s'])
    H0 = float(g['diam'])/np.sqrt(8.)
    o = tables.diag_orders(2, -2-2*s, -1-2*s, float(g['hmin']), H0, int(g['num_dofs']), 0.5)
    assert o['qod'] == int(g['quad_order_diagonal'])
    assert o['qodV'] == int(g['quad_order_diagonalV'])
    assert o['b_qod'] == int(g['boundary_quad_order_diagonal'])
    R = tables.near_rules(2, -2-2*s, -1-2*s, o)
    for nm, key in (('qrId', ('interior', -3)), ('qrEdge', ('interior', -2)), ('qrVertex', ('interior', -1)),
                    ('bqrEdge', ('boundary', -2)), ('bqrVertex', ('boundary', -1))):
        b, w = R[key]
        assert b.shape == g[nm+'_nodes'].shape
        assert np.abs(b-g[nm+'_nodes']).max() < 1e-15
        assert np.allclose(w, g[nm+'_weights'], rtol=1e-14, atol=0)


@pytest.mark.parametrize('name', CASES_1D)
def test_singular_tables_1d(golden_dir, name):
    g = load(golden_dir, name)
    s = float(g['s'])
    H0 = float(g['diam'])/np.sqrt(8.)
    o = tables.diag_orders(1, -1-2*s, -2*s, float(g['hmin']), H0, int(g['num_dofs']))
    assert o['qod'] == int(g['quad_order_diagonal'])
    assert o['b_qod'] == int(g['boundary_quad_order_diagonal'])
    assert o['target_order'] == float(g['target_order'])
    R = tables.near_rules(1, -1-2*s, -2*s, o)
    for nm, key in (('qrId', ('interior', -2)), ('qrVertex', ('interior', -1)), ('bqrVertex', ('boundary', -1))):
        b, w = R[key]
        assert b.shape == g[nm+'_nodes'].shape
        assert np.abs(b-g[nm+'_nodes']).max() < 1e-15
        assert np.allclose(w, g[nm+'_weights'], rtol=1e-14, atol=0)


def test_scaling_constants(golden_dir):
    g = load(golden_dir, 'kernel_values')
    for dim in (1, 2):
        for s in (0.25, 0.75):
            C = tables.fractional_scaling(dim, s)
            assert np.isclose(C, float(g['C_%dd_s%g' % (dim, s)]), rtol=1e-15)
            assert np.isclose(C/s, float(g['Cb_%dd_s%g' % (dim, s)]), rtol=1e-15)
            x, y = g['x_%dd' % dim], g['y_%dd' % dim]
            d2 = ((x-y)**2).sum(axis=1)
            assert np.allclose(C*d2**(-dim/2.-s), g['k_%dd_s%g' % (dim, s)], rtol=1e-14)
            assert np.allclose(C/s*d2**(-(dim-1)/2.-s), g['kb_%dd_s%g' % (dim, s)], rtol=1e-14)


@pytest.mark.parametrize('name', CASES_2D+CASES_1D)
def test_pairs_and_dense_match_reference(golden_dir, name):
    g = load(golden_dir, name)
    P = problem_from_golden(g)
    panels, p1, p2, C = P.pairs(g['pairs'])
    # bit-exact classification, panel order and permutations
    assert np.array_equal(panels, g['panels'])
    touching = panels < 0
    assert np.array_equal(p1[touching], g['perm1'][touching])
    assert np.array_equal(p2[touching], g['perm2'][touching])
    # local matrices: same algorithm, same summation order
    scale = np.abs(g['contribs']).max(axis=1, keepdims=True)
    assert (np.abs(C-g['contribs'])/scale).max() < 1e-14
    bpan, bC = P.boundary_pairs(g['bpairs'])
    assert np.array_equal(bpan, g['bpanels'])
    scale = np.abs(g['bcontribs']).max(axis=1, keepdims=True)
    assert (np.abs(bC-g['bcontribs'])/scale).max() < 1e-14
    # assembled matrices (thread-private partial sums: order differs)
    A = P.dense(True)
    A0 = P.dense(False)
    nz = np.abs(g['A']) > 0
    assert (np.abs(A-g['A'])[nz]/np.abs(g['A'])[nz]).max() < 1e-12
    assert np.abs(A0-g['A_interior']).max()/np.abs(g['A_interior']).max() < 1e-13


def test_all_pair_panels_match_reference(golden_dir):
    """every pair c1<=c2 of the r=2 disc and the r=6 interval"""
    for name in ('disc_s0.75_r2', 'interval_s0.25_r6'):
        g = load(golden_dir, name)
        P = problem_from_golden(g)
        nc = g['cells'].shape[0]
        iu = np.triu_indices(nc)
        pairs = np.stack(iu, axis=1).astype(np.int32)
        panels = P.pairs(pairs, with_contrib=False)[0]
        assert np.array_equal(panels, g['panel_matrix'][iu])


def test_slices_add_up():
    """rank slices of the cell loop sum to the full matrix (the reference's Allreduce)"""
    P = oracle.disc_problem(2)
    A = P.dense(True)
    nc = P.P.nc
    B = sum(P.dense(True, start=a, end=b) for a, b in ((0, 30), (30, 61), (61, nc)))
    assert np.abs(A-B).max()/np.abs(A).max() < 1e-14


@pytest.mark.parametrize('name', ['h2_disc_s0.75_r4', 'h2_interval_s0.25_r8'])
def test_farfield_blocks_match_reference(golden_dir, name):
    from oracle import h2
    g = load(golden_dir, name)
    dim = g['vertices'].shape[1]
    ptr = g['far_ptr']
    for k in range(g['far_m1'].shape[0]):
        blk = h2.farfield_block(dim, float(g['s']), g['far_box1'][k], g['far_box2'][k], g['far_m1'][k], g['far_m2'][k])
        ref = g['far_blocks'][ptr[k]:ptr[k+1]].reshape(blk.shape)
        assert (np.abs(blk-ref)/np.abs(ref)).max() < 1e-14


def test_sampled_rows_at_2977_dofs(golden_dir):
    """disc, 5 refinements (the size of the reference's driver test): sampled rows, the diagonal and products
    with seeded vectors of the reference's getDense output."""
    g = load(golden_dir, 'disc_s0.75_r5_rows')
    P = oracle.Problem(g['vertices'], g['cells'], g['dofs'], int(g['num_dofs']), float(g['s']),
                       bfacets=g['boundaryEdges'], target_order=0.5, hVector=g['hVector'], volVector=g['volVector'],
                       hmin=float(g['hmin']), diam=float(g['diam']))
    A = P.dense(True)
    d = np.sqrt(g['diagonal'])
    scale = np.maximum(np.abs(g['A_rows']), 1e-2*np.outer(d[g['rows']], d))
    assert (np.abs(A[g['rows']]-g['A_rows'])/scale).max() < 1e-12
    assert np.abs(np.diag(A)/g['diagonal']-1).max() < 1e-13
    assert np.abs(A.dot(g['x'])-g['Ax']).max() < 1e-12*np.abs(g['Ax']).max()
    assert np.abs(A.dot(np.ones(A.shape[0]))-g['ones_Ax']).max() < 1e-12*np.abs(g['diagonal']).max()
    assert abs(np.linalg.norm(A)/float(g['frobenius'])-1) < 1e-13


@pytest.mark.parametrize('name', ['disc_dm2_s0.75_r2', 'disc_dm2_s0.25_r3'])
def test_two_dofmaps_match_reference(golden_dir, name):
    """rows: interior DoFs, columns: the complementary (boundary) DoFs; the reference assembles over the combined
    map and keeps that block (nonlocalAssembly_{SCALAR}.pxi:1366-1378)"""
    g = load(golden_dir, name)
    n1, n2 = int(g['num_dofs']), int(g['num_dofs2'])
    combined = np.where(g['dofs'] >= 0, g['dofs'], n1+g['dofs2'])
    P = oracle.Problem(g['vertices'], g['cells'], combined, n1+n2, float(g['s']), bfacets=g['boundaryEdges'], target_order=0.5,
                       hVector=g['hVector'], volVector=g['volVector'], hmin=float(g['hmin']), diam=float(g['diam']),
                       order_num_dofs=n1)
    for ze, key in ((True, 'A_bc'), (False, 'A_bc_interior')):
        A = P.dense(ze)[:n1, n1:]
        assert A.shape == g[key].shape
        assert np.abs(A-g[key]).max() < 1e-13*np.abs(g[key]).max()


@pytest.mark.parametrize('name', ['disc_leftright_r2', 'disc_leftright_r3'])
def test_piecewise_variable_order_matches_reference(golden_dir, name):
    """leftRightFractionalOrder (fractionalOrders.pyx:285-335): the kernel parameters are set per cell pair from the cell
    centres (NO.pxi:509-513).  Restated as a sum over the classes of label pairs, each with its constant order; the
    singular quadrature orders follow s.max (fractionalLaplacian2D.pyx:606-611)."""
    g = load(golden_dir, name)
    svals = [float(g['sll']), float(g['slr']), float(g['srr'])]
    labels = (g['vertices'][g['cells']].mean(axis=1)[:, 0] >= float(g['interface'])).astype(np.uint8)
    blabels = (g['vertices'][g['boundaryEdges']].mean(axis=1)[:, 0] >= float(g['interface'])).astype(np.uint8)
    pc = np.zeros((4, 4), dtype=np.uint8)
    pc[0, 1] = pc[1, 0] = 1
    pc[1, 1] = 2
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        A = 0.
        for k, s in enumerate(svals):
            P = oracle.Problem(g['vertices'], g['cells'], g['dofs'], int(g['num_dofs']), s, bfacets=g['boundaryEdges'],
                               target_order=0.5, hVector=g['hVector'], volVector=g['volVector'], hmin=float(g['hmin']),
                               diam=float(g['diam']), s_max=max(svals), labels=labels, blabels=blabels, pair_class=pc,
                               active_class=k, max_order=40)
            assert P.orders['qod'] == int(g['quad_order_diagonal']) and P.orders['qodV'] == int(g['quad_order_diagonalV'])
            A = A+P.dense(ze)
        assert np.abs(A-g[key]).max() < 1e-13*np.abs(g[key]).max()


@pytest.mark.parametrize('name', ['entry_disc_s0.75_r3', 'entry_interval_s0.25_r6'])
def test_single_entries_match_reference(golden_dir, name):
    """getEntry / getDiagonal (nonlocalAssembly_{SCALAR}.pxi:1539-1660, 2269-2289): the form over the patch of the two
    basis functions plus the surface integral around the patch.  Restated as the dense operator of the patch sub-mesh
    with the quadrature parameters (hmin, diam, DoF count) of the whole problem."""
    g = load(golden_dir, name)
    dofs, N = g['dofs'], int(g['num_dofs'])

    def entry(I, J):
        patch = np.where(((dofs == I) | (dofs == J)).any(axis=1))[0]
        sub = np.where(dofs[patch] == I, 0, np.where(dofs[patch] == J, 1, -1))
        P = oracle.Problem(g['vertices'], g['cells'][patch], sub, 2, float(g['s']), target_order=float(g['target_order']),
                           hVector=g['hVector'][patch], volVector=g['volVector'][patch], hmin=float(g['hmin']),
                           diam=float(g['diam']), order_num_dofs=N)
        return P.dense(True)[0, 0 if I == J else 1]
    scale = np.abs(g['diagonal']).max()
    for (I, J), ref in zip(g['IJ'], g['entries']):
        assert abs(entry(I, J)-ref) < 1e-12*max(abs(ref), 1e-2*scale)
    for I in range(0, N, max(1, N//8)):
        assert abs(entry(I, I)/g['diagonal'][I]-1) < 1e-12


def test_oracle_mesh_sizes_bit_exact_vs_reference(golden_dir):
    """the oracle's own hVector / h / hmin (orc_edge_lengths) against every fixture that stores the reference's"""
    import glob
    from oracle import meshes
    n = 0
    for f in sorted(glob.glob(os.path.join(golden_dir, '*.npz'))):
        g = np.load(f)
        if 'hVector' not in g.files or 'cells' not in g.files:
            continue
        m = meshes.Mesh(g['vertices'], g['cells'])
        assert np.array_equal(m.hVector, g['hVector']), f
        assert m.hmin == float(g['hmin']) and m.h == float(g['h']), f
        assert np.array_equal(m.volVector, g['volVector']), f
        n += 1
    assert n >= 10


NONSYM = ['nonsym_disc_leftright_r2', 'nonsym_disc_leftright_r3', 'nonsym_interval_leftright_r5', 'nonsym_disc_layers_r3',
          'disc_layers_sym_r2', 'nonsym_interval_innerouter_r5', 'nonsym_disc_constant0.75_r2', 'nonsym_disc_constant0.25_r3',
          'nonsym_interval_constant0.75_r5']


def order_from_fixture(g):
    import pynucleus_b200 as pb
    kind = str(g['kind'])
    dim = g['vertices'].shape[1]
    if kind == 'leftRight':
        return pb.leftRightFractionalOrder(float(g['sll']), float(g['srr']), float(g['slr']), float(g['srl']), float(g['interface']))
    if kind == 'layers':
        return pb.layersFractionalOrder(dim, g['layerBoundaries'], g['layerOrders'])
    if kind == 'innerOuter':
        return pb.innerOuterFractionalOrder(dim, float(g['sii']), float(g['soo']), float(g['r']), g['center'], float(g['sio']), float(g['soi']))
    if kind == 'constantNonSym':
        return pb.constantNonSymFractionalOrder(float(g['s']))
    raise NotImplementedError(kind)


@pytest.mark.parametrize('name', NONSYM)
def test_unsymmetric_piecewise_order_matches_reference(golden_dir, name):
    """Unsymmetric piecewise orders (fractionalLaplacian{1,2}D_nonsym, fractionalLaplacian2D.pyx:894-1184; both
    orientations of a pair visited, nonlocalAssembly_{SCALAR}.pxi:1412-1428).  With piecewise parameters the orientation
    (c1, c2) contributes the symmetric local matrix of s(c1, c2) at weight 1, with c1 as first cell of the singular rule:
    restated as two constant-order passes per order at half weight, over the orientations (smaller, larger cell index)
    and (larger, smaller) of that order."""
    g = load(golden_dir, name)
    sFun = order_from_fixture(g)
    dim = g['vertices'].shape[1]
    assert bool(g['symmetric']) == sFun.symmetric
    bf = g['boundaryEdges'] if dim == 2 else g['boundaryVertices'].reshape(-1, 1)
    labels = sFun.labels(g['vertices'][g['cells']].mean(axis=1))
    blabels = sFun.labels(g['vertices'][bf].mean(axis=1))
    sV = sFun.blockOrders()
    nb = sV.shape[0]
    vals = sorted(set(sV.ravel().tolist()))
    pc = np.array([[vals.index(sV[i, j]) for j in range(nb)] for i in range(nb)])
    kw = dict(bfacets=bf, hVector=g['hVector'], volVector=g['volVector'], hmin=float(g['hmin']), diam=float(g['diam']),
              s_max=max(vals), s_min=min(vals), labels=labels, blabels=blabels, active_class=1, max_order=40)
    if dim == 2:
        kw['target_order'] = 0.5

    def problem(k, M, orientation):
        P4, B4 = np.zeros((4, 4), dtype=np.uint8), np.zeros((4, 4), dtype=np.uint8)
        P4[:nb, :nb] = M
        B4[:nb, :nb] = pc == k
        return oracle.Problem(g['vertices'], g['cells'], g['dofs'], int(g['num_dofs']), vals[k], pair_class=P4, bpair_class=B4,
                              pair_orientation=orientation, **kw)

    for ze, key in ((True, 'A'), (False, 'A_interior')):
        A = 0.
        for k in range(len(vals)):
            if sFun.symmetric:
                A = A+problem(k, pc == k, 0).dense(ze)
            else:
                A = A+0.5*problem(k, pc == k, 0).dense(ze)+0.5*problem(k, (pc == k).T, 1).dense(ze)
        assert np.abs(A-g[key]).max() < 1e-13*np.abs(g[key]).max()


def _varorder_fun(g):
    from oracle import varorder
    cls = varorder.smoothStep if str(g['kind']) == 'smoothedLeftRight' else varorder.linearStep
    return cls(float(g['sl']), float(g['sr']), float(g['r']), float(g['interface']))


@pytest.mark.parametrize('name', ['varorder_interval_smoothed_r5', 'varorder_interval_linear_r5', 'varorder_interval_smoothed_r6',
                                  'varorder_disc_smoothed_r2', 'varorder_p2_interval_smoothed_r4', 'varorder_p2_disc_smoothed_r1',
                                  'varorder_p0_disc_smoothed_r2'])
def test_order_varying_inside_cells_matches_reference(golden_dir, name):
    """orders that vary inside a cell (kernel.piecewise == False; the driver's twoDomainNonSym): the numpy restatement
    oracle/varorder.py against operators assembled by the reference itself (make_golden_varorder.py)"""
    from oracle import varorder
    g = np.load(os.path.join(golden_dir, name+'.npz'))
    sF = _varorder_fun(g)
    dim = g['vertices'].shape[1]
    X, Y = g['points_x'], g['points_y']
    assert np.abs(sF(X)-g['s_values']).max() < 1e-15
    assert np.abs(varorder.kernel_value(dim, sF, X, Y)/g['kernel_values']-1).max() < 1e-14
    assert np.abs(varorder.kernel_value(dim, sF, X, Y, True)/g['bkernel_values']-1).max() < 1e-14
    bf = g['boundaryEdges'] if dim == 2 else g['boundaryVertices']
    for ze, key in ((False, 'A_interior'), (True, 'A')):
        A = varorder.dense(g['vertices'], g['cells'], g['dofs'], int(g['num_dofs']), sF, bf, zero_exterior=ze,
                           hmin=float(g['hmin']), diam=float(g['diam']))
        assert np.abs(A-g[key]).max() < 1e-12*np.abs(g[key]).max()


def test_order_varying_inside_cells_reproduces_cached_driver_run():
    """oracle/varorder.py against the reference's cached driver result
    tests/cache_runFractional.py--domaininterval--stwoDomainNonSym(0.25,0.75)--problemknownSolution--elementP1--solverlu--matrixFormatdense
    (127 DoFs, u = (1-x^2)^0.7): interpolated L2 and Linf errors of the discrete solution"""
    from scipy.special import hyp2f1, gamma as Gamma
    from oracle import varorder
    m = meshes.interval(-1., 1., 7)
    dofs, n = meshes.p1_dofs(m)
    assert n == 127
    sF = varorder.smoothStep(0.25, 0.75, 0.1, 0.)
    A = varorder.dense(m.vertices, m.cells, dofs, n, sF, m.boundary_facets())
    beta = 0.7
    bary, w = tables.regular_rule(3, 1)            # simplexXiaoGimbutas(3, 1), discretizedProblems.py:561
    T = m.vertices[m.cells][:, :, 0]
    pts = np.einsum('kq,ck->cq', bary, T)
    s = sF(pts[..., None])
    f = 2**(2*s)*Gamma(s+0.5)*Gamma(beta+1.)/np.sqrt(np.pi)/Gamma(beta+1.-s)*hyp2f1(s+0.5, -beta+s, 0.5, pts**2)
    vol = np.abs(T[:, 1]-T[:, 0])
    b, M, xs = np.zeros(n), np.zeros((n, n)), np.zeros(n)
    for k in range(2):
        ok = dofs[:, k] >= 0
        np.add.at(b, dofs[ok, k], (vol[:, None]*f*w*bary[k]).sum(axis=1)[ok])
        xs[dofs[ok, k]] = T[ok, k]
        for l in range(2):
            ok2 = ok & (dofs[:, l] >= 0)
            np.add.at(M, (dofs[ok2, k], dofs[ok2, l]), (2. if k == l else 1.)*vol[ok2]/6.)
    u = np.linalg.solve(A, b)
    e = u-(1-xs**2)**beta
    assert abs(np.abs(e).max()/0.003599161364716205-1) < 1e-8
    assert abs(np.sqrt(e.dot(M.dot(e)))/0.001265060713568335-1) < 1e-8


@pytest.mark.parametrize('name', ['tempered_interval_s0.75_l2_r5', 'tempered_interval_s0.25_l0.5_r6', 'tempered_disc_s0.75_l2_r2',
                                  'tempered_disc_s0.25_l1_r3'])
def test_tempered_kernel_matches_reference(golden_dir, name):
    """tempered fractional kernels in the C restatement against operators assembled by the reference (make_golden_tempered.py)"""
    g = np.load(os.path.join(golden_dir, name+'.npz'))
    dim = g['vertices'].shape[1]
    bf = g['boundaryEdges'] if dim == 2 else g['boundaryVertices'].reshape(-1, 1)
    P = oracle.Problem(g['vertices'], g['cells'], g['dofs'], int(g['num_dofs']), float(g['s']), bfacets=bf,
                       target_order=0.5 if dim == 2 else None, hVector=g['hVector'], volVector=g['volVector'],
                       hmin=float(g['hmin']), diam=float(g['diam']), tempered=float(g['tempered']))
    assert abs(P.C/float(g['scaling'])-1) < 1e-14 and abs(P.Cb/float(g['bscaling'])-1) < 1e-14
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        assert np.abs(P.dense(ze)-g[key]).max() < 1e-13*np.abs(g[key]).max()


def smooth_oracle_problem(g, **kw):
    """oracle problem of a Gaussian / exponential fixture (make_golden_smooth.py)"""
    from math import pi, sqrt
    dim = g['vertices'].shape[1]
    bf = g['boundaryEdges'] if dim == 2 else g['boundaryVertices'].reshape(-1, 1)
    C = float(g['scaling'])
    if str(g['kernelType']) == 'gaussian':
        a = 0.5/float(g['variance'])**dim           # fEXPONENTINVERSE, kernelsCy.pyx:693-695
        sm = (C, 2, a, C*sqrt(pi/a), 3, a) if dim == 1 else (C, 2, a, C/a, 4, a)
    else:
        a = float(g['exponentialRate'])
        sm = (C, 1, a, 2*C/a, 1, a)
    return oracle.Problem(g['vertices'], g['cells'], g['dofs'], int(g['num_dofs']), -0.5*dim, bfacets=bf,
                          target_order=0.5 if dim == 2 else None, hVector=g['hVector'], volVector=g['volVector'],
                          hmin=float(g['hmin']), diam=float(g['diam']), smooth=sm, **kw)


@pytest.mark.parametrize('name', ['gaussian_interval_v0.1_r5', 'gaussian_interval_v0.02_r6', 'exponential_interval_a8_r5',
                                  'exponential_interval_a2.5_r6', 'gaussian_disc_v0.1_r2', 'gaussian_disc_v0.3_r3'])
def test_gaussian_and_exponential_kernels_match_reference(golden_dir, name):
    """Gaussian / exponential kernels on the full space in the C restatement against operators assembled by the reference"""
    g = np.load(os.path.join(golden_dir, name+'.npz'))
    P = smooth_oracle_problem(g)
    assert P.orders['qod'] == int(g['quad_order_diagonal']) and P.orders['b_qod'] == int(g['bquad_order_diagonal'])
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        assert np.abs(P.dense(ze)-g[key]).max() < 1e-13*np.abs(g[key]).max()


# cached tests/cache_runNonlocal.py--domaininterval--kernelType{gaussian,exponential}--problem*--solverlu--matrixFormatH2--...
# --interactionfullSpace--horizoninf of the reference: (L2 error interpolated, Linf error interpolated); both errors run over
# ALL vertices of the mesh (the boundary vertices carry the full analytic value: the driver notes that its Dirichlet data are
# "not quite correct"), 511 unknowns
SMOOTH_DRIVER_CASES = {'gaussian': (0.0029565447289171816, 0.006737946999085467),
                       'exponential': (0.00025530396949181036, 0.00033546262790251185)}


def smooth_driver_setup(kt):
    """kernel parameters, forcing and analytic solution of the drivers' `gaussian` / `exponential` problems
    (nonlocalProblems.py:1254-1285) with the flags of tests/test_drivers_intFracLapl.py:70-73"""
    from math import pi, sqrt
    if kt == 'gaussian':
        var = 0.1
        C, a = 1/sqrt(2*pi*var)/2, 0.5/var
        return dict(variance=var), (C, 2, a, C*sqrt(pi/a), 3, a), \
            (lambda x: np.exp(-0.5*x**2/var)-np.exp(-0.25*x**2/var)/np.sqrt(2)), (lambda x: np.exp(-0.5*x**2/var))
    rate = 8.0
    C = rate**3/2/2
    return dict(exponentialRate=rate), (C, 1, rate, 2*C/rate, 1, rate), \
        (lambda x: np.exp(-rate*np.abs(x))*(1/rate-np.abs(x))*C*2.0), (lambda x: np.exp(-rate*np.abs(x)))


def smooth_driver_errors(vertices, cells, dofs, n, A, f, u_ex):
    """load vector with simplexXiaoGimbutas(3, 1), direct solve, interpolated L2 / Linf errors over all vertices"""
    bary, w = tables.regular_rule(3, 1)
    T = vertices[cells][:, :, 0]
    pts = np.einsum('kq,ck->cq', bary, T)
    vol = np.abs(T[:, 1]-T[:, 0])
    b = np.zeros(n)
    nv = vertices.shape[0]
    Mf = np.zeros((nv, nv))
    for k in range(2):
        ok = dofs[:, k] >= 0
        np.add.at(b, dofs[ok, k], (vol[:, None]*f(pts)*w*bary[k]).sum(axis=1)[ok])
        for l in range(2):
            np.add.at(Mf, (cells[:, k], cells[:, l]), (2. if k == l else 1.)*vol/6.)
    u = np.linalg.solve(A, b)
    uf = np.zeros(nv)
    for k in range(2):
        ok = dofs[:, k] >= 0
        uf[cells[ok, k]] = u[dofs[ok, k]]
    e = uf-u_ex(vertices[:, 0])
    return np.sqrt(e.dot(Mf.dot(e))), np.abs(e).max()


@pytest.mark.parametrize('kt', ['gaussian', 'exponential'])
def test_gaussian_and_exponential_kernels_reproduce_cached_driver_runs(kt):
    """the C oracle against the reference's cached runNonlocal results for the Gaussian and the exponential kernel on the
    full space (the cached runs use the H2 format: 1.5e-8 away from the dense operator)"""
    m = meshes.interval(-1., 1., 9)
    dofs, n = meshes.p1_dofs(m)
    assert n == 511
    _, sm, f, u_ex = smooth_driver_setup(kt)
    A = oracle.Problem(m.vertices, m.cells, dofs, n, -0.5, smooth=sm).dense(True)
    L2i, Linf = smooth_driver_errors(m.vertices, m.cells, dofs, n, A, f, u_ex)
    assert abs(L2i/SMOOTH_DRIVER_CASES[kt][0]-1) < 1e-6 and abs(Linf/SMOOTH_DRIVER_CASES[kt][1]-1) < 1e-12


def fe_order_vertex_values(g):
    """vertex values of the P1 order function stored in a varorder_fe_* fixture"""
    vs = np.zeros(g['vertices'].shape[0])
    od = g['order_dofs']
    for k in range(od.shape[1]):
        ok = od[:, k] >= 0
        vs[g['cells'][ok, k]] = g['order_values'][od[ok, k]]
    return vs


@pytest.mark.parametrize('name', ['varorder_fe_interval_r5', 'varorder_fe_disc_r2'])
def test_order_given_by_a_fe_function_matches_reference(golden_dir, name):
    """feFractionalOrder (the order is a P1 function on the mesh of the operator) in oracle/varorder.py against operators
    assembled by the reference itself (make_golden_varorder.py fe)"""
    from oracle import varorder
    g = np.load(os.path.join(golden_dir, name+'.npz'))
    dim = g['vertices'].shape[1]
    sF = varorder.feOrder(fe_order_vertex_values(g), float(g['smin']), float(g['smax']))
    bf = g['boundaryEdges'] if dim == 2 else g['boundaryVertices']
    for ze, key in ((True, 'A'), (False, 'A_interior')):
        A = varorder.dense(g['vertices'], g['cells'], g['dofs'], int(g['num_dofs']), sF, bf, zero_exterior=ze,
                           hmin=float(g['hmin']), diam=float(g['diam']))
        assert np.abs(A-g[key]).max() < 1e-12*np.abs(g[key]).max()
