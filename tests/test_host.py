"""CPU-side tests: C ABI loads and exports what include/pnb200.h declares, the
host-side mirror (tables, meshes, DoFMap, kernels) agrees with the oracle and
with the reference goldens, and the product fails loudly without a GPU."""
import ctypes
import os
import re

import numpy as np
import pytest

import pynucleus_b200 as pb
from pynucleus_b200 import _lib, quadrature
from oracle import tables, meshes

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))


@pytest.fixture(scope='module')
def lib():
    import __graft_entry__ as g
    g.build()
    return _lib.lib()


def test_abi_exports_every_declared_symbol(lib):
    header = open(os.path.join(ROOT, 'include', 'pnb200.h')).read()
    declared = set(re.findall(r'\b(pnb_[a-z0-9_]+)\s*\(', header))
    assert declared == set(_lib.EXPORTS)
    for name in declared:
        assert hasattr(lib, name)
    assert lib.pnb_version() >= 100
    assert lib.pnb_far_max_order() == 5


def test_no_cpu_fallback(lib):
    """without a CUDA device every compute entry point fails with PNB_ERR_NO_DEVICE"""
    import torch
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    mesh = pb.refined(pb.uniform_disc(), 1)
    dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
    with pytest.raises(RuntimeError):
        b.getDense()
    y = ctypes.c_double()
    assert lib.pnb_fp64_peak(0, ctypes.byref(y)) == -1
    assert b'no CUDA device' in lib.pnb_last_error()


def test_product_does_not_import_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, 'pynucleus_b200')):
        for f in files:
            if f.endswith(('.py', '.cu', '.cuh', '.h')):
                src = open(os.path.join(dirpath, f)).read()
                assert not re.search(r'^\s*(from|import)\s+oracle', src, re.M), f
                assert 'liboracle' not in src


@pytest.mark.parametrize('dim,s,hmin,diam,N', [(2, 0.75, 0.05, 2.8, 3000), (2, 0.25, 0.1, 2.8, 300),
                                               (1, 0.25, 0.03, 2., 63), (1, 0.75, 0.06, 2., 31)])
def test_tables_match_oracle(dim, s, hmin, diam, N):
    H0 = diam/np.sqrt(8)
    sing, bs = -dim-2*s, 1-dim-2*s
    to = 0.5 if dim == 2 else None
    o = quadrature.localMatrixOrders(dim, sing, bs, hmin, H0, N, to)
    oo = tables.diag_orders(dim, sing, bs, hmin, H0, N, to)
    assert (o.quad_order_diagonal, o.bquad_order_diagonal, o.target_order, o.btarget_order) == \
        (oo['qod'], oo['b_qod'], oo['target_order'], oo['b_target_order'])
    T = quadrature.singular_tables(dim, sing, bs, o)
    R = tables.near_rules(dim, sing, bs, oo)
    names = {'identical': ('interior', -3 if dim == 2 else -2), 'edge': ('interior', -2), 'vertex': ('interior', -1),
             'bedge': ('boundary', -2), 'bvertex': ('boundary', -1)}
    for k, (b, w) in T.items():
        bb, ww = R[names[k]]
        assert np.array_equal(b, bb)
        assert np.allclose(w, ww, rtol=1e-15, atol=0)


def test_regular_rules_match_oracle_and_are_exact():
    from oracle.triangle_rules import check_exactness
    for p in range(1, 25):
        for md in (0, 1, 2):
            b, w = quadrature.regular(p, md)
            bb, ww = tables.regular_rule(p, md)
            assert np.array_equal(b, bb) and np.array_equal(w, ww)
        b, w = quadrature.regular(p, 2)
        assert check_exactness(b, w, p) < 5e-15
        assert w.min() > 0 and b.min() > 0 and abs(w.sum()-1) < 1e-14


def test_meshes_match_reference(golden_dir):
    for r in (0, 4):
        g = np.load(os.path.join(golden_dir, 'disc_mesh_r%d.npz' % r))
        m = pb.refined(pb.uniform_disc(), r)
        assert np.array_equal(m.cells, g['cells'])
        assert np.abs(m.vertices-g['vertices']).max() < 1e-15
        assert np.array_equal(m.boundaryFacets, g['boundaryEdges'])
        dm = pb.P1_DoFMap(m)
        assert dm.num_dofs == int(g['num_dofs'])
        assert np.array_equal(np.where(dm.dofs >= 0, dm.dofs, -1), np.where(g['dofs'] >= 0, g['dofs'], -1))
        assert np.allclose(m.hVector, g['hVector'], rtol=1e-14)
        assert np.allclose(m.volVector, g['volVector'], rtol=1e-13)
        assert abs(m.diam-float(g['diam'])) < 1e-15
    g = np.load(os.path.join(golden_dir, 'interval_s0.25_r6.npz'))
    m = pb.refined(pb.simpleInterval(-1., 1.), 6)
    assert np.array_equal(m.cells, g['cells']) and np.array_equal(m.vertices, g['vertices'])
    assert pb.P1_DoFMap(m).num_dofs == 63


def test_dof_counts_of_the_bench_meshes():
    # N = 1 + nc/2 - 3*2^r
    for r, N in ((5, 2977), (6, 12097)):
        m = pb.refined(pb.uniform_disc(), r)
        assert m.num_cells == 6*4**r
        assert pb.P1_DoFMap(m).num_dofs == N


def test_kernel_values_match_reference(golden_dir):
    g = np.load(os.path.join(golden_dir, 'kernel_values.npz'))
    for dim in (1, 2):
        for s in (0.25, 0.75):
            k = pb.getFractionalKernel(dim, s)
            kb = k.getBoundaryKernel()
            assert np.isclose(k.scalingValue, float(g['C_%dd_s%g' % (dim, s)]), rtol=1e-15)
            assert np.isclose(kb.scalingValue, float(g['Cb_%dd_s%g' % (dim, s)]), rtol=1e-15)
            assert k.singularityValue == -dim-2*s and kb.singularityValue == 1-dim-2*s
            x, y = g['x_%dd' % dim], g['y_%dd' % dim]
            assert np.allclose([k(a, b) for a, b in zip(x, y)], g['k_%dd_s%g' % (dim, s)], rtol=1e-14)
            assert np.allclose([kb(a, b) for a, b in zip(x, y)], g['kb_%dd_s%g' % (dim, s)], rtol=1e-14)


def test_builder_argument_checks():
    mesh = pb.refined(pb.uniform_disc(), 1)
    dm = pb.P1_DoFMap(mesh)
    with pytest.raises(AssertionError):
        pb.nonlocalBuilder(dm, pb.getFractionalKernel(1, 0.75), {})
    with pytest.raises(AssertionError):
        pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'quadType': 'general'})
    b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2, 0.75), {'target_order': 0.5})
    assert b.zeroExterior and b.orders.quad_order_diagonal >= 4


def test_complement_and_combined_dofmaps(golden_dir):
    """getComplementDoFMap / combine (fem/PyNucleus_fem/DoFMaps.pyx:1170-1184, 1563-1588) against the reference's arrays"""
    import pynucleus_b200 as pb
    g = np.load(os.path.join(golden_dir, 'disc_dm2_s0.75_r2.npz'))
    # boundary vertices in the reference's order (= order of the negative DoF numbers)
    neg = g['dofs'] < 0
    bv = np.empty(int(g['num_dofs2']), dtype=np.int32)
    bv[-g['dofs'][neg]-1] = g['cells'][neg]
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'], boundaryVertices=bv)
    dm = pb.P1_DoFMap(mesh)
    assert np.array_equal(dm.dofs, g['dofs'])
    dm2 = dm.getComplementDoFMap()
    assert np.array_equal(dm2.dofs, g['dofs2']) and dm2.num_dofs == int(g['num_dofs2'])
    dmc = dm.combine(dm2)
    assert dmc.num_dofs == dm.num_dofs+dm2.num_dofs and dmc.dofs.min() == 0
    assert np.array_equal(np.sort(np.unique(dmc.dofs)), np.arange(dmc.num_dofs))
    assert np.array_equal(dmc.dofs[dm.dofs >= 0], dm.dofs[dm.dofs >= 0])


@pytest.mark.parametrize('name', ['h2_interval_s0.25_r8', 'h2_disc_s0.75_r4'])
def test_cluster_tree_and_farfield_chain_match_reference(golden_dir, name):
    """H2 structure node for node, and the far-field part of the reference's H2 matvec
    (leaf moments, transfer operators, three passes): H x - Anear x of the reference's getH2"""
    import scipy.sparse as sp
    import torch
    from pynucleus_b200 import cluster_tree as ct, h2
    from oracle import h2 as oracle_h2
    g = np.load(os.path.join(golden_dir, name+'.npz'))
    dim = g['vertices'].shape[1]
    bf = g['boundaryEdges'] if dim == 2 else g['boundaryVertices'].reshape(-1, 1)
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=bf)
    dm = pb.P1_DoFMap(mesh)
    kernel = pb.getFractionalKernel(dim, float(g['s']))
    b = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5} if dim == 2 else {})
    root = ct.build_tree(mesh, dm, kernel, b.orders.target_order)
    Pnear, Pfar = ct.admissible_clusters(root)
    nodes = list(root.get_tree_nodes())
    assert np.array_equal([n.id for n in nodes], g['tree_ids'])
    assert np.array_equal(np.array([n.box for n in nodes]), g['tree_boxes'])
    assert np.array_equal([n.parent.id if n.parent else -1 for n in nodes], g['tree_parent'])
    assert np.array_equal([n.isLeaf for n in nodes], g['tree_isleaf'])
    assert np.array_equal([n.levelNo for n in nodes], g['tree_level'])
    assert np.array_equal([n.interpolation_order for n in nodes], g['tree_order'])
    assert np.array_equal(np.concatenate([np.sort(n.dofs) for n in nodes]), g['tree_dofs'])
    assert np.array_equal(np.array([(a.id, c.id) for a, c in Pnear]), g['near_pairs'])
    assert np.array_equal(np.array([(lvl, a.id, c.id) for lvl in sorted(Pfar) for a, c in Pfar[lvl]]), g['far_pairs'])
    # far-field chain on the CPU (torch), kernel blocks from the oracle
    for n in nodes:
        if n.isLeaf:
            n.value = h2.leaf_values(n, mesh, dm)
        if n.parent is not None:
            n.transferOperator = h2.transfer_operator(n.parent, n)
    P = {lvl: [h2.farFieldClusterPair(a, c, oracle_h2.farfield_block(dim, float(g['s']), a.box, c.box, a.interpolation_order,
                                                                      c.interpolation_order)) for a, c in Pfar[lvl]]
         for lvl in sorted(Pfar)}
    H = h2.H2Matrix(root, P, None, dm.num_dofs, 'cpu')
    yfar = H.farfield_device(torch.as_tensor(g['x'])).numpy()
    N = dm.num_dofs
    low = sp.csr_matrix((g['Anear_data'], g['Anear_indices'], g['Anear_indptr']), shape=(N, N))   # SSS: strict lower + diagonal
    ref = g['Hx']-(low+low.T+sp.diags(g['Anear_diagonal'])).dot(g['x'])
    assert np.abs(yfar-ref).max() < 1e-11*np.abs(ref).max()


def test_restriction_operators_match_reference(golden_dir):
    """P1 restriction / prolongation of the uniformly refined disc meshes against the hierarchy of the reference's
    driver (multilevelSolver/PyNucleus_multilevelSolver/restriction_2D_P1.pxi; P = R^T), and the 1D operator against
    linear interpolation"""
    import pynucleus_b200 as pb
    g = np.load(os.path.join(golden_dir, 'mg_disc_varconst0.75_r4.npz'))
    mesh = pb.uniform_disc()
    dms = [pb.P1_DoFMap(mesh)]
    for k in range(1, int(g['num_levels'])):
        mesh = mesh.refine()
        dms.append(pb.P1_DoFMap(mesh))
        R, P = pb.buildRestrictionProlongation(dms[-2], dms[-1])
        assert R.shape == (dms[-2].num_dofs, dms[-1].num_dofs) == (g['level_num_dofs'][k-1], g['level_num_dofs'][k])
        assert np.array_equal(R.indptr, g['R%d_indptr' % k]) and np.array_equal(R.indices, g['R%d_indices' % k])
        assert np.array_equal(R.data, g['R%d_data' % k])
        assert (P != R.T).nnz == 0
    m0 = pb.refined(pb.simpleInterval(-1, 1), 3)
    m1 = m0.refine()
    d0, d1 = pb.P1_DoFMap(m0), pb.P1_DoFMap(m1)
    R, P = pb.buildRestrictionProlongation(d0, d1)
    x0, x1 = d0.getDoFCoordinates()[:, 0], d1.getDoFCoordinates()[:, 0]
    f = 0.3*x0+0.1            # linear functions vanishing nowhere special: interpolation is exact away from the boundary
    inner = np.abs(x1) < 1-2*m0.h
    assert np.abs(P.dot(f)-(0.3*x1+0.1))[inner].max() < 1e-14


def test_near_field_container_is_consistent():
    """nearFieldBlocks (h2.py): block form, compiled CSR form, dense form, diagonal and the scattered correction of
    the regional operator agree (CPU tensors; the blocks themselves come from the CUDA path)"""
    import torch
    from pynucleus_b200.h2 import nearFieldBlocks
    rng = np.random.default_rng(1)
    n = 23
    near = nearFieldBlocks(n, torch.device('cpu'))
    ref = np.zeros((n, n))
    for rows, cols in ((np.arange(0, 7), np.arange(0, 7)), (np.arange(0, 7), np.arange(7, 12)), (np.arange(7, 12), np.arange(0, 7)),
                       (np.arange(7, 23), np.arange(7, 23))):
        B = rng.standard_normal((rows.shape[0], cols.shape[0]))
        near.add(rows, cols, torch.from_numpy(B))
        ref[np.ix_(rows, cols)] += B
    cr, cc = np.array([0, 3, 3, 22]), np.array([0, 4, 3, 21])
    cv = rng.standard_normal(4)
    near.correction = (cr, cc, cv)
    np.add.at(ref, (cr, cc), cv)
    x = torch.from_numpy(rng.standard_normal(n))
    y_blocks = near.matvec_device(x).numpy()
    assert np.abs(near.toarray()-ref).max() < 1e-15
    near.compile()
    assert np.abs(near.matvec_device(x).numpy()-ref.dot(x.numpy())).max() < 1e-13
    assert np.abs(y_blocks-ref.dot(x.numpy())).max() < 1e-13
    assert np.abs(near.diagonal_device().numpy()-np.diag(ref)).max() < 1e-15


def test_mesh_sizes_bit_exact_vs_reference(golden_dir):
    """hVector / h / hmin from the reference's own vertex arrays must equal the reference's values to the last bit
    (hdeltaCy, meshCy.pyx:1654-1732: longest edge per cell, hmin = SHORTEST edge of the mesh, edge lengths through a
    fused dot product): getQuadOrder and the singular quadrature orders round logarithms of these numbers up"""
    import glob
    n = 0
    for f in sorted(glob.glob(os.path.join(golden_dir, '*.npz'))):
        g = np.load(f)
        if 'hVector' not in g.files or 'cells' not in g.files:
            continue
        m = pb.meshNd(g['vertices'], g['cells'])
        assert np.array_equal(m.hVector, g['hVector']), f
        assert m.hmin == float(g['hmin']) and m.h == float(g['h']), f
        assert np.array_equal(m.volVector, g['volVector']), f
        n += 1
    assert n >= 10


def test_cython_shim_builds_and_binds_the_c_abi():
    """integration/pnb200_shim.pyx compiles against include/pnb200.h, links libpnb200.so and fails loudly without a
    device (no compute here)"""
    import subprocess
    import sys
    root = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
    sys.path.insert(0, root)
    from integration.build_shim import build
    so = build()
    assert os.path.exists(so)
    out = subprocess.run(['nm', '-D', '--undefined-only', so], capture_output=True, text=True).stdout
    for sym in ('pnb_problem_create', 'pnb_dense_assemble', 'pnb_problem_destroy', 'pnb_max_order', 'pnb_last_error'):
        assert sym in out


@pytest.mark.parametrize('name', ['p2_interval_s0.25_r4', 'p2_disc_s0.75_r1', 'p2_disc_s0.75_r2', 'p2_disc_s0.75_r3'])
def test_p2_dofmap_numbering_matches_reference(name):
    """P2_DoFMap (fem/PyNucleus_fem/DoFMaps.pyx:157-322, 1978-2031): cell -> dof table equal to the reference's"""
    import pynucleus_b200 as pb
    g = dict(np.load(os.path.join(os.path.dirname(__file__), 'golden', name+'.npz')))
    dim = g['vertices'].shape[1]
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'] if dim == 2 else g['boundaryVertices'],
                     boundaryVertices=g['boundaryVertices'])
    dm = pb.P2_DoFMap(mesh)
    assert dm.num_dofs == int(g['num_dofs']) and dm.num_boundary_dofs == int(g['num_boundary_dofs'])
    assert np.array_equal(dm.dofs, g['dofs'])
    assert np.array_equal(dm.vertexPart().dofs, g['dofs'][:, :dim+1])


@pytest.mark.parametrize('name', ['varorder_interval_smoothed_r6', 'varorder_interval_linear_r5', 'varorder_disc_smoothed_r3'])
def test_orders_varying_inside_cells_host_side(golden_dir, name):
    """host side of the path for orders s(x,y) = sFun(x) (no GPU needed): the order function and the kernel objects against
    values of the reference's own objects, the quadrature orders against the reference's local matrices, the ranking of the
    pair singularities against the oracle's evalParamsOnSimplices restatement"""
    from oracle import varorder
    g = np.load(os.path.join(golden_dir, name+'.npz'))
    dim = g['vertices'].shape[1]
    cls, ocls = ((pb.smoothedLeftRightFractionalOrder, varorder.smoothStep) if str(g['kind']) == 'smoothedLeftRight'
                 else (pb.linearLeftRightFractionalOrder, varorder.linearStep))
    order = cls(float(g['sl']), float(g['sr']), r=float(g['r']), interface=float(g['interface']))
    sF = ocls(float(g['sl']), float(g['sr']), float(g['r']), float(g['interface']))
    X, Y = g['points_x'], g['points_y']
    assert np.abs(order.evaluate(X)-g['s_values']).max() < 1e-15
    kernel = pb.getFractionalKernel(dim, order)
    assert not kernel.piecewise and not kernel.symmetric and kernel.variable
    kb = kernel.getBoundaryKernel()
    for i in range(X.shape[0]):
        assert abs(kernel(X[i], Y[i])/g['kernel_values'][i]-1) < 1e-14
        assert abs(kb(X[i], Y[i])/g['bkernel_values'][i]-1) < 1e-14
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'] if dim == 2 else g['boundaryVertices'])
    dm = pb.P1_DoFMap(mesh)
    b = pb.nonlocalBuilder(dm, kernel, {'target_order': 0.5} if dim == 2 else {})
    assert b.orders.target_order == float(g['target_order_used']) and b.orders.btarget_order == float(g['btarget_order_used'])
    assert b.orders.quad_order_diagonal == int(g['quad_order_diagonal'])
    assert b.orders.bquad_order_diagonal == int(g['bquad_order_diagonal'])
    V = b._varorder
    T = mesh.vertices[mesh.cells]
    smax = np.maximum(sF(T.mean(axis=1)), sF(T).max(axis=1))
    assert np.abs(V['values'][V['cell_value']]-smax).max() < 1e-15
    assert (np.diff(V['values']) > 0).all() and V['values'][0] >= order.min and V['values'][-1] <= order.max
    st = b._varorder_struct()
    assert st.num_values == V['values'].shape[0] and st.identical[0].n > 0 and st.bvertex[st.num_values-1].n > 0
    # nothing but getDense() may use the stand-in device problem of this path
    with pytest.raises(NotImplementedError):
        b.getEntry(0, 0)


@pytest.mark.parametrize('name', ['tempered_interval_s0.75_l2_r5', 'tempered_disc_s0.25_l1_r3'])
def test_tempered_kernel_objects_match_reference(golden_dir, name):
    """kernel objects of the tempered fractional kernels against values of the reference's own objects: interior kernel with
    the exponential factor, boundary kernel without it (getBoundaryKernel does not hand `tempered` on)"""
    g = np.load(os.path.join(golden_dir, name+'.npz'))
    dim = g['vertices'].shape[1]
    kernel = pb.getFractionalKernel(dim, float(g['s']), tempered=float(g['tempered']))
    kb = kernel.getBoundaryKernel()
    assert kernel.tempered == float(g['tempered']) and kb.tempered == 0.
    assert abs(kernel.scalingValue/float(g['scaling'])-1) < 1e-14 and abs(kb.scalingValue/float(g['bscaling'])-1) < 1e-14
    X, Y = g['points_x'], g['points_y']
    for i in range(X.shape[0]):
        assert abs(kernel(X[i], Y[i])/g['kernel_values'][i]-1) < 1e-14
        assert abs(kb(X[i], Y[i])/g['bkernel_values'][i]-1) < 1e-14
    with pytest.raises(NotImplementedError):
        pb.getFractionalKernel(dim, float(g['s']), horizon=0.3, tempered=1.)


SMOOTH_CASES = ['gaussian_interval_v0.1_r5', 'gaussian_interval_v0.02_r6', 'exponential_interval_a8_r5', 'exponential_interval_a2.5_r6',
                'gaussian_disc_v0.1_r2', 'gaussian_disc_v0.3_r3']


def _smooth_kernel(g):
    dim = g['vertices'].shape[1]
    kw = {'variance': float(g['variance'])} if str(g['kernelType']) == 'gaussian' else {'exponentialRate': float(g['exponentialRate'])}
    return pb.getIntegrableKernel(dim, str(g['kernelType']), np.inf, interaction='fullSpace', **kw)


@pytest.mark.parametrize('name', SMOOTH_CASES)
def test_gaussian_and_exponential_kernel_objects_match_reference(golden_dir, name):
    """Gaussian / exponential kernels on the full space (the reference's driver tests `--interaction fullSpace`): kernel and
    boundary-kernel values, scaling constants and the quadrature orders of the local matrices against the reference's own
    objects (make_golden_smooth.py); the smooth-factor description handed to the library reproduces both kernels"""
    from scipy.special import erfc
    g = np.load(os.path.join(golden_dir, name+'.npz'))
    dim = g['vertices'].shape[1]
    kernel = _smooth_kernel(g)
    kb = kernel.getBoundaryKernel()
    assert abs(kernel.scalingValue/float(g['scaling'])-1) < 1e-14 and abs(kb.scalingValue/float(g['bscaling'])-1) < 1e-14
    X, Y = g['points_x'], g['points_y']
    mode, a, bmode, ba, bconst = kernel.smoothFactors()
    f = {1: lambda a_, d2: np.exp(-a_*np.sqrt(d2)), 2: lambda a_, d2: np.exp(-a_*d2), 3: lambda a_, d2: erfc(np.sqrt(a_*d2)),
         4: lambda a_, d2: np.exp(-a_*d2)/np.sqrt(d2)}
    for i in range(X.shape[0]):
        assert abs(kernel(X[i], Y[i])/g['kernel_values'][i]-1) < 1e-13
        assert abs(kb(X[i], Y[i])/g['bkernel_values'][i]-1) < 1e-13
        d2 = ((X[i]-Y[i])**2).sum()
        assert abs(kernel.scalingValue*f[mode](a, d2)/g['kernel_values'][i]-1) < 1e-13
        # (2D: the library's table is the constant divided by |x-y|, and the surface form multiplies by |x-y| again)
        assert abs(bconst*f[bmode](ba, d2)/g['bkernel_values'][i]-1) < 1e-13
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'] if dim == 2 else g['boundaryVertices'])
    b = pb.nonlocalBuilder(pb.P1_DoFMap(mesh), kernel, {'target_order': 0.5} if dim == 2 else {})
    assert b.orders.target_order == float(g['target_order_used']) and b.orders.btarget_order == float(g['btarget_order_used'])
    assert b.orders.quad_order_diagonal == int(g['quad_order_diagonal'])
    assert b.orders.bquad_order_diagonal == int(g['bquad_order_diagonal'])
    with pytest.raises(NotImplementedError):
        pb.getIntegrableKernel(dim, str(g['kernelType']), 0.3)


def test_fe_order_host_side(golden_dir):
    """feFractionalOrder on the host: vertex values through the order's own DoFMap, ranking of the pair singularities (largest
    vertex value per cell / facet), quadrature orders as in the reference's local matrices; a mesh other than the order's is refused"""
    from test_oracle_golden import fe_order_vertex_values
    g = np.load(os.path.join(golden_dir, 'varorder_fe_disc_r2.npz'))
    mesh = pb.meshNd(g['vertices'], g['cells'], boundary=g['boundaryEdges'])
    dms = pb.P1_DoFMap(mesh, tag=np.zeros(mesh.num_vertices, dtype=bool))
    assert dms.num_dofs == mesh.num_vertices
    vs = fe_order_vertex_values(g)
    u = np.zeros(dms.num_dofs)
    for k in range(3):
        u[dms.dofs[:, k]] = vs[mesh.cells[:, k]]
    order = pb.feFractionalOrder(dms, u, float(g['smin']), float(g['smax']))
    b = pb.nonlocalBuilder(pb.P1_DoFMap(mesh), pb.getFractionalKernel(2, order), {'target_order': 0.5})
    assert b.orders.quad_order_diagonal == int(g['quad_order_diagonal']) and b.orders.bquad_order_diagonal == int(g['bquad_order_diagonal'])
    V = b._varorder
    assert np.array_equal(V['values'][V['cell_value']], vs[mesh.cells].max(axis=1))
    assert np.array_equal(V['values'][V['bfacet_value']], vs[np.asarray(mesh.boundaryFacets)].max(axis=1))
    assert b._varorder_struct().vertex_values
    with pytest.raises(NotImplementedError):
        pb.nonlocalBuilder(pb.P1_DoFMap(pb.refined(mesh, 1)), pb.getFractionalKernel(2, order), {'target_order': 0.5})
