"""ORACLE INFRASTRUCTURE: generates tests/golden/*.npz by running the
REFERENCE ITSELF (the stub-built copy in oracle/_ref, see build_reference.sh).

    PYTHONPATH=oracle/_ref python oracle/refbuild/make_golden.py

Every array below is an output of reference code (PyNucleus, nl/ and fem/);
nothing here is computed by this repository's own oracle or product.  The only
substitution is the 2D regular triangle rule family (oracle/triangle_rules.py,
provided through the modepy stand-in), see SURVEY.md section 8c.

The fixtures are committed; this script only needs to be re-run when a new
fixture is added.
"""
import os
import sys
import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..', '..'))
OUT = os.path.join(ROOT, 'tests', 'golden')
os.makedirs(OUT, exist_ok=True)

from PyNucleus_fem.mesh import simpleInterval, uniform_disc  # noqa: E402
from PyNucleus_fem.DoFMaps import P1_DoFMap  # noqa: E402
from PyNucleus_nl.kernels import getFractionalKernel  # noqa: E402
from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder  # noqa: E402
from PyNucleus_nl.fractionalOrders import constFractionalOrder  # noqa: E402
from PyNucleus_base.myTypes import REAL  # noqa: E402


def mesh_arrays(mesh, dm):
    out = dict(vertices=np.array(mesh.vertices), cells=np.array(mesh.cells),
               dofs=np.array(dm.dofs), num_dofs=dm.num_dofs,
               hVector=np.array(mesh.hVector), volVector=np.array(mesh.volVector),
               h=mesh.h, hmin=mesh.hmin, diam=mesh.diam)
    if mesh.dim == 2:
        out['boundaryEdges'] = np.array(mesh.boundaryEdges)
    else:
        out['boundaryVertices'] = np.array(mesh.boundaryVertices)
    return out


def rule_arrays(prefix, qr):
    return {prefix+'_nodes': np.array(qr.nodes), prefix+'_weights': np.array(qr.weights)}


def pair_dump(builder, mesh, pairs):
    """panel type, permutations and local matrix for the given cell pairs"""
    lm = builder.local_matrix
    nloc = np.array(builder.contrib).shape[0] if hasattr(builder, 'contrib') else None
    dpe = builder.dm.dofs_per_element
    nloc = (2*dpe)*(2*dpe+1)//2
    nv = mesh.cells.shape[1]
    panels = np.zeros(len(pairs), dtype=np.int32)
    perm1 = np.zeros((len(pairs), nv), dtype=np.int32)
    perm2 = np.zeros((len(pairs), nv), dtype=np.int32)
    perm = np.zeros((len(pairs), 2*dpe), dtype=np.int32)
    contribs = np.zeros((len(pairs), nloc))
    contrib = np.zeros((nloc, 1), dtype=REAL)
    for n, (c1, c2) in enumerate(pairs):
        lm.setCell1_py(int(c1))
        lm.setCell2_py(int(c2))
        p = lm.getPanelType()
        panels[n] = p
        perm1[n] = np.array(lm.perm1)
        perm2[n] = np.array(lm.perm2)
        pp = np.array(lm.perm)
        perm[n, :pp.shape[0]] = pp
        lm.eval_py(contrib, p)
        contribs[n] = contrib[:, 0]
    return dict(pairs=np.array(pairs, dtype=np.int32), panels=panels, perm1=perm1, perm2=perm2,
                perm=perm, contribs=contribs)


def boundary_dump(builder, mesh, pairs):
    lm = builder.local_matrix_zeroExterior
    surface = mesh.get_surface_mesh()
    lm.setMesh2_py(surface)
    dpe = builder.dm.dofs_per_element
    nloc = dpe*(dpe+1)//2
    panels = np.zeros(len(pairs), dtype=np.int32)
    contribs = np.zeros((len(pairs), nloc))
    contrib = np.zeros((nloc, 1), dtype=REAL)
    for n, (c1, c2) in enumerate(pairs):
        lm.setCell1_py(int(c1))
        lm.setCell2_py(int(c2))
        p = lm.getPanelType()
        panels[n] = p
        lm.eval_py(contrib, p)
        contribs[n] = contrib[:, 0]
    return dict(bpairs=np.array(pairs, dtype=np.int32), bpanels=panels, bcontribs=contribs,
                surface_cells=np.array(surface.cells))


def all_panels(builder, mesh):
    lm = builder.local_matrix
    nc = mesh.num_cells
    P = np.zeros((nc, nc), dtype=np.int8)
    for c1 in range(nc):
        lm.setCell1_py(c1)
        for c2 in range(c1, nc):
            lm.setCell2_py(c2)
            P[c1, c2] = lm.getPanelType()
    return P


def interval_case(noRef, s, name):
    mesh = simpleInterval(-1, 1)
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    kernel = getFractionalKernel(1, constFractionalOrder(s), np.inf)
    b = nonlocalBuilder(dm, kernel, {})
    A = np.array(b.getDense().data)
    b0 = nonlocalBuilder(dm, kernel, {}, zeroExterior=False)
    A0 = np.array(b0.getDense().data)
    out = mesh_arrays(mesh, dm)
    out.update(A=A, A_interior=A0, s=s, scaling=kernel.scalingValue,
               boundary_scaling=b.local_matrix_zeroExterior.kernel.scalingValue,
               target_order=b.local_matrix.target_order,
               quad_order_diagonal=b.local_matrix.quad_order_diagonal,
               boundary_target_order=b.local_matrix_zeroExterior.target_order,
               boundary_quad_order_diagonal=b.local_matrix_zeroExterior.quad_order_diagonal)
    out.update(rule_arrays('qrId', b.local_matrix.qrId))
    out.update(rule_arrays('qrVertex', b.local_matrix.qrVertex))
    out.update(rule_arrays('bqrVertex', b.local_matrix_zeroExterior.qrVertex))
    out['panel_matrix'] = all_panels(b, mesh)
    nc = mesh.num_cells
    pairs = [(c1, c2) for c1 in range(nc) for c2 in range(c1, nc)]
    if len(pairs) > 600:
        rng = np.random.RandomState(0)
        near = [(c1, c2) for (c1, c2) in pairs if c2-c1 <= 2]
        idx = rng.choice(len(pairs), 300, replace=False)
        pairs = near + [pairs[i] for i in idx]
    out.update(pair_dump(b, mesh, pairs))
    nb = mesh.get_surface_mesh().num_cells
    out.update(boundary_dump(b, mesh, [(c1, e) for c1 in range(nc) for e in range(nb)]))
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, A.shape)


def disc_case(noRef, s, name, full_pairs=False, with_A=True):
    mesh = uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    out = mesh_arrays(mesh, dm)
    if with_A:
        kernel = getFractionalKernel(2, constFractionalOrder(s), np.inf)
        params = {'target_order': 0.5}
        b = nonlocalBuilder(dm, kernel, params)
        A = np.array(b.getDense().data)
        b0 = nonlocalBuilder(dm, kernel, params, zeroExterior=False)
        A0 = np.array(b0.getDense().data)
        out.update(A=A, A_interior=A0, s=s, scaling=kernel.scalingValue,
                   boundary_scaling=b.local_matrix_zeroExterior.kernel.scalingValue,
                   target_order=b.local_matrix.target_order,
                   quad_order_diagonal=b.local_matrix.quad_order_diagonal,
                   quad_order_diagonalV=b.local_matrix.quad_order_diagonalV,
                   boundary_quad_order_diagonal=b.local_matrix_zeroExterior.quad_order_diagonal)
        out.update(rule_arrays('qrId', b.local_matrix.qrId))
        out.update(rule_arrays('qrEdge', b.local_matrix.qrEdge))
        out.update(rule_arrays('qrVertex', b.local_matrix.qrVertex))
        out.update(rule_arrays('bqrEdge', b.local_matrix_zeroExterior.qrEdge))
        out.update(rule_arrays('bqrVertex', b.local_matrix_zeroExterior.qrVertex))
        P = all_panels(b, mesh)
        out['panel_matrix'] = P
        nc = mesh.num_cells
        rng = np.random.RandomState(1)
        touching = [(c1, c2) for c1 in range(nc) for c2 in range(c1, nc) if P[c1, c2] < 0]
        distant = [(c1, c2) for c1 in range(nc) for c2 in range(c1, nc) if P[c1, c2] > 0]
        if not full_pairs:
            if len(touching) > 400:
                idx = rng.choice(len(touching), 400, replace=False)
                touching = [touching[i] for i in idx]
            # keep every order present plus a random sample
            byorder = {}
            for pr in distant:
                byorder.setdefault(int(P[pr]), []).append(pr)
            sample = []
            for o, lst in sorted(byorder.items()):
                idx = rng.choice(len(lst), min(len(lst), 12), replace=False)
                sample += [lst[i] for i in idx]
            distant = sample
        out.update(pair_dump(b, mesh, touching+distant))
        nb = mesh.get_surface_mesh().num_cells
        bp = [(c1, e) for c1 in range(nc) for e in range(nb)]
        if len(bp) > 700:
            idx = rng.choice(len(bp), 700, replace=False)
            bp = [bp[i] for i in idx]
        out.update(boundary_dump(b, mesh, bp))
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, out['cells'].shape, out.get('A', np.zeros((0, 0))).shape)


def rows_case(noRef, s, name, nrows=12):
    """Larger mesh (N = 2977 at noRef 5): the full matrix is too big to commit, keep sampled rows, the diagonal,
    and products with seeded vectors."""
    mesh = uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    out = mesh_arrays(mesh, dm)
    kernel = getFractionalKernel(2, constFractionalOrder(s), np.inf)
    A = np.array(nonlocalBuilder(dm, kernel, {'target_order': 0.5}).getDense().data)
    rng = np.random.default_rng(7)
    rows = np.sort(rng.choice(dm.num_dofs, nrows, replace=False))
    x = rng.standard_normal(dm.num_dofs)
    out.update(s=s, rows=rows, A_rows=A[rows], diagonal=np.diag(A).copy(), x=x, Ax=A.dot(x),
               ones_Ax=A.dot(np.ones(dm.num_dofs)), frobenius=np.linalg.norm(A))
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, dm.num_dofs)


def dm2_case(noRef, s, name):
    """two DoFMaps (rows: interior DoFs, columns: the complementary boundary DoFs), NA.pxi:1366-1378"""
    mesh = uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    dm2 = dm.getComplementDoFMap()
    out = mesh_arrays(mesh, dm)
    kernel = getFractionalKernel(2, constFractionalOrder(s), np.inf)
    b = nonlocalBuilder(dm, kernel, {'target_order': 0.5}, dm2=dm2)
    A = np.array(b.getDense().data)
    b0 = nonlocalBuilder(dm, kernel, {'target_order': 0.5}, zeroExterior=False, dm2=dm2)
    A0 = np.array(b0.getDense().data)
    out.update(s=s, dofs2=np.array(dm2.dofs), num_dofs2=dm2.num_dofs, A_bc=A, A_bc_interior=A0)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, A.shape)


def entry_case(dim, noRef, s, name, nsample=10):
    """nonlocalBuilder.getEntry / getDiagonal (NA.pxi:1539-1660, 2269-2289): single entries integrated over the patch
    of the two basis functions, the rest of the space replaced by a surface integral around the patch"""
    mesh = simpleInterval(-1, 1) if dim == 1 else uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    out = mesh_arrays(mesh, dm)
    kernel = getFractionalKernel(dim, constFractionalOrder(s), np.inf)
    b = nonlocalBuilder(dm, kernel, {'target_order': 0.5})
    rng = np.random.default_rng(11)
    IJ = [(int(i), int(i)) for i in rng.choice(dm.num_dofs, nsample, replace=False)]
    IJ += [(int(i), int(j)) for i, j in rng.integers(0, dm.num_dofs, (nsample, 2))]
    # neighbouring DoFs (sharing a cell)
    dofs = np.array(dm.dofs)
    for c in rng.choice(mesh.num_cells, nsample, replace=False):
        d = dofs[c][dofs[c] >= 0]
        if d.shape[0] >= 2:
            IJ.append((int(d[0]), int(d[1])))
    IJ = np.array(IJ, dtype=np.int64)
    vals = np.array([b.getEntry(int(i), int(j)) for i, j in IJ])
    diag = np.array(b.getDiagonal().data)
    out.update(s=s, target_order=0.5, IJ=IJ, entries=vals, diagonal=diag)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, dm.num_dofs, vals[:3], diag[:3])


def solver_case(noRef, s, name):
    """Krylov solvers of the reference (base/PyNucleus_base/solvers.pyx: cg_solver :329-445, gmres_solver :458-660) with
    the Jacobi preconditioner (invDiagonal) on the assembled operator: solutions and residual histories"""
    from PyNucleus_base.solvers import cg_solver, gmres_solver
    from PyNucleus_base.linear_operators import invDiagonal
    mesh = uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    out = mesh_arrays(mesh, dm)
    kernel = getFractionalKernel(2, constFractionalOrder(s), np.inf)
    A = nonlocalBuilder(dm, kernel, {'target_order': 0.5}).getDense()
    rng = np.random.default_rng(3)
    b = rng.standard_normal(dm.num_dofs)
    out.update(s=s, target_order=0.5, b=b)
    for tag, prec in (('', True), ('_noprec', False)):
        cg = cg_solver(A)
        if prec:
            cg.setPreconditioner(invDiagonal(A))
        cg.tolerance = 1e-10
        cg.maxIter = 200
        cg.setup()
        x = np.zeros(dm.num_dofs)
        its = cg(b, x)
        out.update({'cg_x'+tag: x.copy(), 'cg_iterations'+tag: its, 'cg_residuals'+tag: np.array(cg.residuals)})
        for left in (True, False):
            gm = gmres_solver(A)
            if prec:
                gm.setPreconditioner(invDiagonal(A), left)
            gm.tolerance = 1e-10
            gm.maxIter = 12
            gm.restarts = 20
            gm.setup()
            x = np.zeros(dm.num_dofs)
            its = gm(b, x)
            key = 'gmres_'+('left' if left else 'right')+tag
            out.update({key+'_x': x.copy(), key+'_iterations': its, key+'_residuals': np.array(gm.residuals)})
            print(key, its, len(gm.residuals), gm.residuals[-1])
        print('cg'+tag, its, len(cg.residuals), cg.residuals[-1])
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, dm.num_dofs)


def varconst_case(noRef, s, name):
    """variable-order code path of the reference with s(x,y) = const (config 4): dense matrix only"""
    from PyNucleus_nl.fractionalOrders import variableConstFractionalOrder
    mesh = uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    kernel = getFractionalKernel(2, variableConstFractionalOrder(s), np.inf)
    assert kernel.variable
    b = nonlocalBuilder(dm, kernel, {'target_order': 0.5})
    A = np.array(b.getDense().data)
    out = mesh_arrays(mesh, dm)
    out.update(A=A, s=s, target_order=0.5)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, A.shape)


def leftright_case(noRef, sll, srr, slr, name):
    """piecewise constant, symmetric variable order s(x,y) (leftRightFractionalOrder, fractionalOrders.pyx:285-335):
    the kernel parameters change per cell pair (kernel.evalParams at the cell centres, NO.pxi:509-513)"""
    from PyNucleus_nl.fractionalOrders import leftRightFractionalOrder
    mesh = uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    kernel = getFractionalKernel(2, leftRightFractionalOrder(sll, srr, slr, slr, 0.), np.inf)
    assert kernel.variable and kernel.symmetric
    b = nonlocalBuilder(dm, kernel, {'target_order': 0.5})
    A = np.array(b.getDense().data)
    b0 = nonlocalBuilder(dm, kernel, {'target_order': 0.5}, zeroExterior=False)
    A0 = np.array(b0.getDense().data)
    out = mesh_arrays(mesh, dm)
    out.update(A=A, A_interior=A0, sll=sll, srr=srr, slr=slr, interface=0., target_order=0.5,
               quad_order_diagonal=b.local_matrix.quad_order_diagonal, quad_order_diagonalV=b.local_matrix.quad_order_diagonalV,
               boundary_quad_order_diagonal=b.local_matrix_zeroExterior.quad_order_diagonal)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, A.shape)


def h2_case(dim, noRef, s, name, max_far=60):
    """getH2 of the reference: cluster tree, admissible / near pairs, far-field kernel blocks
    (clusterMethodCy.pyx:2153-2238), near-field CSR matrix, and H2 matvec of a fixed vector"""
    if dim == 2:
        mesh = uniform_disc()
        params = {'target_order': 0.5}
    else:
        mesh = simpleInterval(-1, 1)
        params = {}
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    kernel = getFractionalKernel(dim, constFractionalOrder(s), np.inf)
    b = nonlocalBuilder(dm, kernel, params)
    H, Pnear = b.getH2(returnNearField=True)
    out = mesh_arrays(mesh, dm)
    out.update(s=s, target_order=params.get('target_order', np.nan), repr=str(H))
    # tree
    nodes = list(H.tree.get_tree_nodes())
    ids = np.array([n.id for n in nodes], dtype=np.int64)
    out['tree_ids'] = ids
    out['tree_boxes'] = np.array([np.array(n.box) for n in nodes])
    out['tree_parent'] = np.array([n.parent.id if n.parent is not None else -1 for n in nodes], dtype=np.int64)
    out['tree_isleaf'] = np.array([n.isLeaf for n in nodes], dtype=bool)
    out['tree_level'] = np.array([n.levelNo for n in nodes], dtype=np.int32)
    out['tree_order'] = np.array([n.interpolation_order for n in nodes], dtype=np.int32)
    dofptr = [0]
    dofs = []
    for n in nodes:
        d = np.sort(np.array(n.dofs.toArray()))
        dofs.append(d)
        dofptr.append(dofptr[-1]+d.shape[0])
    out['tree_dofptr'] = np.array(dofptr, dtype=np.int64)
    out['tree_dofs'] = np.concatenate(dofs).astype(np.int32)
    out['near_pairs'] = np.array([(cP.n1.id, cP.n2.id) for cP in Pnear], dtype=np.int64)
    far = []
    for lvl in sorted(H.Pfar):
        for cP in H.Pfar[lvl]:
            far.append((lvl, cP))
    out['far_pairs'] = np.array([(lvl, cP.n1.id, cP.n2.id) for lvl, cP in far], dtype=np.int64)
    rng = np.random.RandomState(3)
    sel = np.sort(rng.choice(len(far), min(max_far, len(far)), replace=False))
    out['far_sel'] = sel
    out['far_box1'] = np.array([np.array(far[i][1].n1.box) for i in sel])
    out['far_box2'] = np.array([np.array(far[i][1].n2.box) for i in sel])
    out['far_m1'] = np.array([far[i][1].n1.interpolation_order for i in sel], dtype=np.int32)
    out['far_m2'] = np.array([far[i][1].n2.interpolation_order for i in sel], dtype=np.int32)
    blocks = [np.array(far[i][1].kernelInterpolant).ravel() for i in sel]
    out['far_ptr'] = np.concatenate(([0], np.cumsum([b_.shape[0] for b_ in blocks]))).astype(np.int64)
    out['far_blocks'] = np.concatenate(blocks)
    An = H.Anear
    out['Anear_indptr'] = np.array(An.indptr)
    out['Anear_indices'] = np.array(An.indices)
    out['Anear_data'] = np.array(An.data)
    if hasattr(An, 'diagonal') and An.__class__.__name__.startswith('SSS'):
        out['Anear_diagonal'] = np.array(An.diagonal)
    out['Anear_type'] = An.__class__.__name__
    x = np.sin(np.arange(dm.num_dofs)*0.37)+0.1
    out['x'] = x
    out['Hx'] = H*x
    A = np.array(b.getDense().data)
    out['Ax'] = A.dot(x)
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, H)


def h2_regional_case(dim, noRef, s, name):
    """regional operator (zeroExterior=False): near field of getH2 (assembleClusters :1840-1912: surface terms around the
    cluster unions minus the surface terms of the domain boundary), H2 / dense matvec, and getEntry values"""
    if dim == 2:
        mesh = uniform_disc()
        params = {'target_order': 0.5}
    else:
        mesh = simpleInterval(-1, 1)
        params = {}
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    kernel = getFractionalKernel(dim, constFractionalOrder(s), np.inf)
    b = nonlocalBuilder(dm, kernel, params, zeroExterior=False)
    H, Pnear = b.getH2(returnNearField=True)
    out = mesh_arrays(mesh, dm)
    out.update(s=s, target_order=params.get('target_order', np.nan), repr=str(H))
    out['near_pairs'] = np.array([(cP.n1.id, cP.n2.id) for cP in Pnear], dtype=np.int64)
    An = H.Anear
    out['Anear_indptr'] = np.array(An.indptr)
    out['Anear_indices'] = np.array(An.indices)
    out['Anear_data'] = np.array(An.data)
    out['Anear_diagonal'] = np.array(An.diagonal)
    out['Anear_type'] = An.__class__.__name__
    x = np.sin(np.arange(dm.num_dofs)*0.37)+0.1
    out['x'] = x
    out['Hx'] = H*x
    out['Ax'] = np.array(b.getDense().data).dot(x)
    rng = np.random.default_rng(5)
    IJ = [(int(i), int(i)) for i in rng.choice(dm.num_dofs, 6, replace=False)]
    dofs = np.array(dm.dofs)
    for c in rng.choice(mesh.num_cells, 6, replace=False):
        d = dofs[c][dofs[c] >= 0]
        if d.shape[0] >= 2:
            IJ.append((int(d[0]), int(d[1])))
    out['IJ'] = np.array(IJ, dtype=np.int64)
    out['entries'] = np.array([b.getEntry(i, j) for i, j in IJ])
    np.savez_compressed(os.path.join(OUT, name), **out)
    print(name, H, An.__class__.__name__)


def kernel_values():
    """closed-form style spot values straight from the reference kernels"""
    rng = np.random.RandomState(2)
    out = {}
    for dim in (1, 2):
        x = rng.rand(20, dim)
        y = rng.rand(20, dim)+1.5
        for s in (0.25, 0.75):
            k = getFractionalKernel(dim, constFractionalOrder(s), np.inf)
            kb = k.getBoundaryKernel()
            out['x_%dd' % dim] = x
            out['y_%dd' % dim] = y
            out['k_%dd_s%g' % (dim, s)] = np.array([k(x[i], y[i]) for i in range(20)])
            out['kb_%dd_s%g' % (dim, s)] = np.array([kb(x[i], y[i]) for i in range(20)])
            out['C_%dd_s%g' % (dim, s)] = k.scalingValue
            out['Cb_%dd_s%g' % (dim, s)] = kb.scalingValue
    np.savez_compressed(os.path.join(OUT, 'kernel_values'), **out)
    print('kernel_values')


if __name__ == '__main__':
    which = sys.argv[1:] or ['all']
    if 'all' in which or 'interval' in which:
        interval_case(3, 0.25, 'interval_s0.25_r3')
        interval_case(6, 0.25, 'interval_s0.25_r6')
        interval_case(5, 0.75, 'interval_s0.75_r5')
    if 'all' in which or 'disc' in which:
        disc_case(0, 0.75, 'disc_mesh_r0', with_A=False)
        disc_case(1, 0.75, 'disc_s0.75_r1', full_pairs=True)
        disc_case(2, 0.75, 'disc_s0.75_r2')
        disc_case(3, 0.75, 'disc_s0.75_r3')
        disc_case(2, 0.25, 'disc_s0.25_r2')
        disc_case(4, 0.75, 'disc_mesh_r4', with_A=False)
    if 'all' in which or 'rows' in which:
        rows_case(5, 0.75, 'disc_s0.75_r5_rows')
    if 'all' in which or 'leftright' in which:
        leftright_case(2, 0.25, 0.75, 0.5, 'disc_leftright_r2')
        leftright_case(3, 0.4, 0.8, 0.3, 'disc_leftright_r3')
    if 'all' in which or 'dm2' in which:
        dm2_case(2, 0.75, 'disc_dm2_s0.75_r2')
        dm2_case(3, 0.25, 'disc_dm2_s0.25_r3')
    if 'all' in which or 'entry' in which:
        entry_case(2, 3, 0.75, 'entry_disc_s0.75_r3')
        entry_case(1, 6, 0.25, 'entry_interval_s0.25_r6')
    if 'all' in which or 'solvers' in which:
        solver_case(3, 0.75, 'solvers_disc_s0.75_r3')
    if 'all' in which or 'varconst' in which:
        varconst_case(2, 0.75, 'disc_varconst0.75_r2')
        varconst_case(3, 0.4, 'disc_varconst0.4_r3')
    if 'all' in which or 'h2' in which:
        h2_case(2, 4, 0.75, 'h2_disc_s0.75_r4')
        h2_case(1, 8, 0.25, 'h2_interval_s0.25_r8')
    if 'all' in which or 'h2regional' in which:
        h2_regional_case(2, 4, 0.75, 'h2_regional_disc_s0.75_r4')
        h2_regional_case(1, 8, 0.25, 'h2_regional_interval_s0.25_r8')
    if 'all' in which or 'kernels' in which:
        kernel_values()
