"""The compiled Cython shim (integration/pnb200_shim.pyx) between the REFERENCE's own objects and libpnb200:
in one process, PyNucleus' nonlocalBuilder.getDense (its Cython loops, run from the stub-built copy in oracle/_ref)
against the same call served by the CUDA library through the C ABI -- the reference's mesh, DoFMap, kernel objects and
quadrature tables on both sides.  Mirrors what the reference's tests/test_fracLapl.py does with its builders.

Test infrastructure: oracle/_ref is only imported here, as the checker."""
import os
import sys

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

ROOT = os.path.abspath(os.path.join(os.path.dirname(__file__), '..'))
REF = os.path.join(ROOT, 'oracle', '_ref')
TOL = 1e-12


@pytest.fixture(scope='module')
def shim():
    if not os.path.isdir(os.path.join(REF, 'PyNucleus_nl')):
        pytest.skip('stub-built reference (oracle/_ref) not present')
    for p in (REF, os.path.join(ROOT, 'integration')):
        if p not in sys.path:
            sys.path.insert(0, p)
    sys.path.insert(0, ROOT)
    from integration.build_shim import build
    build()
    import pnb200_shim
    return pnb200_shim


def entry_err(A, Aref):
    d = np.sqrt(np.abs(np.diag(Aref)))
    scale = np.maximum(np.abs(Aref), 1e-2*np.outer(d, d))
    return (np.abs(A-Aref)/scale).max()


def ref_problem(dim, noRef, ktype, s, horizon):
    from PyNucleus_fem.mesh import simpleInterval, uniform_disc
    from PyNucleus_fem.DoFMaps import P1_DoFMap
    from PyNucleus_fem.functions import constant
    from PyNucleus_nl.kernels import getFractionalKernel, getIntegrableKernel
    from PyNucleus_nl.fractionalOrders import constFractionalOrder
    mesh = simpleInterval(-1, 1) if dim == 1 else uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = P1_DoFMap(mesh)
    if ktype == 'fractional':
        kernel = getFractionalKernel(dim, constFractionalOrder(s), constant(horizon))
    else:
        kernel = getIntegrableKernel(dim, ktype, constant(horizon))
    return mesh, dm, kernel


@pytest.mark.parametrize('dim,noRef,ktype,s,horizon,zeroExterior', [
    (1, 5, 'fractional', 0.25, np.inf, True), (1, 5, 'fractional', 0.75, np.inf, False),
    (2, 2, 'fractional', 0.75, np.inf, True), (2, 3, 'fractional', 0.25, np.inf, True),
    (2, 3, 'fractional', 0.75, 0.7, True), (2, 3, 'constant', 0., 0.7, True), (1, 5, 'inverseDistance', 0., 0.4, True)])
def test_reference_builder_vs_shim(shim, dim, noRef, ktype, s, horizon, zeroExterior):
    from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder
    mesh, dm, kernel = ref_problem(dim, noRef, ktype, s, horizon)
    params = {'target_order': 0.5} if dim == 2 else {}
    Aref = np.array(nonlocalBuilder(dm, kernel, dict(params), zeroExterior=zeroExterior).getDense().data)
    B = shim.builder_class()
    b = B(dm, kernel, dict(params), zeroExterior=zeroExterior)
    assert shim.supported(b)
    A = b.getDense()
    assert type(A).__name__ == 'Dense_LinearOperator' and A.shape == Aref.shape
    assert entry_err(np.array(A.data), Aref) < TOL
    # the returned object is the reference's operator class: its own matvec works on it
    x = np.linspace(0., 1., dm.num_dofs)
    assert np.allclose(A*x, Aref.dot(x), rtol=1e-12, atol=1e-12*np.abs(Aref.dot(x)).max())


def test_assembleNonlocal_resolves_to_the_shim(shim):
    """DoFMap.assembleNonlocal (fem/PyNucleus_fem/DoFMaps.pyx:877-899) looks the builder up in PyNucleus_nl at call
    time; after install() it is the accelerated subclass"""
    import PyNucleus_nl
    from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder
    mesh, dm, kernel = ref_problem(2, 2, 'fractional', 0.75, np.inf)
    Aref = np.array(nonlocalBuilder(dm, kernel, {'target_order': 0.5}).getDense().data)
    old = PyNucleus_nl.nonlocalBuilder
    try:
        cls = shim.install()
        assert PyNucleus_nl.nonlocalBuilder is cls
        A = dm.assembleNonlocal(kernel, matrixFormat='dense', params={'target_order': 0.5})
        assert entry_err(np.array(A.data), Aref) < TOL
    finally:
        PyNucleus_nl.nonlocalBuilder = old


def test_unsupported_configuration_falls_through(shim):
    """two DoFMaps are outside the shim: the subclass hands the call to the reference's own getDense"""
    mesh, dm, kernel = ref_problem(1, 4, 'fractional', 0.25, np.inf)
    dm2 = dm.getComplementDoFMap()
    B = shim.builder_class()
    b = B(dm, kernel, {}, dm2=dm2)
    assert not shim.supported(b)
    A = b.getDense()
    assert A.shape == (dm.num_dofs, dm2.num_dofs)


@pytest.mark.parametrize('dim,noRef,s,element', [(1, 5, 0.75, 'P2'), (2, 2, 0.75, 'P2'), (2, 2, 0.25, 'P2'), (1, 5, 0.25, 'P0'),
                                                 (2, 2, 0.3, 'P0'), (1, 4, 0.75, 'P3')])
def test_reference_builder_vs_shim_p2(shim, dim, noRef, s, element):
    """P2_DoFMap / P0_DoFMap of the reference through the shim (pnb_dense_assemble_element) against its own Cython getDense"""
    from PyNucleus_fem.mesh import simpleInterval, uniform_disc
    from PyNucleus_fem.DoFMaps import P0_DoFMap, P2_DoFMap, P3_DoFMap
    from PyNucleus_nl.kernels import getFractionalKernel
    from PyNucleus_nl.fractionalOrders import constFractionalOrder
    from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder
    mesh = simpleInterval(-1, 1) if dim == 1 else uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = {'P2': P2_DoFMap, 'P0': P0_DoFMap, 'P3': P3_DoFMap}[element](mesh)
    kernel = getFractionalKernel(dim, constFractionalOrder(s), np.inf)
    params = {'target_order': 0.5} if dim == 2 else {}
    Aref = np.array(nonlocalBuilder(dm, kernel, dict(params)).getDense().data)
    b = shim.builder_class()(dm, kernel, dict(params))
    assert shim.supported(b)
    A = b.getDense()
    assert A.shape == Aref.shape
    assert entry_err(np.array(A.data), Aref) < TOL


@pytest.mark.parametrize('dim,noRef,s,lam,element', [(1, 5, 0.75, 2.0, 'P1'), (2, 2, 0.25, 1.0, 'P1'), (1, 4, 0.75, 1.5, 'P2')])
def test_reference_builder_vs_shim_tempered(shim, dim, noRef, s, lam, element):
    """a tempered fractional kernel of the reference (kernelType FRACTIONAL with temperedValue != 0) through the shim: the row-owner
    kernel with the rate (pnb_dense_assemble_element_tempered), not the plain power law, against the reference's Cython getDense"""
    from PyNucleus_fem.mesh import simpleInterval, uniform_disc
    from PyNucleus_fem.DoFMaps import P1_DoFMap, P2_DoFMap
    from PyNucleus_nl.kernels import getFractionalKernel
    from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder
    mesh = simpleInterval(-1, 1) if dim == 1 else uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = {'P1': P1_DoFMap, 'P2': P2_DoFMap}[element](mesh)
    kernel = getFractionalKernel(dim, s, np.inf, tempered=lam)
    params = {'target_order': 0.5} if dim == 2 else {}
    for ze in (True, False):
        Aref = np.array(nonlocalBuilder(dm, kernel, dict(params), zeroExterior=ze).getDense().data)
        b = shim.builder_class()(dm, kernel, dict(params), zeroExterior=ze)
        assert shim.supported(b)
        assert entry_err(np.array(b.getDense().data), Aref) < TOL


@pytest.mark.parametrize('dim,noRef,ktype,kw,element', [(1, 5, 'gaussian', {'variance': 0.1}, 'P1'), (1, 5, 'exponential', {'exponentialRate': 8.0}, 'P1'),
                                                        (2, 2, 'gaussian', {'variance': 0.2}, 'P1'), (1, 4, 'gaussian', {'variance': 0.1}, 'P2')])
def test_reference_builder_vs_shim_gaussian_exponential(shim, dim, noRef, ktype, kw, element):
    """the Gaussian / exponential kernels of the reference's driver tests (full space) through the shim
    (pnb_dense_assemble_element_smooth) against the reference's Cython getDense, with and without the surface terms"""
    from PyNucleus_fem.mesh import simpleInterval, uniform_disc
    from PyNucleus_fem.DoFMaps import P1_DoFMap, P2_DoFMap
    from PyNucleus_nl.kernels import getIntegrableKernel
    from PyNucleus_nl.nonlocalAssembly import nonlocalBuilder
    mesh = simpleInterval(-1, 1) if dim == 1 else uniform_disc()
    for _ in range(noRef):
        mesh = mesh.refine()
    dm = {'P1': P1_DoFMap, 'P2': P2_DoFMap}[element](mesh)
    kernel = getIntegrableKernel(dim, ktype, np.inf, **kw)
    params = {'target_order': 0.5} if dim == 2 else {}
    for ze in (True, False):
        Aref = np.array(nonlocalBuilder(dm, kernel, dict(params), zeroExterior=ze).getDense().data)
        b = shim.builder_class()(dm, kernel, dict(params), zeroExterior=ze)
        assert shim.supported(b)
        assert entry_err(np.array(b.getDense().data), Aref) < TOL
