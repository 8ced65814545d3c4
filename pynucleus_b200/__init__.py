"""pynucleus_b200: B200-native nonlocal operator assembly behind PyNucleus'
nonlocalBuilder API.  The compute path is hand-written CUDA for
sm_100a in libpnb200.so (C ABI: include/pnb200.h); this package is the
host-side mirror of the reference interface."""
from .mesh import simpleInterval, uniform_disc, polygon_disc, refined, meshNd  # noqa: F401
from .dofmap import P0_DoFMap, P1_DoFMap, P2_DoFMap, P3_DoFMap  # noqa: F401
from .kernels import (getFractionalKernel, getKernel, FractionalKernel, constFractionalOrder, variableConstFractionalOrder, leftRightFractionalOrder,  # noqa: F401
                      piecewiseConstantFractionalOrder, constantNonSymFractionalOrder, layersFractionalOrder, innerOuterFractionalOrder, islandsFractionalOrder,
                      constantFractionalLaplacianScaling, FRACTIONAL, INDICATOR, PERIDYNAMIC, Kernel, getIntegrableKernel,
                      constantIntegrableScaling, constant, singleVariableUnsymmetricFractionalOrder,
                      smoothedLeftRightFractionalOrder, linearLeftRightFractionalOrder, feFractionalOrder,
                      variableFractionalLaplacianScaling)
from .assembly import nonlocalBuilder, assembleNonlocalOperator, release_staging_pool  # noqa: F401
from .linear_operators import Dense_LinearOperator, diagonalOperator  # noqa: F401
from .solvers import cg, gmres, lu, DistributedDenseOperator  # noqa: F401

__version__ = '0.1.0'
from .multigrid import multigrid, hierarchy, buildRestrictionProlongation  # noqa: F401,E402
