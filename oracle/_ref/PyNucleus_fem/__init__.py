###################################################################################
# Copyright 2021 National Technology & Engineering Solutions of Sandia,           #
# LLC (NTESS). Under the terms of Contract DE-NA0003525 with NTESS, the           #
# U.S. Government retains certain rights in this software.                        #
# If you want to use this code, please refer to the README.rst and LICENSE files. #
###################################################################################

from . mesh import (PHYSICAL, INTERIOR_NONOVERLAPPING, INTERIOR, NO_BOUNDARY,
                    DIRICHLET, HOMOGENEOUS_DIRICHLET,
                    NEUMANN, HOMOGENEOUS_NEUMANN,
                    NORM, boundaryConditions)
from . DoFMaps import (P0_DoFMap, P1_DoFMap, P2_DoFMap, P3_DoFMap,
                       str2DoFMap, str2DoFMapOrder, getAvailableDoFMaps)
from . factories import functionFactory, dofmapFactory, meshFactory
from . pdeProblems import diffusionProblem, helmholtzProblem
__all__ = ['functionFactory', 'dofmapFactory', 'meshFactory',
           'diffusionProblem', 'helmholtzProblem',
           'PHYSICAL', 'INTERIOR_NONOVERLAPPING', 'INTERIOR', 'NO_BOUNDARY',
           'DIRICHLET', 'HOMOGENEOUS_DIRICHLET',
           'NEUMANN', 'HOMOGENEOUS_NEUMANN',
           'NORM', 'boundaryConditions',
           'P0_DoFMap', 'P1_DoFMap', 'P2_DoFMap', 'P3_DoFMap',
           'str2DoFMap', 'str2DoFMapOrder', 'getAvailableDoFMaps']
