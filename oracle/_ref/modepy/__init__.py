"""Stand-in for the un-vendored ``modepy`` dependency of the reference
(fem/PyNucleus_fem/quadrature.pyx:13-14).

``XiaoGimbutasSimplexQuadrature(order, 2)`` returns the repository's adopted
triangle-rule family (oracle/triangle_rules.py) instead of the unavailable
Xiao-Gimbutas tables, packaged so that the reference's own post-processing
(quadrature.pyx:531-545: ``unit_to_barycentric`` + weights*0.5) reproduces the
family's barycentric nodes and weights BIT-EXACTLY.  Oracle build
infrastructure only."""
import os
import sys
import numpy as np

_root = os.path.abspath(os.path.join(os.path.dirname(__file__), '..', '..', '..'))
if _root not in sys.path:
    sys.path.insert(0, _root)
from oracle.triangle_rules import family  # noqa: E402


class _UnitNodes(np.ndarray):
    """unit-coordinate nodes that remember the barycentric originals"""
    pass


class XiaoGimbutasSimplexQuadrature:
    def __init__(self, order, dims):
        if dims != 2:
            raise NotImplementedError('only triangles are provided by the stub')
        bary, w = family(order)
        # modepy convention: unit coords r in [-1,1]^2, bary[:2]=(r+1)/2, bary[2]=1-sum
        unit = (2. * bary[:2] - 1.).view(_UnitNodes)
        unit.bary = np.ascontiguousarray(np.vstack((bary[0:1], bary[1:2], bary[2:3])))
        self.nodes = unit
        self.weights = 2. * w     # reference multiplies by 0.5 in place
        self.exact_to = int(order)
        self.dim = dims
