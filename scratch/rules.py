import sys
sys.path.insert(0,'.')
import numpy as np, pynucleus_b200 as pb
from pynucleus_b200 import quadrature
sides, r = int(sys.argv[1]), int(sys.argv[2])
mesh = pb.refined(pb.polygon_disc(sides), r); dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5})
import inspect
print([k for k in dir(b) if not k.startswith('__')])
o = b.orders
print(o)
t = quadrature.singular_tables(2, b.kernel.singularityValue, b.kernelBoundary.singularityValue if hasattr(b.kernelBoundary,'singularityValue') else -2.5, o)
for k,(bary,w) in t.items(): print(k, bary.shape, w.shape)
