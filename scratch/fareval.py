import sys, os
sys.path.insert(0,'.')
import numpy as np, torch
import pynucleus_b200 as pb
from pynucleus_b200 import _lib
mesh = pb.refined(pb.polygon_disc(10), 5); dm = pb.P1_DoFMap(mesh)
b = pb.nonlocalBuilder(dm, pb.getFractionalKernel(2,0.75), {'target_order':0.5})
b.problem
pk = np.zeros(1)
_lib.check(_lib.lib().pnb_fp64_peak(0, pk.ctypes.data_as(_lib.c_double_p))); print('peak', pk[0])
