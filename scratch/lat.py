import sys, ctypes, numpy as np
sys.path.insert(0,'.')
from pynucleus_b200 import _lib
pk = np.zeros(1)
_lib.check(_lib.lib().pnb_fp64_peak(0, pk.ctypes.data_as(_lib.c_double_p))); print('peak', pk[0])
