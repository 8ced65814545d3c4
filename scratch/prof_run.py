import sys, os
sys.path.insert(0,'.')
import pynucleus_b200._lib as L
L.LIB_PATH = os.path.join(os.path.dirname(L.LIB_PATH), 'libpnb200_prof.so')
exec(open('scratch/timing3.py').read())
