import sys
sys.path.insert(0, '.')
from math import gamma
import numpy as np
import pynucleus_b200 as pb
for radial in (False, True):
    for refinements in (2, 3):
        for element in ('P1', 'P2'):
            for s in (0.3, 0.7):
                fan = pb.polygon_disc(10)
                mesh = fan if radial else pb.meshNd(fan.vertices, fan.cells)
                for _ in range(refinements):
                    mesh = mesh.refine()
                dm = pb.P1_DoFMap(mesh) if element == 'P1' else pb.P2_DoFMap(mesh)
                A = pb.assembleNonlocalOperator(mesh, dm, pb.constFractionalOrder(s)).data
                rhs = dm.assembleRHS(1.)
                u = np.linalg.solve(A, rhs)
                err = np.sqrt(abs(np.dot(rhs, u)-2*np.pi*2**(-2*s)*gamma(1)/gamma(1+s)**2/2/(s+1)))
                print('radial', radial, 'ref', refinements, element, s, 'cells', mesh.num_cells, 'err %.3f' % err)
