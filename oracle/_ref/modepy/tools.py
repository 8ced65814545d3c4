import numpy as np


def unit_to_barycentric(unit):
    if hasattr(unit, 'bary'):
        return np.array(unit.bary, copy=True)
    dims = unit.shape[0]
    bary = np.empty((dims + 1,) + unit.shape[1:], dtype=unit.dtype)
    bary[:dims] = (unit + 1.) / 2.
    bary[dims] = 1. - bary[:dims].sum(axis=0)
    return bary
