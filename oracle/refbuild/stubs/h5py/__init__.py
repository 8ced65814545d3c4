"""Stub: the reference imports h5py at module level (nl/PyNucleus_nl/helpers.py:35)
but the assembly path never touches HDF5.  Oracle build infrastructure only."""


class Empty:
    def __init__(self, dtype=None):
        self.dtype = dtype


class Group:
    pass


class File:
    def __init__(self, *args, **kwargs):
        raise NotImplementedError('h5py is not available in the oracle build')
