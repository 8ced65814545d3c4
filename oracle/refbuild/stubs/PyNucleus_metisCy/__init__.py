"""Stub for the METIS bindings: only the type names that
fem/PyNucleus_fem/repartitioner.pyx:16-18 needs at import time.  Partitioning
calls raise.  Oracle build infrastructure only."""
from . import metisCy  # noqa


def __getattr__(name):
    raise NotImplementedError('METIS is not available in the oracle build ({})'.format(name))
