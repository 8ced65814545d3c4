"""Serial stand-in for mpi4py used ONLY to build/import the reference
(PyNucleus) in a container without MPI.  rank=0/size=1 semantics; a
``fakeComm(rank, size)`` lets ``nonlocalBuilder.getDense`` compute the cell
slice of one MPI rank (nl/PyNucleus_nl/nonlocalAssembly_{SCALAR}.pxi:1280-1285).
Test infrastructure: never imported by the product."""
import os
__version__ = '4.0.0'


def get_include():
    return os.path.join(os.path.dirname(__file__), 'include')


def get_config():
    return {'mpicc': 'gcc', 'mpicxx': 'g++'}


from . import rc  # noqa
