from setuptools import setup, Extension
from Cython.Build import cythonize
setup(name='mpi4py_stub',
      ext_modules=cythonize([Extension('mpi4py.MPI', ['mpi4py/MPI.pyx'])],
                            compiler_directives={'language_level': '3'}))
