from setuptools import setup, Extension
from Cython.Build import cythonize
import numpy
setup(name='metisCy_stub',
      ext_modules=cythonize([Extension('PyNucleus_metisCy.metisCy', ['PyNucleus_metisCy/metisCy.pyx'],
                                       include_dirs=[numpy.get_include()])],
                            compiler_directives={'language_level': '3'}))
