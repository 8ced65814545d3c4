"""Compiles integration/pnb200_shim.pyx in-tree against include/pnb200.h and pynucleus_b200/libpnb200.so
(Cython + gcc; no part of the reference is needed at build time)."""
import os
import subprocess
import sys
import sysconfig

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.abspath(os.path.join(HERE, '..'))


def build(force=False):
    import numpy as np
    src = os.path.join(HERE, 'pnb200_shim.pyx')
    csrc = os.path.join(HERE, 'pnb200_shim.c')
    out = os.path.join(HERE, 'pnb200_shim'+sysconfig.get_config_var('EXT_SUFFIX'))
    deps = [src, os.path.join(HERE, 'pnb200.pxd'), os.path.join(ROOT, 'include', 'pnb200.h')]
    if not force and os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return out
    subprocess.check_call([sys.executable, '-m', 'cython', '-3', '-I', HERE, src, '-o', csrc])
    inc = [sysconfig.get_paths()['include'], np.get_include(), os.path.join(ROOT, 'include')]
    libdir = os.path.join(ROOT, 'pynucleus_b200')
    cmd = ['gcc', '-O2', '-fPIC', '-shared', '-Wno-unused-function', '-DNPY_NO_DEPRECATED_API=NPY_1_7_API_VERSION']
    cmd += ['-I'+i for i in inc]
    cmd += [csrc, '-o', out, '-L'+libdir, '-lpnb200', '-Wl,-rpath,$ORIGIN/../pynucleus_b200']
    subprocess.check_call(cmd)
    return out


if __name__ == '__main__':
    print(build(force=True))
