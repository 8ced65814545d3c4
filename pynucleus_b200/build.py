"""In-tree build of libpnb200.so with nvcc for sm_100a (no JIT cache: the .so
travels with the repository snapshot)."""
import os
import subprocess

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'pnb200.cu')
OUT = os.path.join(HERE, 'libpnb200.so')
DEPS = [SRC, os.path.join(HERE, 'csrc', 'pnb_device.cuh'), os.path.join(HERE, 'csrc', 'pnb_pair.cuh'),
        os.path.join(HERE, 'csrc', 'pnb_group.cuh'), os.path.join(HERE, 'csrc', 'pnb_h2.cuh'), os.path.join(HERE, 'csrc', 'pnb_element.cuh'), os.path.join(HERE, 'csrc', 'pnb_krylov.cuh'), os.path.join(HERE, 'csrc', 'pnb_varorder.cuh'),
        os.path.join(HERE, '..', 'include', 'pnb200.h')]

# per-thread default streams: host threads that assemble independent problems (H2 near field) overlap on the device
NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17', '--default-stream', 'per-thread',
              '-Xcompiler', '-fPIC', '-shared', '-lpthread']


def build(force=False, verbose=False):
    if not force and os.path.exists(OUT) and all(os.path.getmtime(OUT) >= os.path.getmtime(d) for d in DEPS):
        return OUT
    nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
    cmd = [nvcc]+NVCC_FLAGS+(['-Xptxas', '-v'] if verbose else [])+['-o', OUT, SRC]
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    print(build(force=True, verbose=True))
