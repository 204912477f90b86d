"""Build libmvb.so (the C-ABI CUDA library) in-tree for sm_100a.

    python -m mvpy_b200.build          # or  __graft_entry__.build()

nvcc cross-compiles without a GPU; the resulting mvpy_b200/libmvb.so is git-ignored but travels
to the GPU box with the repo snapshot.
"""
from __future__ import annotations

import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, 'csrc', 'mvb_api.cu')
DEPS = [SRC, os.path.join(HERE, 'csrc', 'mvb_kernels.cuh'), os.path.join(HERE, 'csrc', 'tc_gemm.cuh'), os.path.join(HERE, 'csrc', 'jacobi_block.cuh'), os.path.join(HERE, 'csrc', 'lanczos_band.cuh'), os.path.join(HERE, 'csrc', 'toeplitz_gram.cuh'), os.path.join(HERE, 'csrc', 'gram_recurrence.cuh'),
        os.path.join(HERE, 'csrc', 'band_loo.cuh'), os.path.join(HERE, 'csrc', 'sweep_chain.cuh'), os.path.join(HERE, 'csrc', 'rf_kernels.cuh'), os.path.join(HERE, 'csrc', 'dense_batched.cuh'),
        os.path.join(HERE, '..', 'include', 'mvb.h')]
OUT = os.path.join(HERE, 'libmvb.so')

NVCC_FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-shared', '--cudart', 'static']


def needs_build() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(d) > t for d in DEPS)


def build(force: bool = False, verbose: bool = True) -> str:
    if not force and not needs_build():
        return OUT
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found: libmvb.so must be built before use (python -m mvpy_b200.build)')
    cmd = [nvcc] + NVCC_FLAGS + ['-o', OUT, SRC]
    if verbose:
        print(' '.join(cmd), flush=True)
    subprocess.check_call(cmd)
    return OUT


if __name__ == '__main__':
    build(force='--force' in sys.argv)
