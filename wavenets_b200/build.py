"""In-tree build of libwavenet_b200.so (nvcc, sm_100a only).

`python -m wavenets_b200.build` or `__graft_entry__.build()`.  nvcc cross-compiles without a
GPU.  The .so is git-ignored but travels with the repo snapshot to the GPU box.
"""
from __future__ import annotations

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'libwavenet_b200.so')
STAMP = os.path.join(HERE, '.libwavenet_b200.stamp')

SOURCES = ['wn_api.cu']
HEADERS = ['common.cuh', 'generate.cuh', 'epilogues.cuh', 'gemm_simt.cuh', 'gemm_tc.cuh', 'gemm_tc_block.cuh', 'gemm_tc_wgroup.cuh', 'gemm_tc_stack.cuh', 'tc_common.cuh', 'tc_epilogues.cuh', 'kernels_misc.cuh', 'nccl_dl.cuh',
           os.path.join('..', '..', 'include', 'wavenet_b200.h')]

NVCC_FLAGS = [
  '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
  '-Xcompiler', '-fPIC', '-shared', '--expt-relaxed-constexpr',
  '-Xptxas', '-v', '-ldl',
]


def _flags():
  return NVCC_FLAGS + os.environ.get('WN_NVCC_EXTRA', '').split()


def _digest() -> str:
  h = hashlib.sha256()
  for f in SOURCES + HEADERS:
    with open(os.path.join(CSRC, f), 'rb') as fh:
      h.update(fh.read())
  h.update(' '.join(_flags()).encode())
  return h.hexdigest()


def build(force: bool = False, verbose: bool = False) -> str:
  dig = _digest()
  if not force and os.path.exists(LIB) and os.path.exists(STAMP):
    with open(STAMP) as fh:
      if fh.read().strip() == dig:
        return LIB
  nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
  cmd = [nvcc] + _flags() + [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB]
  res = subprocess.run(cmd, capture_output=True, text=True)
  log = os.path.join(HERE, 'build.log')
  with open(log, 'w') as fh:
    fh.write(' '.join(cmd) + '\n' + res.stdout + res.stderr)
  if res.returncode != 0:
    sys.stderr.write(res.stdout[-4000:] + res.stderr[-8000:])
    raise RuntimeError('nvcc failed building libwavenet_b200.so (see %s)' % log)
  if verbose:
    print(res.stderr[-3000:])
  with open(STAMP, 'w') as fh:
    fh.write(dig)
  return LIB


if __name__ == '__main__':
  print(build(force='--force' in sys.argv, verbose=True))
