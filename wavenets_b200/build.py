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
# flavours: '' = the product; 'precise' = same sources with -DWN_PRECISE_MATH (accurate tanh / sigmoid in the bf16 tier's gates
# instead of MUFU.TANH) — test infrastructure for the bf16-faithful parity tests, loaded with WN_LIB=<path>
FLAVOURS = {'': [], 'precise': ['-DWN_PRECISE_MATH'], 'tl': ['-DTC_TIMELINE']}   # 'tl': in-kernel clock64 phase accounting (printf), diagnostics only

SOURCES = ['wn_api.cu']
HEADERS = ['common.cuh', 'generate.cuh', 'epilogues.cuh', 'gemm_simt.cuh', 'gemm_tc.cuh', 'gemm_tc_block.cuh', 'gemm_tc_wgroup.cuh', 'gemm_tc_stack.cuh', 'gemm_tc_stack_bwd.cuh', 'tc_common.cuh', 'tc_epilogues.cuh', 'kernels_misc.cuh', 'nccl_dl.cuh',
           os.path.join('..', '..', 'include', 'wavenet_b200.h')]

NVCC_FLAGS = [
  '-gencode', 'arch=compute_100a,code=sm_100a', '-lineinfo', '-O3', '-std=c++17',
  '-Xcompiler', '-fPIC', '-shared', '--expt-relaxed-constexpr',
  '-Xptxas', '-v', '-ldl',
]


def _flags(flavour: str = ''):
  return NVCC_FLAGS + FLAVOURS[flavour] + os.environ.get('WN_NVCC_EXTRA', '').split()


def lib_path(flavour: str = '') -> str:
  return LIB if not flavour else os.path.join(HERE, f'libwavenet_b200_{flavour}.so')


def _digest(flavour: str = '') -> str:
  h = hashlib.sha256()
  for f in SOURCES + HEADERS:
    with open(os.path.join(CSRC, f), 'rb') as fh:
      h.update(fh.read())
  h.update(' '.join(_flags(flavour)).encode())
  return h.hexdigest()


def build(force: bool = False, verbose: bool = False, flavour: str = '') -> str:
  dig = _digest(flavour)
  LIB = lib_path(flavour)
  STAMP = os.path.join(HERE, f'.libwavenet_b200{"_" + flavour if flavour else ""}.stamp')
  if not force and os.path.exists(LIB) and os.path.exists(STAMP):
    with open(STAMP) as fh:
      if fh.read().strip() == dig:
        return LIB
  nvcc = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
  cmd = [nvcc] + _flags(flavour) + [os.path.join(CSRC, s) for s in SOURCES] + ['-o', LIB]
  res = subprocess.run(cmd, capture_output=True, text=True)
  log = os.path.join(HERE, f'build{"_" + flavour if flavour else ""}.log')
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
  flav = [a for a in sys.argv[1:] if not a.startswith('--')]
  print(build(force='--force' in sys.argv, verbose=True, flavour=flav[0] if flav else ''))
