"""Builds libhsgb200.so in-tree with nvcc for sm_100a (cross-compiles without a GPU).

    python -m hsg_b200.build [--force] [--verbose]

The shared library lands next to this file (git-ignored, but it travels with the
gpurun snapshot).  Nothing here JIT-compiles at import time: hsg_b200/_lib.py only
dlopens the file this script produced and fails loudly when it is missing.
"""

import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.path.join(HERE, 'libhsgb200.so')
STAMP = os.path.join(HERE, 'build', 'stamp')

SOURCES = ['abi.cu', 'prep.cu', 'segreduce.cu', 'kmeans.cu', 'tc_estep.cu', 'nce.cu', 'nce_tc.cu', 'gemm_tc.cu', 'relabel.cu', 'attention.cu', 'attention_tc.cu', 'attention_bwd_tc.cu', 'graph.cu', 'topk.cu', 'syncbn.cu', 'exchange.cu']
NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden', '--expt-relaxed-constexpr',
]


def nvcc():
  for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
    if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
      return cand
  return 'nvcc'


def _digest():
  h = hashlib.sha256()
  names = sorted(os.listdir(CSRC)) + ['../../include/hsg_b200.h']
  for name in names:
    path = os.path.join(CSRC, name)
    if os.path.isfile(path):
      h.update(name.encode())
      with open(path, 'rb') as f:
        h.update(f.read())
  h.update(' '.join(NVCC_FLAGS).encode())
  return h.hexdigest()


def build(force=False, verbose=False):
  digest = _digest()
  if not force and os.path.exists(OUT) and os.path.exists(STAMP):
    with open(STAMP) as f:
      if f.read().strip() == digest:
        return OUT
  objdir = os.path.join(HERE, 'build')
  os.makedirs(objdir, exist_ok=True)
  procs = []
  objs = []
  for src in SOURCES:
    obj = os.path.join(objdir, src.replace('.cu', '.o'))
    objs.append(obj)
    cmd = [nvcc()] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + [
        '-c', os.path.join(CSRC, src), '-o', obj]
    procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
  failed = False
  for src, p in procs:
    out, _ = p.communicate()
    if p.returncode != 0 or verbose:
      sys.stderr.write('---- %s\n%s\n' % (src, out))
    failed = failed or p.returncode != 0
  if failed:
    raise RuntimeError('nvcc failed building libhsgb200.so')
  link = [nvcc(), '-shared', '-o', OUT] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                 '-lcudart_static', '-ldl', '-lrt', '-lpthread']
  r = subprocess.run(link, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
  if r.returncode != 0:
    sys.stderr.write(r.stdout)
    raise RuntimeError('link failed for libhsgb200.so')
  with open(STAMP, 'w') as f:
    f.write(digest)
  return OUT


if __name__ == '__main__':
  path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
  print(path)
