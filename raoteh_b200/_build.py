"""Build librt_b200.so in-tree with nvcc for sm_100a (no torch types involved).

The shared library is the C-ABI drop-in boundary (include/rt_b200.h); it links
the static CUDA runtime, so it only needs the NVIDIA driver at run time.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'librt_b200.so')
STAMP = os.path.join(HERE, 'csrc', '.build_stamp')

NVCC_FLAGS = [
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-O3', '-lineinfo', '-std=c++17',
    '-Xcompiler', '-fPIC',
    '--expt-relaxed-constexpr',
] + os.environ.get('RT_EXTRA_NVCC_FLAGS', '').split()


def sources():
    return sorted(os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith('.cu'))


def _digest():
    h = hashlib.sha256()
    for f in sorted(os.listdir(CSRC)):
        if f.endswith(('.cu', '.cuh', '.h')):
            h.update(open(os.path.join(CSRC, f), 'rb').read())
    h.update(open(os.path.join(HERE, '..', 'include', 'rt_b200.h'), 'rb').read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def find_nvcc():
    nvcc = shutil.which('nvcc') or '/usr/local/cuda/bin/nvcc'
    if not os.path.exists(nvcc):
        raise RuntimeError('nvcc not found; librt_b200.so cannot be built')
    return nvcc


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ and link librt_b200.so (idempotent)."""
    dig = _digest()
    if not force and os.path.exists(LIB) and os.path.exists(STAMP) and open(STAMP).read() == dig:
        return LIB
    nvcc = find_nvcc()
    objdir = os.path.join(HERE, 'csrc', 'build')
    os.makedirs(objdir, exist_ok=True)
    objs = []
    procs = []
    for src in sources():
        obj = os.path.join(objdir, os.path.basename(src)[:-3] + '.o')
        cmd = [nvcc] + NVCC_FLAGS + (['-Xptxas', '-v'] if verbose else []) + ['-c', src, '-o', obj]
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            failed = True
            sys.stderr.write('nvcc failed for %s:\n%s\n' % (src, out))
        elif verbose:
            sys.stderr.write(out)
    if failed:
        raise RuntimeError('nvcc compilation failed')
    cmd = [nvcc, '-shared', '-o', LIB] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a',
                                                 '-Xcompiler', '-fPIC']
    subprocess.check_call(cmd)
    with open(STAMP, 'w') as f:
        f.write(dig)
    return LIB


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv))
