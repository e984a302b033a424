"""Build recipe for libssdcodec.so (hand-written sm_100a CUDA kernels + C ABI).

    python -m jpeg_detection_resnet_ssd_b200.build [--force] [--verbose]

nvcc cross-compiles for sm_100a without a GPU.  The library is built IN-TREE
(`jpeg_detection_resnet_ssd_b200/lib/libssdcodec.so`, git-ignored) so that it
travels with a snapshot of the repository.

Flags that matter:
  --fmad=false   numpy never contracts a*b+c into an FMA; the kernels must not
                 either or box coordinates / IoU decisions drift from the
                 reference (hard part 1).
  -lineinfo      so `ncu --import-source on` maps stalls to source lines.
"""
from __future__ import annotations

import hashlib
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIBDIR = os.path.join(HERE, 'lib')
LIBPATH = os.path.join(LIBDIR, 'libssdcodec.so')
STAMP = os.path.join(LIBDIR, 'libssdcodec.stamp')
INCLUDE = os.path.join(os.path.dirname(HERE), 'include')

SOURCES = ['ctx.cu', 'decode.cu', 'encode.cu', 'thin.cu', 'voc.cu', 'loss.cu', 'evalprep.cu']
NVCC_FLAGS = [
    '-O3', '-std=c++17',
    '-gencode', 'arch=compute_100a,code=sm_100a',
    '-lineinfo', '--fmad=false',
    '-Xcompiler', '-fPIC',
    '--expt-relaxed-constexpr',
]


def _nvcc():
    for cand in (os.environ.get('NVCC'), shutil.which('nvcc'), '/usr/local/cuda/bin/nvcc'):
        if cand and os.path.exists(cand):
            return cand
    raise RuntimeError('nvcc not found; libssdcodec cannot be built (there is no CPU fallback)')


def _source_hash():
    h = hashlib.sha256()
    files = [os.path.join(CSRC, f) for f in sorted(os.listdir(CSRC))] + [os.path.join(INCLUDE, 'ssdcodec.h')]
    for f in files:
        with open(f, 'rb') as fh:
            h.update(os.path.basename(f).encode())      # (not the absolute path: a snapshot of the tree elsewhere keeps its stamp)
            h.update(fh.read())
    h.update(' '.join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_current():
    if not (os.path.exists(LIBPATH) and os.path.exists(STAMP)):
        return False
    try:
        with open(STAMP) as fh:
            return fh.read().strip() == _source_hash()
    except OSError:
        return False


def build(force=False, verbose=False):
    """Compile every .cu under csrc/ for sm_100a into lib/libssdcodec.so."""
    if not force and is_current():
        return LIBPATH
    os.makedirs(LIBDIR, exist_ok=True)
    nvcc = _nvcc()
    objs = []
    procs = []
    for src in SOURCES:
        obj = os.path.join(LIBDIR, src.replace('.cu', '.o'))
        cmd = [nvcc] + NVCC_FLAGS + ['-I', INCLUDE, '-c', os.path.join(CSRC, src), '-o', obj]
        if verbose:
            cmd.insert(1, '-Xptxas')
            cmd.insert(2, '-v')
            print(' '.join(cmd), flush=True)
        procs.append((src, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
        objs.append(obj)
    failed = False
    for src, p in procs:
        out, _ = p.communicate()
        if p.returncode != 0 or verbose:
            sys.stderr.write('[%s]\n%s\n' % (src, out))
        failed |= p.returncode != 0
    if failed:
        raise RuntimeError('nvcc failed; see output above')
    cmd = [nvcc, '-shared', '-gencode', 'arch=compute_100a,code=sm_100a', '-o', LIBPATH] + objs
    subprocess.run(cmd, check=True)
    with open(STAMP, 'w') as fh:
        fh.write(_source_hash())
    return LIBPATH


if __name__ == '__main__':
    path = build(force='--force' in sys.argv, verbose='--verbose' in sys.argv)
    print(path)
