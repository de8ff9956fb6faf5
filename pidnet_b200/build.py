"""In-tree build of libpidnet_b200.so (hand-written sm_100a CUDA behind a C ABI).

nvcc cross-compiles without a GPU; the .so is git-ignored but travels with the repo snapshot to the
GPU box.  `python -m pidnet_b200.build` rebuilds when a source is newer than the library.
"""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
LIB = os.path.join(HERE, 'lib', 'libpidnet_b200.so')
PROBE_LIB = os.path.join(HERE, 'lib', 'libpidnet_b200_probe.so')
SOURCES = ['conv_tc.cu', 'conv3_ws.cu', 'stem_tc.cu', 'stem2_tc.cu', 'criterion.cu', 'train_kernels.cu', 'wgrad_tc.cu', 'kernels.cu', 'engine.cu']
HEADERS = ['conv_tc.cuh', 'kernels.cuh', 'ptx.cuh', 'criterion.cuh', 'train_kernels.cuh', 'train.inc', os.path.join('..', '..', 'include', 'pidnet_b200.h')]


def _nvcc():
    for cand in (os.environ.get('NVCC'), '/usr/local/cuda/bin/nvcc', 'nvcc'):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    return 'nvcc'


def needs_build():
    if not os.path.exists(LIB):
        return True
    t = os.path.getmtime(LIB)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.exists(d) and os.path.getmtime(d) > t for d in deps)


def build(force=False, verbose=False, probes=False):
    """probes=True builds libpidnet_b200_probe.so instead: the same sources + csrc/probe.cu with -DPIDNET_PROBES (hardware
    probes for tools/probe_*.py; the product library never contains them)."""
    lib = PROBE_LIB if probes else LIB
    if not force and not probes and not needs_build():
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    objs = []
    common = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17',
              '-Xcompiler', '-fPIC', '-Xcompiler', '-fvisibility=hidden']
    if probes:
        common += ['-DPIDNET_PROBES']
    if verbose:
        common += ['-Xptxas', '-v']
    procs = []
    for s in SOURCES + (['probe.cu'] if probes else []):
        o = os.path.join(HERE, 'lib', s.replace('.cu', '.probe.o' if probes else '.o'))
        objs.append(o)
        cmd = [_nvcc()] + common + ['-c', os.path.join(CSRC, s), '-o', o]
        procs.append((cmd, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)))
    for cmd, p in procs:
        out, _ = p.communicate()
        if verbose or p.returncode != 0:
            sys.stderr.write(out)
        if p.returncode != 0:
            raise RuntimeError('nvcc failed: ' + ' '.join(cmd))
    cmd = [_nvcc(), '-shared', '-o', lib] + objs + ['-cudart', 'static', '-Xcompiler', '-fPIC']
    r = subprocess.run(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout)
        raise RuntimeError('link failed: ' + ' '.join(cmd))
    return lib


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='-v' in sys.argv, probes='--probes' in sys.argv))
