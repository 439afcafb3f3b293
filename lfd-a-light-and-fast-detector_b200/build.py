# -*- coding: utf-8 -*-
"""Builds liblfd_b200.so (sm_100a only) in-tree with nvcc.  No GPU is needed to compile.

    python build.py [--force] [--verbose]

LFD_B200_TRACE=1 compiles the clock64() per-role hooks of the convolution kernel in (tests/debug_trace.py) and
LFD_B200_TIMELINE=1 the per-launch %globaltimer stamps (tests/debug_timeline.py); both are absent from the normal build.
"""
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, 'csrc')
OUT = os.environ.get('LFD_B200_OUT') or os.path.join(HERE, 'liblfd_b200.so')      # LFD_B200_OUT / LFD_B200_EXTRA_FLAGS: tuning experiments only
SOURCES = ['api.cu', 'conv_umma.cu', 'conv_simt.cu', 'postprocess.cu', 'losses.cu', 'train.cu', 'wgrad_umma.cu']
HEADERS = ['ptx.cuh', 'conv_common.cuh', 'kernels.cuh', 'train.cuh', os.path.join('..', '..', 'include', 'lfd_b200.h')]
NVCC = os.environ.get('NVCC', '/usr/local/cuda/bin/nvcc')
FLAGS = ['-gencode', 'arch=compute_100a,code=sm_100a', '-O3', '-lineinfo', '-std=c++17', '-Xcompiler', '-fPIC',
         '--expt-relaxed-constexpr'] + (['-DLFD_B200_TRACE'] if os.environ.get('LFD_B200_TRACE') else []) + \
        (['-DLFD_B200_TIMELINE'] if os.environ.get('LFD_B200_TIMELINE') else []) + os.environ.get('LFD_B200_EXTRA_FLAGS', '').split()


def _stale():
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    deps = [os.path.join(CSRC, s) for s in SOURCES + HEADERS] + [os.path.abspath(__file__)]
    return any(os.path.getmtime(d) > t for d in deps)


def _check_stack_frames(ptxas_log, limit=64):
    """The warp-specialised conv kernel keeps everything in registers; a large stack frame means a role lambda was not
    inlined and its closure lives in local memory (measured: the stem kernel 2.5x slower).  Fail the build instead."""
    import re
    name = None
    for line in ptxas_log.splitlines():
        m = re.search(r'Function properties for (\S+)', line)
        if m:
            name = m.group(1)
        m = re.search(r'(\d+) bytes stack frame', line)
        if m and name and 'conv_umma_kernel' in name and int(m.group(1)) > limit:
            raise RuntimeError('%s has a %s-byte stack frame (limit %d): a role lambda is no longer inlined' % (name, m.group(1), limit))


def build(force=False, verbose=False):
    if not force and not _stale():
        return OUT
    objs = []
    procs = []
    for s in SOURCES:
        o = os.path.join(CSRC, s.replace('.cu', '.o')) if not os.environ.get('LFD_B200_OUT') else os.path.join('/tmp', os.path.basename(OUT) + '.' + s.replace('.cu', '.o'))
        # conv_umma.cu is always compiled with ptxas -v: see _check_stack_frames
        cmd = [NVCC] + FLAGS + (['-Xptxas', '-v'] if (verbose or s == 'conv_umma.cu') else []) + ['-c', os.path.join(CSRC, s), '-o', o]
        procs.append((s, subprocess.Popen(cmd, stdout=subprocess.PIPE, stderr=subprocess.STDOUT)))
        objs.append(o)
    for s, p in procs:
        out = p.communicate()[0].decode()
        if p.returncode != 0:
            raise RuntimeError('nvcc failed on %s:\n%s' % (s, out))
        if s == 'conv_umma.cu':
            _check_stack_frames(out)
            if not verbose:
                out = '\n'.join(l for l in out.splitlines() if 'ptxas info' not in l and 'bytes stack frame' not in l)
        if verbose or out.strip():
            sys.stderr.write(out)
    subprocess.check_call([NVCC, '-shared', '-o', OUT] + objs + ['-gencode', 'arch=compute_100a,code=sm_100a', '-lcudart'])
    return OUT


if __name__ == '__main__':
    print(build(force='--force' in sys.argv, verbose='--verbose' in sys.argv))
