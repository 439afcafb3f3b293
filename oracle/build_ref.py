# -*- coding: utf-8 -*-
"""Recipe: compile the REFERENCE's own CPU NMS (lfd/model/utils/build/nms/src/{nms_ext.cpp,cpu/nms_cpu.cpp}) from the
sources where they lie under /root/reference into oracle/_ref/nms_ext_ref.so (git-ignored, travels to the GPU box).
Test infrastructure only: used to validate oracle/lfd_oracle.py::nms and as part of the CPU baseline.
The reference's two .cu files include the removed THC/THC.h and do not build against torch >= 1.11.

    python oracle/build_ref.py
"""
import importlib.util
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get('LFD_REFERENCE_ROOT', '/root/reference')
OUT_DIR = os.path.join(HERE, '_ref')
NAME = 'nms_ext_ref'


def so_path():
    return os.path.join(OUT_DIR, NAME + '.so')


def build(verbose=False):
    src = os.path.join(REF, 'lfd', 'model', 'utils', 'build', 'nms', 'src')
    if not os.path.isdir(src):
        return None
    if os.path.exists(so_path()):
        return so_path()
    from torch.utils.cpp_extension import load
    os.makedirs(OUT_DIR, exist_ok=True)
    load(name=NAME, sources=[os.path.join(src, 'nms_ext.cpp'), os.path.join(src, 'cpu', 'nms_cpu.cpp')],
         build_directory=OUT_DIR, verbose=verbose, extra_cflags=['-O2'])
    return so_path() if os.path.exists(so_path()) else None


def load_module():
    """-> the compiled reference module exposing nms(dets, thr), soft_nms(...), nms_match(...) or None."""
    import torch  # noqa: F401  (the extension links against libtorch)
    if not os.path.exists(so_path()):
        return None
    spec = importlib.util.spec_from_file_location(NAME, so_path())
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod


if __name__ == '__main__':
    p = build(verbose='-v' in sys.argv)
    print(p if p else 'reference sources not found under %s' % REF)
