# -*- coding: utf-8 -*-
"""Debug aid (not a test): runs conv cases through both kernels and prints error patterns."""
import os
import sys
import traceback

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [HERE, os.path.dirname(HERE), os.path.join(os.path.dirname(HERE), 'lfd-a-light-and-fast-detector_b200')]
import torch  # noqa: E402

from gpu_ops import run_conv, ref_conv  # noqa: E402
from lfd import _native as nat  # noqa: E402
from test_gpu_conv import CASES, _make  # noqa: E402


def describe(out, ref):
    o, r = out.float().cpu(), ref.float()
    d = (o - r).abs()
    tol = r.abs() * 2.0 ** -7 + 2e-3 * float(r.abs().max()) * 2.0 ** -7 + 1e-6
    bad = d > tol
    msg = 'max_abs_err %.3e (ref max %.3e) bad %d/%d' % (float(d.max()), float(r.abs().max()), int(bad.sum()), o.numel())
    if bad.any():
        N, H, W, Cc = o.shape
        rows_bad = bad.any(dim=3).float().mean().item()
        ch_bad = bad.reshape(-1, Cc).any(dim=0).float().mean().item()
        idx = torch.nonzero(bad)[:4].tolist()
        msg += ' | pixels with errors %.1f%% channels with errors %.1f%% first %s' % (100 * rows_bad, 100 * ch_bad, idx)
        i = tuple(idx[0])
        msg += ' got %.4f want %.4f' % (float(o[i]), float(r[i]))
        zero = (o == 0).float().mean().item()
        msg += ' zeros in out %.1f%%' % (100 * zero)
    return msg, not bool(bad.any())


def main():
    only = sys.argv[1:] 
    print('device', torch.cuda.get_device_name(0), 'SMs', nat.lib().lfd_device_sm_count())
    for ci, case in enumerate(CASES):
        if only and str(ci) not in only:
            continue
        x, w, scale, shift, res = _make(case)
        N, H, W, Cin, Cout, k, s, relu, use_res, gn = case
        ref = ref_conv(x, w, scale, shift, s, relu, res=res)
        for impl, nm in ((nat.CONV_SIMT, 'simt'), (nat.CONV_UMMA, 'umma')):
            try:
                out, stats, q = run_conv(x, w, scale, shift, s, relu, res=res, gn_groups=gn, impl=impl)
                msg, ok = describe(out, ref)
                print('[%2d] %s %-4s %s  %s  plan=%s' % (ci, 'OK  ' if ok else 'FAIL', nm, case, msg, q if nm == 'umma' else ''))
            except Exception as e:
                print('[%2d] EXC  %-4s %s %s' % (ci, nm, case, repr(e)[:300]))
                traceback.print_exc()
                return 1
        sys.stdout.flush()
    return 0


if __name__ == '__main__':
    sys.exit(main())
