# -*- coding: utf-8 -*-
"""Helpers that drive single native ops through the C-ABI (lfd_run_op) for the GPU parity tests."""
import ctypes as C

import torch
import torch.nn.functional as F

from lfd import _native as nat
from lfd._engine import pack_conv_weight, fold_scale


def bf16r(t):
    return t.to(torch.bfloat16).float()


def fp16r(t):
    return t.to(torch.float16).float()


# act dtype name -> (torch storage type, rounding function, relative size of one ulp step, native code)
DTYPES = {'bf16': (torch.bfloat16, bf16r, 2.0 ** -7, nat.DTYPE_BF16), 'fp16': (torch.float16, fp16r, 2.0 ** -10, nat.DTYPE_FP16)}


def conv_out(size, k, s):
    return (size + 2 * (k // 2) - k) // s + 1


def run_conv(x_nhwc, weight, scale, shift, stride, relu, res=None, gn_groups=0, impl=nat.CONV_UMMA, tail=None, dtype='bf16'):
    """x_nhwc: cuda bf16 / fp16 [N,H,W,Cin]; weight fp32 [Cout,Cin,k,k] (already representable in the 16-bit type).
    -> (out bf16 [N,Ho,Wo,Cout], stats double [N,groups,2] or None)"""
    tdt, _, _, code = DTYPES[dtype]
    assert x_nhwc.dtype == tdt
    dev = x_nhwc.device
    N, H, W, Cin = x_nhwc.shape
    Cout, _, k, _ = weight.shape
    Ho, Wo = conv_out(H, k, stride), conv_out(W, k, stride)
    Cf = tail[0].shape[0] if tail is not None else Cout
    q = nat.conv_query(N, H, W, Cin, Ho, Wo, Cout, k, stride, Cf if tail is not None else 0)
    wp = pack_conv_weight(fold_scale(weight, scale), q['cc'], tdt).to(dev)   # BatchNorm scale folded before the bf16 rounding
    sh = shift.float().to(dev).contiguous()
    in_b = x_nhwc.numel() * 2
    out_b = N * Ho * Wo * Cf * 2
    al = lambda v: (v + 255) & ~255
    off_in, off_out = 4096, 4096 + al(in_b)
    off_res = off_out + al(out_b)
    total = off_res + al(out_b) + 256
    ws = torch.zeros(total, dtype=torch.uint8, device=dev)
    ws[off_in:off_in + in_b] = x_nhwc.contiguous().view(torch.uint8).reshape(-1)
    if res is not None:
        ws[off_res:off_res + out_b] = res.contiguous().view(torch.uint8).reshape(-1)
    op = nat.Op()
    op.kind = nat.OP_CONV
    op.dtype = code
    op.N, op.H, op.W, op.Cin, op.Ho, op.Wo, op.Cout = N, H, W, Cin, Ho, Wo, Cout
    op.ksize, op.stride, op.relu, op.gn_groups, op.cc = k, stride, int(relu), gn_groups, q['cc']
    op.in_off, op.out_off, op.res_off = off_in, off_out, (off_res if res is not None else -1)
    op.stats_off = 0 if gn_groups else -1
    op.weight, op.shift = wp.data_ptr(), sh.data_ptr()
    if tail is not None:
        w2, sc2, sh2, relu2 = tail
        w2p = pack_conv_weight(fold_scale(w2, sc2), Cout, tdt).to(dev)
        sh2d = sh2.float().to(dev).contiguous()
        op.tail_cout, op.tail_relu = Cf, int(relu2)
        op.tail_weight, op.tail_shift = w2p.data_ptr(), sh2d.data_ptr()
    with torch.cuda.device(dev):
        nat.check(nat.lib().lfd_run_op(C.byref(op), None, 0, nat.ptr(ws), None, None, 0, 0, impl, nat.stream_ptr()))
        torch.cuda.synchronize()
    out = ws[off_out:off_out + out_b].view(tdt).view(N, Ho, Wo, Cf).clone()
    stats = ws[0:N * gn_groups * 16].view(torch.float64).view(N, gn_groups, 2).clone() if gn_groups else None
    return out, stats, q


def ref_conv(x_nhwc, weight, scale, shift, stride, relu, res=None, dtype='bf16'):
    """fp32 CPU reference of the fused layer on the operands the kernel sees: weights = round16(weight * scale) (BatchNorm
    fold, rounding point Rw), shift rounded to the 16-bit type; result NOT yet rounded."""
    rnd = DTYPES[dtype][1]
    x = x_nhwc.float().cpu().permute(0, 3, 1, 2)
    k = weight.shape[-1]
    y = F.conv2d(x, rnd(fold_scale(weight, scale)), None, stride=stride, padding=k // 2)
    y = y + rnd(shift.float().cpu())[None, :, None, None]
    if res is not None:
        y = y + res.float().cpu().permute(0, 3, 1, 2)
    if relu:
        y = F.relu(y)
    return y.permute(0, 2, 3, 1).contiguous()


def assert_bf16_close(out_bf16, ref_fp32, what='', dtype='bf16'):
    """out must equal the fp32 reference rounded to the 16-bit type up to 1 ulp (accumulation-order noise next to a rounding
    boundary) -- Gate A of the parity protocol."""
    ulp = DTYPES[dtype][2]
    o = out_bf16.float().cpu()
    r = ref_fp32.float()
    tol = r.abs() * ulp + 2e-3 * float(r.abs().max()) * ulp + 1e-6
    bad = (o - r).abs() > tol
    if bool(bad.any()):
        idx = torch.nonzero(bad)
        i = tuple(idx[0].tolist())
        raise AssertionError('%s: %d / %d elements off; first at %s: got %g want %g; max abs err %g (ref max %g)'
                             % (what, int(bad.sum()), o.numel(), i, float(o[i]), float(r[i]), float((o - r).abs().max()), float(r.abs().max())))
