# -*- coding: utf-8 -*-
"""GPU parity (Gate A): the tcgen05 implicit-GEMM convolution (and its SIMT cross-check) against an fp32 CPU
convolution of the same bf16 operands, through the C-ABI (lfd_run_op)."""
import pytest
import torch

from gpu_ops import bf16r, run_conv, ref_conv, assert_bf16_close
from lfd import _native as nat

pytestmark = pytest.mark.gpu

# (N, H, W, Cin, Cout, k, stride, relu, residual, gn)
CASES = [
    (1, 16, 8, 64, 64, 1, 1, True, False, 0),      # one full flat tile
    (2, 23, 31, 64, 64, 1, 1, True, False, 0),     # ragged flat tiles
    (2, 23, 31, 64, 128, 1, 1, True, False, 0),    # neck shape
    (2, 12, 20, 128, 128, 1, 1, False, False, 16), # head tower + GroupNorm statistics
    (1, 45, 80, 32, 32, 1, 1, True, False, 0),     # XS stem 1x1
    (1, 16, 8, 64, 64, 3, 1, True, False, 0),      # exactly one 16x8 tile
    (2, 23, 40, 64, 64, 3, 1, True, True, 0),      # 720p stage-2 shape, residual
    (1, 37, 29, 64, 64, 3, 1, False, False, 0),
    (2, 12, 20, 128, 128, 3, 1, True, True, 0),    # streamed weights
    (1, 45, 80, 64, 64, 3, 2, True, False, 0),     # stride 2 (odd output size 23x40)
    (2, 23, 40, 64, 128, 3, 2, True, False, 0),    # stage-3 entry
    (1, 46, 62, 32, 32, 3, 2, True, False, 0),     # XS stem 3x3/s2
    (2, 23, 40, 64, 128, 1, 2, False, False, 0),   # downsample path
    (1, 45, 80, 64, 64, 1, 2, False, False, 0),
]


def _make(case, seed=0):
    N, H, W, Cin, Cout, k, s, relu, use_res, gn = case
    g = torch.Generator().manual_seed(seed)
    x = bf16r(torch.randn((N, H, W, Cin), generator=g)).to(torch.bfloat16).cuda()
    w = bf16r(torch.randn((Cout, Cin, k, k), generator=g) * (2.0 / (Cin * k * k)) ** 0.5)
    scale = torch.rand((Cout,), generator=g) + 0.5
    shift = torch.randn((Cout,), generator=g) * 0.2
    if gn:
        scale, shift = torch.ones(Cout), torch.zeros(Cout)
    Ho, Wo = (H + 2 * (k // 2) - k) // s + 1, (W + 2 * (k // 2) - k) // s + 1
    res = bf16r(torch.randn((N, Ho, Wo, Cout), generator=g)).to(torch.bfloat16).cuda() if use_res else None
    return x, w, scale, shift, res


@pytest.mark.parametrize('impl', [nat.CONV_SIMT, nat.CONV_UMMA], ids=['simt', 'umma'])
@pytest.mark.parametrize('case', CASES, ids=lambda c: 'N%d_%dx%d_%d-%d_k%ds%d_r%d_res%d_gn%d' % c)
def test_conv_matches_fp32_reference(case, impl):
    N, H, W, Cin, Cout, k, s, relu, use_res, gn = case
    x, w, scale, shift, res = _make(case)
    out, stats, q = run_conv(x, w, scale, shift, s, relu, res=res, gn_groups=gn, impl=impl)
    ref = ref_conv(x, w, scale, shift, s, relu, res=res)
    assert_bf16_close(out, ref, 'conv %s (plan %s)' % (case, q))
    if gn:
        o = out.float().cpu().reshape(N, -1, gn, Cout // gn).double()
        s1, s2 = o.sum(dim=(1, 3)), (o * o).sum(dim=(1, 3))
        assert torch.allclose(stats[..., 0].cpu(), s1, rtol=1e-6, atol=1e-3)
        assert torch.allclose(stats[..., 1].cpu(), s2, rtol=1e-6, atol=1e-3)


def test_conv_large_grid_persistent_loop():
    """More tiles than SMs: exercises the persistent tile loop, both accumulator stages and ring wrap-around."""
    case = (3, 90, 160, 64, 64, 3, 1, True, True, 0)
    x, w, scale, shift, res = _make(case, seed=3)
    out, _, q = run_conv(x, w, scale, shift, 1, True, res=res)
    assert q['num_tiles'] > 2 * nat.lib().lfd_device_sm_count()
    assert_bf16_close(out, ref_conv(x, w, scale, shift, 1, True, res=res), 'large conv')
