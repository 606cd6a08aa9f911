"""GPU parity of the two convolution kernel families on single layers, against torch conv1d on the CPU."""
import ctypes

import numpy as np
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu


def _run(tc, x, w, b, dil, pre_slope=1.0, residual=None, slope=1.0, div=1.0, want_b=True):
    from infernos_b200 import _lib
    lib = _lib.load()
    W, T, Cin = x.shape
    Cout, _, k = w.shape
    out32 = torch.empty(W, T, Cout, device="cuda", dtype=torch.float32)
    outb = torch.empty(W, T, Cout, device="cuda", dtype=torch.bfloat16) if want_b else None
    wc, bc = w.contiguous(), b.contiguous()
    res = residual.cuda().contiguous() if residual is not None else None
    st = torch.cuda.current_stream().cuda_stream
    if tc:
        xin = x.to(torch.bfloat16).cuda().contiguous()
        rc = lib.b2_conv1d_tc(xin.data_ptr(), wc.data_ptr(), bc.data_ptr(), W, T, Cin, Cout, k, dil,
                              res.data_ptr() if res is not None else None, out32.data_ptr(),
                              outb.data_ptr() if outb is not None else None, slope, div, st)
    else:
        xin = x.cuda().contiguous()
        rc = lib.b2_conv1d_f32(xin.data_ptr(), wc.data_ptr(), bc.data_ptr(), W, T, Cin, Cout, k, dil, pre_slope,
                               res.data_ptr() if res is not None else None, out32.data_ptr(),
                               outb.data_ptr() if outb is not None else None, slope, div, st)
    _lib.check(rc, "conv1d")
    torch.cuda.synchronize()
    return out32.cpu(), (outb.float().cpu() if outb is not None else None)


def _ref(x, w, b, dil, pre_slope=1.0, residual=None, div=1.0):
    k = w.shape[2]
    xin = F.leaky_relu(x, pre_slope) if pre_slope != 1.0 else x
    y = F.conv1d(xin.transpose(1, 2).double(), w.double(), b.double(), dilation=dil, padding=(k - 1) * dil // 2).transpose(1, 2)
    if residual is not None:
        y = y + residual.double()
    return (y / div).float()


SHAPES = [
    # (W, T, Cin, Cout, k, dil)
    (3, 3072, 32, 32, 11, 5), (2, 768, 64, 64, 7, 3), (3, 192, 128, 128, 3, 1), (5, 48, 256, 256, 11, 5),
    (5, 48, 256, 256, 3, 1), (20, 12, 512, 1024, 3, 1), (3, 48, 256, 512, 3, 1), (2, 192, 128, 256, 3, 1),
    (2, 768, 64, 128, 3, 1), (1, 3072, 32, 32, 3, 1), (2, 20, 256, 256, 7, 1), (1, 300, 64, 64, 11, 3),
    # C = 256 runs as CTA pairs (cta_group::2): an odd tile count (19: the last pair's second tile is a dummy), and more tiles than CTAs
    # (310 tiles on 148 CTAs: three passes over the rings, accumulators and A buffers, the last one with 134 dummy tiles)
    (37, 48, 256, 256, 7, 3), (620, 48, 256, 256, 3, 1), (301, 48, 256, 256, 11, 5),
]


@pytest.mark.parametrize("shape", SHAPES)
def test_simt_conv_matches_torch(shape):
    W, T, Cin, Cout, k, dil = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.randn(W, T, Cin, generator=g)
    w = torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5
    b = torch.randn(Cout, generator=g) * 0.1
    r = torch.randn(W, T, Cout, generator=g)
    out, outb = _run(False, x, w, b, dil, pre_slope=0.1, residual=r, slope=0.1, div=3.0)
    ref = _ref(x, w, b, dil, pre_slope=0.1, residual=r, div=3.0)
    assert (out - ref).abs().max() < 2e-5
    assert (outb - F.leaky_relu(ref, 0.1)).abs().max() < 2e-2


@pytest.mark.parametrize("shape", SHAPES)
def test_tensor_core_conv_matches_torch_on_bf16_operands(shape):
    W, T, Cin, Cout, k, dil = shape
    g = torch.Generator().manual_seed(sum(shape) + 1)
    x = torch.randn(W, T, Cin, generator=g).bfloat16().float()
    w = (torch.randn(Cout, Cin, k, generator=g) / (Cin * k) ** 0.5)
    b = torch.randn(Cout, generator=g) * 0.1
    r = torch.randn(W, T, Cout, generator=g)
    out, outb = _run(True, x, w, b, dil, residual=r, slope=0.1, div=1.0)
    ref = _ref(x, w.bfloat16().float(), b, dil, residual=r)
    err = (out - ref).abs().max().item()
    assert err < 1e-4, f"max abs err {err}"          # fp32 accumulation of exact bf16 products
    assert (outb - F.leaky_relu(ref, 0.1)).abs().max() < 3e-2
