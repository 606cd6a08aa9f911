"""Host logic of the stacked-output ResBlock kernel (csrc/conv_resblock_t.cu), on the CPU: the library's own slab plan (b2_resblock_t_plan, no
device needed) is fed to the index-exact numpy model of the kernel (tools/resblock_t_model.py: operand mappings, extended-tap loop, slot
window, accumulator layouts, halo), whose result must equal torch.conv1d's HiFiGAN ResBlock (modeling_speecht5.py:2954-2962) to rounding."""
import ctypes
import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "tools"))
import resblock_t_model as model   # noqa: E402


def lib_plan(k, dil, T, post, up=False):
    from infernos_b200 import _lib
    lib = _lib.load()
    out = (ctypes.c_int * 10)()
    rc = lib.b2_resblock_t_plan(k, dil[0], dil[1], dil[2], T, int(post) | (2 if up else 0), out)
    if rc:
        return None
    v = list(out)
    return dict(S=v[0], H=v[1], V=v[2], tiles=v[3], off=v[4:7], lim=v[7:10])


CASES = [
    # (k, dilations, T, conv_post fused): the vocoder's stage-3 shapes, window edges inside a slab, a single-row window, other tap counts
    (3, (1, 3, 5), 3072, False), (7, (1, 3, 5), 1100, False), (11, (1, 3, 5), 1300, True), (11, (1, 3, 5), 393, False),
    (3, (1, 3, 5), 1, False), (7, (1, 3, 5), 100, False), (5, (2, 1, 4), 900, False), (11, (1, 3, 5), 501, False), (9, (1, 2, 3), 777, True),
]


@pytest.mark.parametrize("case", CASES)
def test_library_plan_equals_the_model_plan_and_the_model_equals_conv1d(case):
    k, dil, T, post = case
    pl = lib_plan(k, dil, T, post)
    assert pl is not None
    assert pl == model.plan(k, dil, T, post)
    assert pl["S"] <= 512 and pl["H"] + pl["V"] <= pl["S"] - pl["H"] and pl["tiles"] * pl["V"] >= T
    err = model.check(k, dil, T, post, seed=sum(dil) + k + T, pl=pl)
    assert err < 1e-12, err


@pytest.mark.parametrize("case", [(3, (1, 3, 5), 3072, False), (7, (1, 3, 5), 1100, False), (11, (1, 3, 5), 1300, True), (11, (1, 3, 5), 392, True),
                                  (7, (1, 3, 5), 4, False)])
def test_fused_upsampler_variant_equals_conv_transpose_then_resblock(case):
    """UP: x = ConvTranspose1d(k8, s4, p2)(stage input) is computed inside the kernel (three taps x two channel halves of MMAs into X, tiles
    starting on a multiple of four rows) -- the model's restatement on the library's aligned plan against torch.conv_transpose1d + ResBlock"""
    k, dil, T, post = case
    pl = lib_plan(k, dil, T, post, up=True)
    assert pl is not None and pl == model.plan(k, dil, T, post, align4=True)
    assert pl["H"] % 4 == 0 and pl["V"] % 4 == 0 and pl["H"] + pl["V"] <= pl["S"] - pl["H"] and pl["tiles"] * pl["V"] >= T
    err = model.check_up(k, dil, T, post, seed=k + T, pl=pl)
    assert err < 1e-12, err


def test_plan_rejects_what_the_kernel_does_not_cover():
    from infernos_b200 import _lib
    lib = _lib.load()
    assert lib_plan(13, (1, 3, 5), 1000, False) is None            # more taps than weight slots
    assert b"not covered" in lib.b2_last_error(None)
    assert lib_plan(11, (1, 3, 7), 1000, False) is None            # extended taps would leave the guard rows
    assert lib_plan(4, (1, 3, 5), 1000, False) is None             # even tap count
