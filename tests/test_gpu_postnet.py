"""GPU parity of the SpeechT5 decoder post-net (the call at HelloSippyRTPipe.py:230) through the C-ABI:
b2_postnet_forward and the post-net-inside-the-tail flag of b2_tts_tail2, against the golden vectors of the real transformers
module and the oracle.  fp32 mode: max-abs 1e-3 (north_star's fp32 tolerance; measured ~1e-5).  bf16 mode (tcgen05 convolutions,
bf16 operands, fp32 accumulate / tanh / residual): >= 40 dB SNR on the post-net's mel."""
import os

import numpy as np
import pytest
import torch

from infernos_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
FP32_TOL = 1e-3
BF16_SNR_DB = 40.0


@pytest.fixture(scope="module")
def sds():
    return synth.hifigan_state_dict(), synth.chunker_state_dict(), synth.postnet_state_dict()


def make(sds, mode, **kw):
    from infernos_b200.engine import TTSTail
    return TTSTail("cuda:0", sds[0], sds[1], mode=mode, max_sessions=64, max_windows=24, postnet_sd=sds[2], **kw)


@pytest.fixture(scope="module")
def tail32(sds):
    t = make(sds, "fp32")
    yield t
    t.close()


@pytest.fixture(scope="module")
def tail16(sds):
    t = make(sds, "bf16")
    yield t
    t.close()


def snr(ref, x):
    from oracle.tail import snr_db
    return snr_db(torch.as_tensor(ref), torch.as_tensor(x))


def test_postnet_fp32_matches_real_module_golden(tail32):
    d = np.load(os.path.join(G, "postnet_golden.npz"))
    for k_in, k_out in (("mel", "out"), ("mel_short", "out_short")):
        y = tail32.postnet(torch.from_numpy(d[k_in]).cuda()).cpu().numpy()
        assert y.shape == d[k_out].shape
        err = np.abs(y - d[k_out]).max()
        assert err <= FP32_TOL
        assert err < 1e-4, err


def test_postnet_bf16_snr_vs_real_module_golden(tail16):
    d = np.load(os.path.join(G, "postnet_golden.npz"))
    for k_in, k_out in (("mel", "out"), ("mel_short", "out_short")):
        y = tail16.postnet(torch.from_numpy(d[k_in]).cuda()).cpu().numpy()
        assert snr(d[k_out], y) >= BF16_SNR_DB
        # the residual path is exact, so also look at what the five layers add on their own
        assert snr(d[k_out] - d[k_in], y - d[k_in]) >= 30.0


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_postnet_matches_oracle_many_sessions_and_lengths(sds, mode):
    """More sessions than one pass of the workspace holds, call lengths 2..64 frames (zero padding is per session per call)."""
    from oracle.tail import postnet_forward
    t = make(sds, mode)
    try:
        for B, T, seed in ((70, 32, 1), (5, 2, 2), (3, 64, 3), (1, 1, 4), (33, 8, 5)):
            mel = synth.synth_mel(B, T, seed=100 + seed)
            ref = postnet_forward(sds[2], mel)
            y = t.postnet(mel.cuda()).cpu()
            if mode == "fp32":
                assert (y - ref).abs().max() < 1e-4
            else:
                assert snr(ref, y) >= BF16_SNR_DB
    finally:
        t.close()


@pytest.mark.parametrize("mode", ["fp32", "bf16"])
def test_tail_with_postnet_inside_equals_postnet_then_tail(sds, mode):
    """B2_TAIL_APPLY_POSTNET: feeding pre-post-net frames to the fused tail == running the post-net callable, then the tail
    (same kernels, same order), over two calls so the carried pre_frames are post-net frames too."""
    a, b = make(sds, mode), make(sds, mode)
    try:
        slots = torch.arange(5, dtype=torch.int32).cuda()
        for call in range(2):
            pre = synth.synth_mel(5, 32, seed=300 + call).cuda()
            g1, a1 = a.tail(slots, pre, apply_postnet=True)
            g2, a2 = b.tail(slots, b.postnet(pre))
            assert torch.equal(a1, a2) and torch.equal(g1, g2)
        assert torch.equal(a.get_pre_frames(3), b.get_pre_frames(3))
    finally:
        a.close()
        b.close()
    # and the whole thing against the oracle's post-net feeding the same tail (fresh contexts: pre_frames are zero)
    from oracle import tail as otail
    pre = synth.synth_mel(5, 32, seed=300)
    c, d = make(sds, mode), make(sds, mode)
    try:
        _, a_ref = c.tail(slots, otail.postnet_forward(sds[2], pre).cuda())
        _, a_got = d.tail(slots, pre.cuda(), apply_postnet=True)
        if mode == "fp32":
            assert (a_ref - a_got).abs().max() <= FP32_TOL
        else:
            assert snr(a_ref.cpu(), a_got.cpu()) >= 30.0     # bf16 post-net feeding a bf16 vocoder: two roundings stacked
    finally:
        c.close()
        d.close()


def test_postnet_errors(sds):
    from infernos_b200.engine import TTSTail
    t = TTSTail("cuda:0", sds[0], sds[1], mode="fp32", max_sessions=8, max_windows=8)
    try:
        with pytest.raises(RuntimeError):
            t.postnet(synth.synth_mel(1, 8).cuda())
        with pytest.raises(RuntimeError, match="post-net"):
            t.tail(torch.zeros(1, dtype=torch.int32).cuda(), synth.synth_mel(1, 8).cuda(), apply_postnet=True)
    finally:
        t.close()
    with pytest.raises(RuntimeError):
        TTSTail("cuda:0", sds[0], sds[1], mode="fp32", max_sessions=8, max_windows=8, postnet_sd={"layers.0.conv.weight": torch.zeros(256, 80, 5)})
