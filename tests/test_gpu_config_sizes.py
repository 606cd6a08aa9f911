"""GPU parity at BASELINE.json's own configuration sizes.

  C2  "batched vocoder+codec, 64 concurrent sessions, 8-frame mel chunks, 1xB200, fp32 bit-for-tolerance check":
      every stage boundary of the vocoder (modeling_speecht5.py:3062-3072) <= 1e-3 max-abs against oracle.tail on the same
      windows, the chunker output, the 8 kHz audio of four consecutive 8-frame calls, and the PCM behind the G.711 bytes.
  C3  "1,024 concurrent calls ... bf16 tensor-core conv path": the bench-shaped launch (4,096 windows in one pass, and again
      with a workspace that forces three sub-batches), per-SESSION SNR >= 40 dB against the fp32 oracle on 64 sampled sessions
      that include the first and last session of every sub-batch.

The oracle runs on the host cores for a bounded sample (64 sessions), the GPU over the full batch.
"""
import numpy as np
import pytest
import torch

from infernos_b200 import synth

pytestmark = pytest.mark.gpu

FP32_TOL = 1e-3        # north_star: fp32 mode max-abs 1e-3
BF16_SNR_DB = 40.0     # north_star: bf16 mode >= 40 dB


@pytest.fixture(scope="module")
def sds():
    return synth.hifigan_state_dict(), synth.chunker_state_dict()


def _snr(ref, x):
    from oracle.tail import snr_db
    return snr_db(torch.as_tensor(ref), torch.as_tensor(x))


# ------------------------------------------------------------------------------------------------ C2
@pytest.fixture(scope="module")
def c2(sds):
    from infernos_b200.engine import TTSTail
    t = TTSTail("cuda:0", sds[0], sds[1], mode="fp32", max_sessions=64, max_windows=64)
    yield t
    t.close()


def test_c2_every_stage_boundary_64_sessions_fp32(c2, sds):
    from oracle import tail as otail
    B = 64
    mel = synth.synth_mel(B, 8, seed=202)
    pre = synth.synth_mel(B, 4, seed=203)                     # non-zero carried frames: a mid-sentence step
    win, _ = otail.build_windows(pre, mel)                    # (64, 12, 80)
    taps = {}
    with torch.no_grad():
        ref_audio = otail.hifigan_forward(sds[0], win, taps=taps)
        ref_chunk = otail.chunker_forward(sds[1], win, ref_audio)
    audio, got = c2.vocoder_with_taps(win.cuda())
    worst = {}
    for name in c2.TAP_NAMES:
        r, g = taps[name], got[name].cpu()
        assert tuple(g.shape) == tuple(r.shape), name
        worst[name] = float((g - r).abs().max())
        assert worst[name] <= FP32_TOL, (name, worst[name])
    worst["audio"] = float((audio.cpu() - ref_audio).abs().max())
    assert worst["audio"] <= FP32_TOL
    chunk = c2.chunker(win.cuda(), audio)
    worst["chunker"] = float((chunk.cpu() - ref_chunk).abs().max())
    assert worst["chunker"] <= FP32_TOL
    print("C2 stage-boundary max-abs (64 sessions, fp32):", {k: f"{v:.2e}" for k, v in worst.items()})


def test_c2_four_8_frame_calls_audio_and_pcm(c2, sds):
    """64 sessions x four consecutive 8-frame calls through the fused tail: 8 kHz audio <= 1e-3, PCM within 1 LSB of the oracle's for
    (practically) every sample, bytes = the codes of the call's own PCM."""
    from oracle import codec as ocodec
    from oracle import tail as otail
    B = 64
    mel = synth.synth_mel(B, 32, seed=204)
    slots = torch.arange(B, dtype=torch.int32)
    c2.reset_sessions(slots.tolist())
    pre = torch.zeros(B, 4, 80)
    dmax, over1, n = 0.0, 0, 0
    for c in range(4):
        m = mel[:, 8 * c:8 * c + 8].contiguous()
        g, a = c2.tail(slots.cuda(), m.cuda())
        with torch.no_grad():
            ref, pre = otail.tts_tail(sds[0], sds[1], pre, m)
        a, g = a.cpu().numpy(), g.cpu().numpy()
        assert a.shape == (B, 1024) and g.shape == (B, 1024)
        assert np.abs(a - ref.numpy()).max() <= FP32_TOL
        pm, pr = ocodec.f32_to_pcm16(a).astype(np.int32), ocodec.f32_to_pcm16(ref.numpy()).astype(np.int32)
        d = np.abs(pm - pr)
        dmax, over1, n = max(dmax, int(d.max())), over1 + int((d > 1).sum()), n + d.size
        assert np.array_equal(g, ocodec.encode_f32(a, 0))          # bit-exact from the same PCM
    c2.poll_errors()
    print(f"C2 four 8-frame calls: max |dPCM| = {dmax}, samples with |dPCM| > 1: {over1} of {n}")
    assert dmax <= 2 and over1 <= n * 1e-3       # fp32 summation-order noise of ~3e-5 is one LSB at 32767 scale


# ------------------------------------------------------------------------------------------------ C3
def _sample_sessions(B, per_pass, n=64, seed=5):
    edges = set()
    for b0 in range(0, B, per_pass):
        edges.update((b0, min(B, b0 + per_pass) - 1))
    rng = np.random.default_rng(seed)
    rest = [int(i) for i in rng.permutation(B) if int(i) not in edges][: max(0, n - len(edges))]
    return sorted(edges) + sorted(rest)


@pytest.mark.parametrize("max_windows", [4096, 1536], ids=["one-pass-4096-windows", "three-sub-batches"])
def test_c3_1024_sessions_bf16_per_session_snr(sds, max_windows):
    from infernos_b200.engine import TTSTail
    from oracle import tail as otail
    B, F = 1024, 32
    mel = synth.synth_mel(B, F, seed=301)
    pre0 = synth.synth_mel(B, 4, seed=302)
    t = TTSTail("cuda:0", sds[0], sds[1], mode="bf16", max_sessions=B, max_windows=max_windows)
    try:
        slots = torch.arange(B, dtype=torch.int32).cuda()
        # mid-sentence state: the carried frames are set by a priming call whose last four frames are pre0
        prime = torch.cat([synth.synth_mel(B, F - 4, seed=303), pre0], dim=1)
        t.tail(slots, prime.cuda(), want_audio=False)
        g, a = t.tail(slots, mel.cuda())
        t.poll_errors()
        a, g = a.cpu(), g.cpu()
        # vocoder-only waveform of the same windows, in the same bench-shaped launches (sub-batched the same way)
        pick = _sample_sessions(B, max(1, max_windows // 4))
        win_all, _ = otail.build_windows(pre0, mel)          # chunk-major (4B, 12, 80)
        idx = torch.tensor([i * B + b for b in pick for i in range(4)])
        voc_all = t.vocoder(win_all.cuda()).cpu()            # 4,096 windows through the fused kernels
        with torch.no_grad():
            ref_a, _ = otail.tts_tail(sds[0], sds[1], pre0[pick], mel[pick])
            ref_v = otail.hifigan_forward(sds[0], win_all[idx])
    finally:
        t.close()
    snr_v = [_snr(ref_v[4 * j:4 * j + 4], voc_all[idx[4 * j:4 * j + 4]]) for j in range(len(pick))]
    snr_t = [_snr(ref_a[j], a[b]) for j, b in enumerate(pick)]
    print(f"C3 {max_windows}: vocoder waveform per-session SNR min/median {min(snr_v):.2f}/{np.median(snr_v):.2f} dB; "
          f"tail (after chunker + resampler) min/median {min(snr_t):.2f}/{np.median(snr_t):.2f} dB over {len(pick)} sessions")
    assert min(snr_v) >= BF16_SNR_DB
    assert min(snr_t) >= BF16_SNR_DB
    # bytes are the codes of this call's own 8 kHz floats
    from oracle import codec as ocodec
    assert np.array_equal(g[pick].numpy(), ocodec.encode_f32(a[pick].numpy(), 0))


def test_device_side_slot_validation(c2):
    """Slot ids come from the caller in device memory: out-of-range and duplicated ids must not touch the pool and must surface."""
    mel = synth.synth_mel(3, 8, seed=9).cuda()
    c2.reset_sessions([0, 1, 2])
    c2.tail(torch.tensor([0, 1, 2], dtype=torch.int32).cuda(), mel)
    c2.poll_errors()
    before = c2.get_pre_frames(1).clone()
    c2.tail(torch.tensor([0, 64, 2], dtype=torch.int32).cuda(), mel)           # pool has 64 slots: 0..63
    with pytest.raises(RuntimeError, match="outside the session pool"):
        c2.poll_errors()
    c2.tail(torch.tensor([1, 2, 1], dtype=torch.int32).cuda(), mel)
    with pytest.raises(RuntimeError, match="more than once"):
        c2.poll_errors()
    c2.poll_errors()                                                            # the flag is cleared once reported
    g = torch.empty(3, 1024, dtype=torch.uint8).pin_memory()
    with pytest.raises(RuntimeError, match="outside the pool"):
        c2.tail_host(torch.tensor([0, -1, 2], dtype=torch.int32).pin_memory(), mel.cpu().pin_memory(), g)
    with pytest.raises(RuntimeError, match="more than once"):
        c2.tail_host(torch.tensor([2, 2, 0], dtype=torch.int32).pin_memory(), mel.cpu().pin_memory(), g)
    c2.reset_sessions([0, 1, 2])
    c2.tail(torch.tensor([0, 1, 2], dtype=torch.int32).cuda(), mel)
    c2.poll_errors()
    assert torch.equal(c2.get_pre_frames(1), before)          # the pool survived the bad calls and still behaves
