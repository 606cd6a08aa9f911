"""Pins the oracle (oracle/) against the golden vectors the REAL reference modules produced
(oracle/make_golden.py).  CPU only."""
import hashlib
import json
import os

import numpy as np
import pytest
import torch

from infernos_b200 import synth
from oracle import codec, tail

G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def sha(a):
    return hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()


@pytest.fixture(scope="module")
def gold():
    with open(os.path.join(G, "g711_golden.json")) as f:
        return json.load(f)


@pytest.fixture(scope="module")
def taps():
    return np.load(os.path.join(G, "resample_taps.npz"))


def test_g711_tables_exhaustive(gold):
    t = np.load(os.path.join(G, "g711_tables.npz"))
    pcm = np.arange(-32768, 32768, dtype=np.int16)
    b = np.arange(256, dtype=np.uint8)
    for law, name in ((0, "ulaw"), (1, "alaw")):
        enc = codec.encode_pcm16(pcm, law)
        dec = codec.decode_pcm16(b, law)
        assert np.array_equal(enc, t[f"{name}_enc"])
        assert np.array_equal(dec, t[f"{name}_dec"])
        assert sha(enc) == gold["sha256"][f"{name}_enc"]
        assert sha(dec) == gold["sha256"][f"{name}_dec"]
    assert np.array_equal(codec.np_ulaw_enc(pcm), t["ulaw_enc"])
    assert np.array_equal(codec.np_alaw_enc(pcm), t["alaw_enc"])
    assert 0x7F not in set(t["ulaw_enc"].tolist())          # byte 0x7F is never produced


def test_g711_reference_streams(gold):
    g1 = torch.linspace(-1.25, 1.25, 48001).numpy()
    assert sha(codec.encode_f32(g1, 0)) == gold["G1_encode_linspace"]
    g2 = synth.synth_audio(64, 8192).numpy()
    assert sha(codec.encode_f32(g2, 0)) == gold["G2_encode_rand"]
    assert sha(codec.decode_f32(np.arange(256, dtype=np.uint8), 0)) == gold["G4_decode_all"]
    edge = np.array(gold["edge_in"], dtype=np.float32)
    assert codec.encode_f32(edge, 0).tolist() == gold["edge_ulaw"]


def test_float_to_pcm_truncates_toward_zero():
    x = np.array([0.5, -0.5, 0.99999, 1.5, -1.5, 1.0, -1.0], dtype=np.float32)
    assert codec.f32_to_pcm16(x).tolist() == [16383, -16383, 32766, 32767, -32768, 32767, -32767]


def test_resample_kernel_matches_torchaudio(taps):
    k, w = tail.resample_kernel(16000, 8000)
    assert w == 13 and np.array_equal(k.reshape(28).numpy(), taps["down"])
    k, w = tail.resample_kernel(8000, 16000)
    assert w == 7 and np.array_equal(k.reshape(2, 15).numpy(), taps["up"])


def test_resample_torch_restatement(taps):
    x = torch.from_numpy(taps["x"])
    assert np.array_equal(tail.resample(x, 16000, 8000).numpy(), taps["y_down"])
    assert np.array_equal(tail.resample(torch.from_numpy(taps["x_odd"]), 16000, 8000).numpy(), taps["y_down_odd"])
    assert np.array_equal(tail.resample(x[:, :200].contiguous(), 8000, 16000).numpy(), taps["y_up"])


def test_resample_c_chain_close_to_torchaudio(taps):
    """The C oracle uses a defined fmaf chain; torch's conv1d sums in another order.  Bound the gap."""
    y = codec.resample_2to1(taps["x"], taps["down"])
    assert y.shape == taps["y_down"].shape
    assert np.abs(y - taps["y_down"]).max() < 2e-6
    y = codec.resample_2to1(taps["x_odd"], taps["down"])
    assert y.shape == taps["y_down_odd"].shape and np.abs(y - taps["y_down_odd"]).max() < 2e-6
    yu = codec.resample_1to2(taps["x"][:, :200], taps["up"])
    assert np.abs(yu - taps["y_up"]).max() < 2e-6
    pcm_c = codec.f32_to_pcm16(y)
    pcm_t = codec.f32_to_pcm16(taps["y_down_odd"])
    assert np.abs(pcm_c.astype(int) - pcm_t.astype(int)).max() <= 1


def test_decode_16k_matches_reference_codec(taps):
    d = np.load(os.path.join(G, "g711_decode16k.npz"))
    x8 = codec.decode_f32(d["inp"], 0)[None]
    y = tail.resample(torch.from_numpy(x8), 8000, 16000).numpy()[0]
    assert np.array_equal(y, d["out"])
    yc = codec.resample_1to2(x8, taps["up"])[0]
    assert np.abs(yc - d["out"]).max() < 2e-6


def test_hifigan_restatement_matches_real_module():
    d = np.load(os.path.join(G, "hifigan_golden.npz"))
    sd = synth.hifigan_state_dict()
    with torch.no_grad():
        a = tail.hifigan_forward(sd, torch.from_numpy(d["mel"]))
        al = tail.hifigan_forward(sd, torch.from_numpy(d["mel_long"]))
    # same torch ops as the real module; allow for thread-count dependent summation order
    assert np.abs(a.numpy() - d["audio"]).max() < 2e-5
    assert np.abs(al.numpy() - d["audio_long"]).max() < 2e-5
    assert np.abs(d["audio"]).max() > 0.3          # the fixture is not vacuous


def test_chunker_restatement_matches_real_module():
    d = np.load(os.path.join(G, "chunker_golden.npz"))
    sd = synth.chunker_state_dict()
    with torch.no_grad():
        c = tail.chunker_forward(sd, torch.from_numpy(d["mel"]), torch.from_numpy(d["audio"]))
    assert np.abs(c.numpy() - d["out"]).max() < 2e-5
    assert np.abs(d["out"]).max() > 0.3


def test_tail_and_unbatch_match_real_infer():
    """Replays the scripted run of the REAL infer()/unbatch_and_dispatch() through the oracle tail."""
    d = np.load(os.path.join(G, "infer_golden.npz"))
    vsd, csd = synth.hifigan_state_dict(), synth.chunker_state_dict()
    plan = torch.from_numpy(d["plan"])
    B = plan.size(0)
    pre = torch.zeros(B, 4, 80)
    starts = [1] * B
    live = [True] * B
    per_session = [[] for _ in range(B)]
    with torch.no_grad():
        for c in range(int(d["ncalls_run"])):
            mel = plan[:, 32 * c:32 * (c + 1)]
            audio, pre = tail.tts_tail(vsd, csd, pre, mel)
            assert np.abs(audio.numpy() - d["audio"][c]).max() < 2e-5
            idx = int(d["idx"][c])
            ends = d["ends_at"][c].tolist()
            sl, fin, more = tail.unbatch_slices(audio.size(1), idx, starts, ends, live)
            for i in range(B):
                n = 0
                if sl[i] is not None:
                    per_session[i].append(audio[i, sl[i][0]:sl[i][1]])
                    n = sl[i][1] - sl[i][0]
                assert n == int(d["emitted"][c][i])
                if fin[i]:
                    live[i] = False
    assert not more and not bool(d["more_last"])
    for i in range(B):
        full = torch.cat(per_session[i]).numpy()
        ref = d[f"session{i}_audio"]
        assert full.shape == ref.shape
        assert full.shape[0] == (int(d["ends_at"][-1][i]) - 1) * 256
        assert np.abs(full - ref).max() < 2e-5
        # bytes: identical wherever the float PCM is not within rounding of a truncation boundary
        mine = codec.encode_f32(ref, 0)
        assert np.array_equal(mine, d[f"session{i}_ulaw"])


def test_bf16_emulation_snr_margin():
    sd = synth.hifigan_state_dict()
    mel = synth.synth_mel(2, 12)
    with torch.no_grad():
        ref = tail.hifigan_forward(sd, mel)
        emu = tail.hifigan_forward_bf16emu(sd, mel)
    assert tail.snr_db(ref, emu) > 42.0


def test_postnet_oracle_matches_real_module_golden():
    """oracle.tail.postnet_forward against the real transformers SpeechT5SpeechDecoderPostnet.postnet (the call at
    HelloSippyRTPipe.py:230), frozen by oracle/make_golden.py: a 32-frame call of 3 sessions and a 10-frame one."""
    d = np.load(os.path.join(G, "postnet_golden.npz"))
    sd = synth.postnet_state_dict()
    wsha = hashlib.sha256(torch.cat([v.flatten() for v in sd.values()]).numpy().tobytes()).hexdigest()
    assert wsha == bytes(d["weights_sha"]).decode()
    for k_in, k_out in (("mel", "out"), ("mel_short", "out_short")):
        y = tail.postnet_forward(sd, torch.from_numpy(d[k_in])).numpy()
        assert y.shape == d[k_out].shape
        assert np.abs(y - d[k_out]).max() < 1e-5
    # the post-net changes the mel by O(1): the fixture is not vacuous
    assert np.sqrt(np.mean((d["out"] - d["mel"]) ** 2)) > 0.5
    # sessions and calls are independent: a session alone gives the same rows
    y0 = tail.postnet_forward(sd, torch.from_numpy(d["mel"][1:2])).numpy()
    assert np.abs(y0 - d["out"][1:2]).max() < 1e-5
