"""CPU-only checks: the C-ABI library loads and exports every symbol the header declares; the host-side
helpers (taps, synthetic weights) are deterministic; no compute call is made (there is no GPU here)."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def lib():
    from infernos_b200 import build, _lib
    build.build()
    return _lib.load()


def test_header_symbols_all_exported(lib):
    from infernos_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "infernos_b200.h")).read()
    syms = set(re.findall(r"\b(b2_[a-z0-9_]+)\s*\(", hdr))
    assert len(syms) >= 25
    assert syms == set(_lib.SIGNATURES), (syms ^ set(_lib.SIGNATURES))
    for s in syms:
        assert getattr(lib, s) is not None
    assert lib.b2_abi_version() == 1


def test_no_cuda_device_fails_loudly(lib):
    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    ctx = lib.b2_ctx_create(0, 1, 4, 4)
    assert not ctx
    assert b"no CPU fallback" in lib.b2_last_error(None)
    from infernos_b200 import synth
    from infernos_b200.engine import TTSTail
    with pytest.raises(RuntimeError):
        TTSTail("cuda:0", synth.hifigan_state_dict())
    with pytest.raises(RuntimeError):
        TTSTail("cpu", synth.hifigan_state_dict())


def test_bad_ctx_arguments(lib):
    assert not lib.b2_ctx_create(0, 7, 4, 4)
    assert b"mode" in lib.b2_last_error(None)
    assert not lib.b2_ctx_create(0, 0, 0, 4)


def test_taps_match_torchaudio_golden():
    from infernos_b200 import resample_taps as rt
    t = np.load(os.path.join(ROOT, "tests", "golden", "resample_taps.npz"))
    assert np.array_equal(rt.down_taps().numpy(), t["down"])
    assert np.array_equal(rt.up_taps().numpy(), t["up"])
    inc = open(os.path.join(ROOT, "infernos_b200", "csrc", "resample_taps.inc")).read()
    vals = [float.fromhex(v) for v in re.findall(r"(-?0x[0-9a-f.]+p[+-]\d+)f", inc)]
    assert len(vals) == 58
    assert np.array_equal(np.array(vals[:28], dtype=np.float32), t["down"])
    assert np.array_equal(np.array(vals[28:], dtype=np.float32).reshape(2, 15), t["up"])


def test_synthetic_weights_are_deterministic_and_complete():
    from infernos_b200 import synth
    a, b = synth.hifigan_state_dict(), synth.hifigan_state_dict()
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert sum(v.numel() for k, v in a.items() if k not in ("mean", "scale")) == 12_656_257 - 0
    c = synth.chunker_state_dict()
    assert sum(v.numel() for v in c.values()) == 549_120
    try:
        from transformers import SpeechT5HifiGan, SpeechT5HifiGanConfig
    except Exception:
        return
    m = SpeechT5HifiGan(SpeechT5HifiGanConfig())
    assert set(m.state_dict().keys()) == set(a.keys())
    for k, v in m.state_dict().items():
        assert tuple(v.shape) == tuple(a[k].shape), k


def test_product_package_never_imports_the_oracle():
    """Only tests/, __graft_entry__.smoke() and bench.py's CPU legs may touch oracle/: not the package, not the tools."""
    for top in ("infernos_b200", "tools", "include"):
        for dp, _, files in os.walk(os.path.join(ROOT, top)):
            for f in files:
                if f.endswith((".py", ".cu", ".cuh", ".h")):
                    src = open(os.path.join(dp, f)).read()
                    assert "import oracle" not in src and "from oracle" not in src, f


def test_postnet_layers_picks_the_25_tensors_whatever_the_prefix():
    """engine.postnet_layers: the post-net's conv / batch-norm tensors out of a SpeechT5SpeechDecoderPostnet state_dict, out of
    a whole-model state_dict (prefix `speech_decoder_postnet.`), never the encoder/decoder `layers.*`, never feat_out / prob_out."""
    import torch
    from infernos_b200 import synth
    from infernos_b200.engine import postnet_layers
    from transformers import SpeechT5Config
    from transformers.models.speecht5.modeling_speecht5 import SpeechT5SpeechDecoderPostnet
    real = SpeechT5SpeechDecoderPostnet(SpeechT5Config()).state_dict()
    got = postnet_layers(real)
    assert len(got) == 25 and set(got) == set(synth.postnet_state_dict())
    for k, v in synth.postnet_state_dict().items():
        assert tuple(v.shape) == tuple(real[k].shape), k
    whole = {"speech_decoder_postnet." + k: v for k, v in real.items()}
    whole["speecht5.decoder.wrapped_decoder.layers.0.self_attn.k_proj.weight"] = torch.zeros(4, 4)
    whole["speecht5.encoder.wrapped_encoder.layers.0.feed_forward.intermediate_dense.weight"] = torch.zeros(4, 4)
    got2 = postnet_layers(whole)
    assert set(got2) == set(got) and all(torch.equal(got2[k], got[k]) for k in got)
    with pytest.raises(RuntimeError):
        postnet_layers({k: v for k, v in real.items() if "layers.3" not in k})


def test_g711_audio_chunk_carries_its_payload_past_the_encoder():
    """a chunk that already has its GPU-made payload is not encoded again (no CUDA needed for that path), a plain
    chunk or a payload of the other law is; resample drops the payload."""
    import torch
    from infernos_b200.Core.AudioChunk import AudioChunk, G711AudioChunk
    from infernos_b200.Core.Codecs.G711 import G711ACodec, G711Codec
    audio = torch.linspace(-0.5, 0.5, 160)
    payload = bytes(range(160))
    ch = G711AudioChunk(audio, 8000, payload, "PCMU")
    assert isinstance(ch, AudioChunk) and ch.duration() == 0.02
    assert G711Codec().encode(ch) is ch.payload == payload
    with pytest.raises(ValueError):
        G711AudioChunk(audio, 8000, payload[:-1])
    with pytest.raises(ValueError):
        G711AudioChunk(audio, 16000, payload)
    with pytest.raises(ValueError):
        G711AudioChunk(audio, 8000, payload, "G722")
    if not torch.cuda.is_available():
        # the other law / a plain chunk have to go through the CUDA encoder: without a GPU that fails loudly (no CPU fallback)
        with pytest.raises(RuntimeError):
            G711ACodec().encode(ch)
        with pytest.raises(RuntimeError):
            G711Codec().encode(AudioChunk(audio, 8000))


def test_no_global_load_is_scheduled_ahead_of_the_programmatic_launch_wait():
    """Kernels launched with programmatic dependent launch are resident while their predecessor still runs; what it writes may only be read after
    `griddepcontrol.wait` (SASS: ACQBULK).  nvcc treats loads through `const __restrict__` parameters as loads of immutable memory and once hoisted
    one above the wait (csrc/common.cuh: pdl_wait).  Scan the built library: in every kernel that waits, no LDG / LD / LDGSTS may precede the wait
    in program order; TMA loads may, in the tcgen05 kernels only (weights: written once at load time, their rings start ahead of the wait by design)."""
    import re
    import shutil
    import subprocess
    cuobjdump = shutil.which("cuobjdump") or "/usr/local/cuda/bin/cuobjdump"
    so = os.path.join(ROOT, "infernos_b200", "libinfernos_b200.so")
    if not os.path.exists(cuobjdump) or not os.path.exists(so):
        pytest.skip("cuobjdump or the built library is not here")
    sass = subprocess.run([cuobjdump, "-sass", so], capture_output=True, text=True, timeout=600).stdout
    name, waited, pre, waits = None, False, {}, set()
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name, waited = m.group(1), False
            continue
        if name is None:
            continue
        if "ACQBULK" in line:
            waited = True
            waits.add(name)
        elif not waited:
            op = re.search(r"\s(LDG|LD|LDGSTS|UTMALDG)[.\s]", line)
            if op:
                pre.setdefault(name, []).append(op.group(1))
    assert len(waits) >= 10, f"only {len(waits)} kernels contain a wait: the scan no longer sees the library's kernels"
    bad = {}
    for k in waits:
        ops = set(pre.get(k, []))
        if ops - {"UTMALDG"}:
            bad[k] = sorted(ops)
        elif "UTMALDG" in ops and not re.search(r"k_conv_umma|k_resblock|k_gemm_tc", k):
            bad[k] = sorted(ops)
    assert not bad, f"global loads ahead of griddepcontrol.wait: {bad}"
