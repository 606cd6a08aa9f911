"""GPU: the reference-facing Python API (HelloSippyRTPipe / InfernTTSWorker / G711Codec) replays the scripted run of the
REAL engine (tests/golden/infer_golden.npz) and must dispatch the same chunks."""
import os
import threading
import uuid

import numpy as np
import pytest
import torch

from infernos_b200 import synth

pytestmark = pytest.mark.gpu
G = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def _engine_kwargs(d, mode="fp32", **extra):
    from infernos_b200.HelloSippyTTSRT.HelloSippyRTPipe import ScriptedFrontend
    plan = torch.from_numpy(d["plan"])
    kw = dict(frontend=ScriptedFrontend(plan, d["stop_step"].tolist(), maxlen=60), vocoder_state_dict=synth.hifigan_state_dict(),
              chunker_state_dict=synth.chunker_state_dict(), mode=mode, max_sessions=8)
    kw.update(extra)
    return kw


def _collect(B):
    got, ended = [[] for _ in range(B)], [0] * B

    def mk(i):
        def cb(chunk):
            if chunk is None:
                ended[i] += 1
            else:
                assert chunk.dim() == 1 and not chunk.is_cuda and chunk.size(0) > 0
                got[i].append(chunk.clone())
        return cb
    return got, ended, [mk(i) for i in range(B)]


@pytest.mark.parametrize("fused", [True, False])
def test_engine_replays_real_infer(fused):
    from infernos_b200.HelloSippyTTSRT.HelloSippyRTPipe import HelloSippyPipeState, HelloSippyPipeStateBatched, HelloSippyPlayRequest, HelloSippyRTPipe
    from oracle import codec as ocodec
    d = np.load(os.path.join(G, "infer_golden.npz"))
    B = d["plan"].shape[0]
    pp = HelloSippyRTPipe("cuda:0", output_sr=8000, fused=fused, **_engine_kwargs(d))
    assert (pp.chunk_size, pp.pre_nframes, pp.post_nframes, pp.model_sr) == (8, 2, 2, 16000)
    got, ended, cbs = _collect(B)
    payload = [bytearray() for _ in range(B)]
    reqs = [HelloSippyPlayRequest(uuid.uuid4(), "hello", pp.get_voice(0), cbs[i], dispatch_g711=payload[i].extend) for i in range(B)]
    state = HelloSippyPipeStateBatched([HelloSippyPipeState(pp, r) for r in reqs], pp)
    calls = 0
    while True:
        n_before = [sum(x.numel() for x in g) for g in got]
        pp.infer(state)
        more = pp.unbatch_and_dispatch(state)
        assert [sum(x.numel() for x in g) - n for g, n in zip(got, n_before)] == d["emitted"][calls].tolist()
        assert state.ends_at.tolist() == d["ends_at"][calls].tolist() and state.idx == int(d["idx"][calls])
        assert np.abs(state.audio.cpu().numpy() - d["audio"][calls]).max() < 1e-4
        if fused:
            assert np.array_equal(state.g711.cpu().numpy(), ocodec.encode_f32(state.audio.cpu().numpy(), 0))
        calls += 1
        if not more:
            break
    assert calls == int(d["ncalls_run"]) and ended == d["ended"].tolist()
    for i in range(B):
        full = torch.cat(got[i]).numpy()
        assert full.shape == d[f"session{i}_audio"].shape
        assert np.abs(full - d[f"session{i}_audio"]).max() < 1e-4
        if fused:      # the GPU-encoded payload handed to dispatch_g711 is exactly the G.711 code of the dispatched samples
            assert bytes(payload[i]) == ocodec.encode_f32(full, 0).tobytes()


def test_pre_encoded_dispatch_reaches_the_packetiser_without_an_encode(monkeypatch):
    """SURVEY 8 f1 end to end on the GPU: engine -> dispatch(G711AudioChunk) -> TTSSndDispatch -> soundout -> payload-aware OutputMTMuxer
    -> G711Codec.encode (pass-through) -> 160-byte packets == G.711 codes of the dispatched samples, with zero encode calls."""
    import queue
    from infernos_b200 import engine
    from infernos_b200.Cluster.TTSSession import TTSSndDispatch
    from infernos_b200.Core.AudioChunk import G711AudioChunk
    from infernos_b200.Core.Codecs.G711 import G711Codec
    from infernos_b200.Core.OutputMuxer import OutputMTMuxer
    from infernos_b200.HelloSippyTTSRT.HelloSippyRTPipe import HelloSippyPipeState, HelloSippyPipeStateBatched, HelloSippyPlayRequest, HelloSippyRTPipe
    from oracle import codec as ocodec
    d = np.load(os.path.join(G, "infer_golden.npz"))
    B = d["plan"].shape[0]
    pp = HelloSippyRTPipe("cuda:0", output_sr=8000, **_engine_kwargs(d))
    codec = G711Codec().to("cuda:0")
    qs = [queue.Queue() for _ in range(B)]
    samples = [[] for _ in range(B)]

    def soundout(i):
        def f(chunk):
            if isinstance(chunk, G711AudioChunk):
                samples[i].append(chunk.audio.clone())
                chunk.audio = chunk.audio.to("cpu")                     # RTPOutputWorker.soundout's device move (:80)
            qs[i].put(chunk)
        return f
    disp = [TTSSndDispatch(soundout(i), 8000) for i in range(B)]
    reqs = [HelloSippyPlayRequest(uuid.uuid4(), "hello", pp.get_voice(0), disp[i].sound_dispatch, pre_encoded=True) for i in range(B)]
    state = HelloSippyPipeStateBatched([HelloSippyPipeState(pp, r) for r in reqs], pp)
    while True:
        pp.infer(state)
        if not pp.unbatch_and_dispatch(state):
            break
    monkeypatch.setattr(engine, "g711_encode", lambda *a, **k: (_ for _ in ()).throw(AssertionError("encode ran")))
    for i in range(B):
        mix, packets = OutputMTMuxer(8000, 800, "cpu"), []
        while True:
            try:
                mix.chunk_in(qs[i].get(block=False))
                continue
            except queue.Empty:
                q = mix.idle(None)
                if q is None:
                    break
            by = codec.encode(q)
            while len(by) >= 160:
                packets.append(by[:160])
                by = by[160:]
        full = torch.cat(samples[i]).numpy()
        assert full.shape == d[f"session{i}_audio"].shape
        want = ocodec.encode_f32(full, 0).tobytes()
        got = b"".join(packets)
        assert len(got) >= (len(want) // 800) * 800 and got == want[:len(got)]


def test_worker_thread_end_to_end_and_codec():
    from infernos_b200.Cluster.InfernTTSWorker import InfernTTSWorker
    from infernos_b200.Core.Codecs.G711 import G711ACodec, G711Codec
    from infernos_b200.HelloSippyTTSRT.HelloSippyRTPipe import HelloSippyPlayRequest
    from oracle import codec as ocodec
    d = np.load(os.path.join(G, "infer_golden.npz"))
    B = d["plan"].shape[0]
    w = InfernTTSWorker("en", 8000, device="cuda:0", **_engine_kwargs(d, mode="bf16"))
    assert w.output_sr == 8000 and w.max_batch_size == 8
    got, ended, cbs = _collect(B)
    done = threading.Event()

    def wrap(i):
        def cb(chunk):
            cbs[i](chunk)
            if sum(ended) == B:
                done.set()
        return cb
    # all three requests must land in one batch: queue them before the thread starts (next_batch drains without waiting)
    for i in range(B):
        w.infer(HelloSippyPlayRequest(uuid.uuid4(), "hello", w.get_voice(0), wrap(i)))
    w.start()
    assert done.wait(60)
    w.stop()
    codec = G711Codec().to("cuda:0")
    for i in range(B):
        full = torch.cat(got[i])
        ref = d[f"session{i}_audio"]
        assert full.numel() == ref.shape[0]
        snr = 10 * np.log10((ref ** 2).sum() / ((ref - full.numpy()) ** 2).sum())
        print(f'worker end-to-end bf16 SNR session {i}: {snr:.2f} dB')
        assert snr >= 40.0        # north_star's bf16 bar, on the dispatched 8 kHz audio
        by = codec.encode(full)                           # RTPOutputWorker's call (RTP/RTPOutputWorker.py:118)
        assert isinstance(by, bytes) and len(by) == full.numel()
        assert by == ocodec.encode_f32(full.numpy(), 0).tobytes()
    assert G711Codec.rtpmap() == "rtpmap:0 PCMU/8000" and G711ACodec.rtpmap() == "rtpmap:8 PCMA/8000"
    assert codec.silence(3) == b"\xff\xff\xff" and codec.e2d_frames(160, 16000) == 320 and codec.d2e_frames(320, 16000) == 160
    ch = codec.decode(bytes(range(256)), resample=False)
    assert ch.samplerate == 8000 and np.array_equal(ch.audio.cpu().numpy(), ocodec.decode_f32(np.arange(256, dtype=np.uint8), 0))
    ch16 = codec.decode(bytes(range(256)) * 2, sample_rate=16000)
    g = np.load(os.path.join(G, "g711_decode16k.npz"))
    assert ch16.samplerate == 16000 and np.abs(ch16.audio.cpu().numpy() - g["out"]).max() < 2e-6
    a = G711ACodec().to("cuda:0")
    x = synth.synth_audio(1, 4000)[0]
    assert a.encode(x) == ocodec.encode_f32(x.numpy(), 1).tobytes()


@pytest.mark.parametrize("fused", [True, False])
def test_engine_with_gpu_postnet_equals_frontend_postnet(fused):
    """postnet_state_dict moves HelloSippyRTPipe.py:230 onto the GPU (inside the tail call when fused).  Same run twice: once with
    the post-net done by the front end on the CPU (the oracle's restatement of the transformers module), once by the engine."""
    from infernos_b200.HelloSippyTTSRT.HelloSippyRTPipe import (HelloSippyPipeState, HelloSippyPipeStateBatched, HelloSippyPlayRequest,
                                                                 HelloSippyRTPipe, ScriptedFrontend)
    from oracle import tail as otail
    d = np.load(os.path.join(G, "infer_golden.npz"))
    B = d["plan"].shape[0]
    psd = synth.postnet_state_dict()
    called = []

    class CpuPostnetFrontend(ScriptedFrontend):
        def postnet(self, spectrogram):
            called.append(1)
            return otail.postnet_forward(psd, spectrogram.float().cpu())

    def run(**kw):
        pp = HelloSippyRTPipe("cuda:0", output_sr=8000, fused=fused, **_engine_kwargs(d, **kw))
        got, ended, cbs = _collect(B)
        reqs = [HelloSippyPlayRequest(uuid.uuid4(), "hello", pp.get_voice(0), cbs[i]) for i in range(B)]
        state = HelloSippyPipeStateBatched([HelloSippyPipeState(pp, r) for r in reqs], pp)
        audios = []
        while True:
            pp.infer(state)
            audios.append(state.audio.cpu())
            if not pp.unbatch_and_dispatch(state):
                break
        return pp, audios, got, ended

    plan = torch.from_numpy(d["plan"])
    _, ref_audio, ref_got, ref_ended = run(frontend=CpuPostnetFrontend(plan, d["stop_step"].tolist(), maxlen=60))
    assert len(called) == len(ref_audio)
    n = len(called)
    pp, audio, got, ended = run(frontend=CpuPostnetFrontend(plan, d["stop_step"].tolist(), maxlen=60), postnet_state_dict=psd)
    assert pp.gpu_postnet and len(called) == n            # the front end's post-net was not called again
    assert ended == ref_ended and len(audio) == len(ref_audio)
    for a, r in zip(audio, ref_audio):
        assert (a - r).abs().max() < 1e-3
    for i in range(B):
        assert torch.cat(got[i]).shape == torch.cat(ref_got[i]).shape
