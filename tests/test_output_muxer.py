"""pre-encoded dispatch behind the muxer.  CPU-only: when the payload survives, the codec is never asked to encode,
so no GPU is needed; the checker for the bytes is the oracle's closed-form encoder.

The loop below has the shape of the reference's RTPOutputWorker.consume_audio (/root/reference/RTP/RTPOutputWorker.py:84-137):
queue -> mix.chunk_in() ... mix.idle() -> codec.encode() -> 160-byte packets.
"""
import queue

import numpy as np
import pytest
import torch

from infernos_b200.Cluster.TTSSession import TTSSndDispatch
from infernos_b200.Core.AStreamMarkers import ASMarkerNewSent, ASMarkerSentDoneCB
from infernos_b200.Core.AudioChunk import AudioChunk, G711AudioChunk
from infernos_b200.Core.Codecs.G711 import G711ACodec, G711Codec
from infernos_b200.Core.OutputMuxer import OutputMTMuxer, OutputMuxer
from oracle import codec as ocodec


def _chunks(sizes, seed=0, law=0, ename="PCMU", track=0):
    g = torch.Generator().manual_seed(seed)
    out = []
    for n in sizes:
        a = (torch.rand(n, generator=g) * 2 - 1) * 0.9
        c = G711AudioChunk(a, 8000, ocodec.encode_f32(a.numpy(), law).tobytes(), ename)
        c.track_id = track
        out.append(c)
    return out


class _Worker:
    """The consume_audio loop of RTPOutputWorker.py:84-137 without the clocking and the RTP header synthesis."""

    def __init__(self, codec, out_ft=20, samplerate_out=8000, device="cpu"):
        self.codec, self.device = codec, device
        self.data_queue = queue.Queue()
        self.out_fsize = samplerate_out * out_ft // 1000
        out_qsize = out_ft * (samplerate_out // 10 // out_ft)           # :91  (~0.1 s rounded to a frame size = 800)
        self.mix = OutputMTMuxer(samplerate_out, out_qsize, device)
        self.packets, self.encoded_quanta = [], []

    def soundout(self, chunk):                                           # :72-82
        if not isinstance(chunk, ASMarkerNewSent):
            chunk.audio = chunk.audio.to(self.device)
        self.data_queue.put(chunk)
        return (self.data_queue.qsize(), False)

    def drain(self):
        while True:
            try:
                self.mix.chunk_in(self.data_queue.get(block=False))
                continue
            except queue.Empty:
                chunk_o_n = self.mix.idle(self)
                if chunk_o_n is None:
                    return
            self.encoded_quanta.append(chunk_o_n)
            by = self.codec.encode(chunk_o_n)                            # :118
            out_psize = self.codec.d2e_frames(self.out_fsize)            # :119  (160)
            while len(by) >= out_psize:
                self.packets.append(by[:out_psize])
                by = by[out_psize:]              # a sub-packet remainder of a flushed quantum is dropped, like the reference's loop


@pytest.fixture
def no_gpu_encode(monkeypatch):
    """Any attempt to actually encode fails the test: the payload must have made it to the codec."""
    from infernos_b200 import engine
    calls = []

    def boom(*a, **k):
        calls.append(1)
        raise AssertionError("G711Codec.encode ran an encode although a pre-encoded payload was available")
    monkeypatch.setattr(engine, "g711_encode", boom)
    return calls


@pytest.mark.parametrize("codec_cls,law,ename", [(G711Codec, 0, "PCMU"), (G711ACodec, 1, "PCMA")])
def test_single_track_packets_are_the_gpu_payload_with_zero_encode_calls(no_gpu_encode, codec_cls, law, ename):
    # what unbatch_and_dispatch emits for one session: 3,840 samples first (A.5), then 4,096 per call, a short last one
    sizes = [3840, 4096, 4096, 1792]
    chunks = _chunks(sizes, seed=3, law=law, ename=ename)
    w = _Worker(codec_cls())
    done = []
    d = TTSSndDispatch(w.soundout, 8000, done_cb=lambda: done.append(1))
    for c in chunks:
        d.sound_dispatch(c)
        w.drain()
    d.sound_dispatch(None)                                              # end of sentence -> ASMarkerSentDoneCB behind the audio
    w.drain()
    audio = torch.cat([c._audio for c in chunks]).numpy()
    want = ocodec.encode_f32(audio, law).tobytes()
    got = b"".join(w.packets)
    n_full = (len(want) // 800) * 800                                    # the muxer emits 800-sample quanta; the last partial one
    assert len(got) >= n_full and got == want[:len(got)]                 # ... goes out when the end marker flushes it
    assert len(got) == len(want) - (len(want) % 160)                     # whole 160-byte packets only (:120-123)
    assert all(len(p) == 160 for p in w.packets)
    assert all(isinstance(q, G711AudioChunk) for q in w.encoded_quanta)
    assert done == [1] and not no_gpu_encode


def test_payload_follows_the_samples_through_concat_slice_and_reinsert():
    m = OutputMuxer(8000, 800, "cpu")
    a, b, c = _chunks([300, 700, 900], seed=5)
    m.chunk_in(a)
    assert m.idle(None) is None                                          # < qsize: nothing yet (:32-34)
    m.chunk_in(b)
    q = m.idle(None)
    assert isinstance(q, G711AudioChunk) and q.audio.size(0) == 800
    full = torch.cat([a._audio, b._audio, c._audio])
    assert torch.equal(q.audio, full[:800]) and q.payload == ocodec.encode_f32(full[:800].numpy(), 0).tobytes()
    m.chunk_in(ASMarkerNewSent())
    m.chunk_in(c)
    q2 = m.idle(None)                                                    # the 200 left before the marker go out on their own (:38-40)
    assert q2.audio.size(0) == 200 and q2.payload == ocodec.encode_f32(full[800:1000].numpy(), 0).tobytes()
    q3 = m.idle(None)
    assert q3.audio.size(0) == 800 and q3.payload == ocodec.encode_f32(full[1000:1800].numpy(), 0).tobytes()
    assert m.idle(None) is None                                          # 100 samples wait for more


def test_plain_chunks_and_mixed_payloads_fall_back_to_tensors():
    m = OutputMuxer(8000, 800, "cpu")
    g = _chunks([500], seed=6)[0]
    m.chunk_in(g)
    m.chunk_in(AudioChunk(torch.zeros(500), 8000))                       # no payload for these samples
    q = m.idle(None)
    assert isinstance(q, torch.Tensor) and q.size(0) == 800
    m2 = OutputMuxer(8000, 800, "cpu")
    m2.chunk_in(_chunks([500], seed=7, law=0, ename="PCMU")[0])
    m2.chunk_in(_chunks([500], seed=8, law=1, ename="PCMA")[0])          # two different laws cannot share one payload
    assert isinstance(m2.idle(None), torch.Tensor)


def test_two_tracks_are_mixed_like_the_reference_and_lose_the_payload():
    mt = OutputMTMuxer(8000, 800, "cpu")
    a = _chunks([800], seed=9, track=0)[0]
    b = _chunks([800], seed=10, track=1)[0]
    mt.chunk_in(a)
    mt.chunk_in(b)
    q = mt.idle(None)
    assert isinstance(q, torch.Tensor)
    assert torch.allclose(q, (a._audio + b._audio) / 2)                  # sum / len(tracks), Core/OutputMuxer.py:81
    # a quantum in which only one of the two tracks has audio keeps that track's payload (the reference returns chunks[0] unmixed, :76)
    c = _chunks([800], seed=11, track=1)[0]
    mt.chunk_in(c)
    q2 = mt.idle(None)
    assert isinstance(q2, G711AudioChunk) and q2.payload == c.payload


def test_assigning_audio_drops_the_payload_but_a_device_move_keeps_it():
    c = _chunks([64], seed=12)[0]
    pl = c.payload
    c.audio = c.audio.to("cpu")                                          # RTPOutputWorker.soundout (:80)
    assert c.payload == pl
    c.audio = c.audio.to(torch.float64)
    assert c.payload == pl
    c.audio = torch.cat((c.audio, c.audio))                              # what the STOCK muxer does (:27)
    assert c.payload is None
    d = _chunks([64], seed=13)[0]
    d.audio = d.audio * 0.5
    assert d.payload is None


def test_done_marker_runs_in_order():
    seen = []
    m = OutputMuxer(8000, 800, "cpu")
    m.chunk_in(_chunks([800], seed=14)[0])
    m.chunk_in(ASMarkerSentDoneCB(lambda: seen.append("done")))
    assert m.idle(None) is not None and seen == []
    assert m.idle(None) is None and seen == ["done"]
