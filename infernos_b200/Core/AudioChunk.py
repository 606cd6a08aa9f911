"""Audio container (mirror of /root/reference/Core/AudioChunk.py:8-24).  resample() runs the library's
8 kHz <-> 16 kHz polyphase kernels; other rate pairs are not on this path and raise.

G711AudioChunk additionally carries the G.711 payload the GPU produced in the same pass as the audio, so that
the RTP side (RTP/RTPOutputWorker.py:118 `self.codec.encode(chunk)`) can skip its per-call CPU encode when the call has a single
active track; the float samples stay available for the mixing path (Core/OutputMuxer.py:81 needs linear PCM)."""
from __future__ import annotations

import torch


class AudioChunk:
    debug: bool = False
    samplerate: int
    audio: torch.Tensor
    track_id: int = 0
    active: bool = True

    def __init__(self, audio: torch.Tensor, samplerate: int):
        assert isinstance(audio, torch.Tensor)
        self.audio = audio
        self.samplerate = samplerate

    def resample(self, sample_rate: int):
        assert sample_rate != self.samplerate
        from infernos_b200 import engine
        pair = (self.samplerate, sample_rate)
        audio = self.audio.to(torch.float)
        if not audio.is_cuda:
            if not torch.cuda.is_available():
                raise RuntimeError("AudioChunk.resample needs a CUDA device (infernos_b200 has no CPU fallback)")
            audio = audio.cuda()
        if pair == (8000, 16000):
            out = engine.resample_1to2(audio)
        elif pair == (16000, 8000):
            out = engine.resample_2to1(audio)
        else:
            raise RuntimeError(f"AudioChunk.resample: {pair[0]} -> {pair[1]} Hz is not on the accelerated path")
        self.audio = out.to(device=self.audio.device, dtype=self.audio.dtype)
        self.samplerate = sample_rate
        return self

    def duration(self) -> float:
        return self.audio.size(0) / self.samplerate


class G711AudioChunk(AudioChunk):
    """8 kHz audio together with its pre-encoded G.711 payload (one byte per sample, `ename` 'PCMU' or 'PCMA').  The payload is
    dropped by anything that changes the samples — resample(), or ASSIGNING `.audio` (what Core/OutputMuxer.py:26-27 does) —
    after which the chunk behaves like a plain AudioChunk.  In-place edits of the tensor (`chunk.audio *= gain`) cannot be seen
    from here: whoever does that must call drop_payload()."""

    @property
    def audio(self) -> torch.Tensor:
        return self._audio

    @audio.setter
    def audio(self, value: torch.Tensor) -> None:
        old = getattr(self, "_audio", None)
        self._audio = value
        if old is None or getattr(self, "payload", None) is None:
            return
        # a device / dtype move of the same samples (RTP/RTPOutputWorker.py:80 `chunk.audio = chunk.audio.to(self.device)`) keeps the
        # payload; anything else (concatenation, gain, a different tensor) invalidates it
        same = value.shape == old.shape and torch.equal(value.detach().to("cpu", torch.float32), old.detach().to("cpu", torch.float32))
        if not same:
            self.payload = None

    def drop_payload(self) -> None:
        self.payload = None

    def __init__(self, audio: torch.Tensor, samplerate: int, payload: bytes, ename: str = "PCMU"):
        self.payload = None
        super().__init__(audio, samplerate)
        if samplerate != 8000:
            raise ValueError("a G.711 payload is 8 kHz audio")
        if len(payload) != audio.size(0):
            raise ValueError(f"payload has {len(payload)} bytes for {audio.size(0)} samples")
        if ename not in ("PCMU", "PCMA"):
            raise ValueError(f"unknown G.711 variant {ename!r}")
        self.payload, self.ename = bytes(payload), ename

    def resample(self, sample_rate: int):
        self.payload = None
        return super().resample(sample_rate)
