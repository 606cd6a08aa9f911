"""G.711 codecs with the call signatures of /root/reference/Core/Codecs/G711.py:21-70, computed by the CUDA
library (integer closed forms, no lookup tables, no CPU path).

  encode(audio_tensor) -> bytes                         G711.py:25-32
  decode(bytes, resample=True, sample_rate=8000)        G711.py:34-47
  device() / to(device)                                 G711.py:49-59  (here per instance, not module globals)
  e2d_frames / d2e_frames / silence                     G711.py:61-70

G711Codec is PCMU (payload type 0) like the reference's; G711ACodec (PCMA, payload type 8) is the A-law
extension north_star asks for, with audioop.lin2alaw / alaw2lin semantics.
"""
from __future__ import annotations

import torch

from infernos_b200 import engine
from infernos_b200._lib import LAW_ALAW, LAW_ULAW
from infernos_b200.Core.AudioChunk import AudioChunk, G711AudioChunk
from .GenCodec import GenCodec


class G711Codec(GenCodec):
    ptype = 0        # G.711u
    ename = "PCMU"
    _law = LAW_ULAW
    _silence = b"\xff"

    def __init__(self):
        super().__init__()
        self._device = None

    def _dev(self) -> torch.device:
        if self._device is None:
            if not torch.cuda.is_available():
                raise RuntimeError("G711Codec needs a CUDA device (infernos_b200 has no CPU fallback)")
            self._device = torch.device("cuda", torch.cuda.current_device())
        return self._device

    def encode(self, audio_tensor) -> bytes:
        """Tensor -> bytes like the reference.  Also takes an AudioChunk; a G711AudioChunk whose payload was produced on the GPU with
        this codec's law is returned as is (no second encode on the RTP side)."""
        x = audio_tensor
        if isinstance(x, G711AudioChunk) and x.payload is not None and x.ename == self.ename and x.samplerate == self.srate:
            return x.payload
        if isinstance(x, AudioChunk):
            x = x.audio
        if not x.is_cuda:
            x = x.to(self._dev(), non_blocking=True)
        if x.dtype != torch.int16:
            x = x.to(torch.float32)
        return engine.g711_encode(x, self._law).cpu().numpy().tobytes()

    def decode(self, ulaw_bytes: bytes, resample: bool = True, sample_rate: int = GenCodec.srate) -> AudioChunk:
        if len(ulaw_bytes) == 0:                        # the reference returns an empty chunk (G711.py:36-47 on b'')
            rate = sample_rate if resample else self.srate
            return AudioChunk(torch.empty(0, dtype=torch.float32, device=self._dev()), rate)
        codes = torch.frombuffer(bytearray(ulaw_bytes), dtype=torch.uint8).to(self._dev())
        if resample and sample_rate != self.srate:
            if sample_rate != 16000:
                raise RuntimeError(f"G711Codec.decode: resampling to {sample_rate} Hz is not on the accelerated path")
            audio = engine.g711_decode_upsample(codes[None], self._law)[0]      # one fused kernel
            return AudioChunk(audio, sample_rate)
        return AudioChunk(engine.g711_decode(codes, self._law), self.srate)

    def decode_many(self, packets, resample: bool = True, sample_rate: int = GenCodec.srate):
        """decode() for the payloads of MANY calls at once: one staging copy, one launch, one copy back instead of
        a `torch.tensor(list(bytes))` per packet group (RTP/InfernRTPIngest.py:63-100).  -> list of AudioChunk on the CPU, in order;
        every packet is decoded (and resampled) on its own, exactly as len(packets) decode() calls would."""
        up = bool(resample and sample_rate != self.srate)
        if up and sample_rate != 16000:
            raise RuntimeError(f"G711Codec.decode_many: resampling to {sample_rate} Hz is not on the accelerated path")
        outs = engine.g711_decode_many(packets, self._law, upsample=up, device=self._dev())
        rate = sample_rate if up else self.srate
        return [AudioChunk(a, rate) for a in outs]

    def device(self):
        return self._dev()

    def to(self, device):
        d = torch.device(device)
        if d.type != "cuda":
            raise RuntimeError("G711Codec.to: only CUDA devices are supported (no CPU fallback)")
        self._device = torch.device("cuda", d.index if d.index is not None else torch.cuda.current_device())
        return self

    def e2d_frames(self, enframes: int, out_srate: int = GenCodec.srate) -> int:
        assert out_srate % self.srate == 0
        return enframes * out_srate // self.srate

    def d2e_frames(self, dnframes: int, in_srate: int = GenCodec.srate) -> int:
        assert in_srate % self.srate == 0
        return dnframes * self.srate // in_srate

    def silence(self, nframes: int) -> bytes:
        return self._silence * nframes


class G711ACodec(G711Codec):
    ptype = 8        # G.711a (RFC 3551)
    ename = "PCMA"
    _law = LAW_ALAW
    _silence = b"\xd5"
