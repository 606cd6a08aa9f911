"""Payload-aware output muxer with the API of /root/reference/Core/OutputMuxer.py:10-85:

    OutputMuxer(output_sr, qsize, device)   .chunk_in(chunk)   .idle(rtp_worker) -> audio | None
    OutputMTMuxer(output_sr, qsize, device) .chunk_in(chunk)   .idle(rtp_worker) -> audio | None

The reference muxer re-slices every track into `qsize`-sample quanta (0.1 s) and hands a bare tensor to
`self.codec.encode()` (RTP/RTPOutputWorker.py:118), which is where the G.711 encode of every call runs today, on the CPU, in a
per-call thread.  The B200 tail already produced those bytes in the same pass as the audio (G711AudioChunk.payload), but the stock
muxer loses them: chunk_in() overwrites `chunk.audio` with a concatenation (:26-27) and idle() returns a new tensor (:31,46,56).

This muxer carries the payload in lock-step with the samples through exactly the same concatenation / slicing / re-insertion
steps, and
  * when exactly ONE track has audio in a quantum, idle() returns a G711AudioChunk (audio + the matching payload bytes), which
    G711Codec.encode() passes through un-encoded -> zero encode work on the RTP side;
  * when several tracks are active the quantum is mixed like the reference does (`sum / len(tracks)`, :74-81: linear PCM is
    needed) and a plain tensor is returned, which the codec encodes as before.
Chunks without a payload (plain AudioChunk, or anything that was resampled) behave exactly as in the reference.
"""
from __future__ import annotations

from time import monotonic
from typing import Dict, List, Optional, Union

import torch
import torch.nn.functional as F

from .AStreamMarkers import ASMarkerGeneric, ASMarkerNewSent
from .AudioChunk import AudioChunk, G711AudioChunk


class _Piece:
    """Audio of one track waiting in the muxer, with the G.711 bytes of the same samples when they are known."""
    __slots__ = ("audio", "payload", "ename")

    def __init__(self, audio: torch.Tensor, payload: Optional[bytes], ename: Optional[str]):
        self.audio, self.payload, self.ename = audio, payload, ename

    @classmethod
    def of(cls, chunk: AudioChunk) -> "_Piece":
        pl = getattr(chunk, "payload", None)
        return cls(chunk.audio, pl, getattr(chunk, "ename", None) if pl is not None else None)

    def append(self, other: "_Piece") -> None:
        keep = self.payload is not None and other.payload is not None and self.ename == other.ename
        self.audio = torch.cat((self.audio, other.audio.to(self.audio.device)), dim=0)
        self.payload = (self.payload + other.payload) if keep else None
        if not keep:
            self.ename = None

    def take(self, n: int) -> "_Piece":
        """Splits the first n samples off (n may exceed what is there)."""
        head = _Piece(self.audio[:n], None if self.payload is None else self.payload[:n], self.ename)
        self.audio = self.audio[n:]
        if self.payload is not None:
            self.payload = self.payload[n:]
        return head

    def size(self) -> int:
        return self.audio.size(0)


class OutputMuxer:
    debug = False
    output_sr: int
    qsize: int
    device: str

    def __init__(self, output_sr: int, qsize: int, device: str):
        self.output_sr = output_sr
        self.qsize = qsize
        self.device = device
        self.chunks_in: List[Union[_Piece, ASMarkerGeneric]] = []

    def chunk_in(self, chunk: Union[AudioChunk, ASMarkerGeneric]):
        if isinstance(chunk, AudioChunk):
            if chunk.samplerate != self.output_sr:
                chunk = chunk.resample(self.output_sr)            # drops a payload: the samples change
            piece = _Piece.of(chunk)
            if len(self.chunks_in) > 0 and isinstance(self.chunks_in[-1], _Piece):
                self.chunks_in[-1].append(piece)
                return
            self.chunks_in.append(piece)
            return
        self.chunks_in.append(chunk)

    def idle(self, rtp_worker) -> Optional[Union[torch.Tensor, G711AudioChunk]]:
        """Same control flow as the reference (:30-56); the quantum is assembled from pieces so that the bytes follow the samples."""
        if len(self.chunks_in) == 1 and isinstance(self.chunks_in[0], _Piece) and self.chunks_in[0].size() < self.qsize:
            return None
        out: Optional[_Piece] = None
        while len(self.chunks_in) > 0 and (rsize := self.qsize - (out.size() if out else 0)) > 0:
            chunk = self.chunks_in[0]
            if isinstance(chunk, ASMarkerNewSent):
                if out is not None and out.size() > 0:
                    return self._emit(out)
                if self.debug:
                    print(f"{monotonic():4.3f}: ASMarkerNewSent {chunk.on_proc=}")
                self.chunks_in.pop(0)
                chunk.on_proc(rtp_worker)
                continue
            if isinstance(chunk, ASMarkerGeneric):                 # other markers carry no audio: drop them in order
                self.chunks_in.pop(0)
                continue
            head = chunk.take(rsize)
            if out is None:
                out = head
            else:
                out.append(head)
            if chunk.size() == 0:
                self.chunks_in.pop(0)
        if out is not None and 0 < out.size() < self.qsize:
            if self.debug:
                print(f"{monotonic():4.3f}: Reinserting {out.size()=}")
            self.chunks_in.insert(0, out)
            return None
        return self._emit(out) if out is not None and out.size() > 0 else None

    def _emit(self, piece: _Piece):
        audio = piece.audio.to(self.device) if self.device is not None and str(piece.audio.device) != str(self.device) else piece.audio
        if piece.payload is not None and self.output_sr == 8000:
            return G711AudioChunk(audio, self.output_sr, piece.payload, piece.ename)
        return audio


def _samples(x) -> torch.Tensor:
    return x.audio if isinstance(x, AudioChunk) else x


class OutputMTMuxer:
    tracks: Dict[int, OutputMuxer]

    def __init__(self, output_sr: int, qsize: int, device: str):
        self.tracks = {}
        self.output_sr = output_sr
        self.qsize = qsize
        self.device = device

    def chunk_in(self, chunk: Union[AudioChunk, ASMarkerGeneric]):
        if chunk.track_id not in self.tracks:
            self.tracks[chunk.track_id] = OutputMuxer(self.output_sr, self.qsize, self.device)
        self.tracks[chunk.track_id].chunk_in(chunk)

    def idle(self, rtp_worker):
        chunks = [c for c in [track.idle(rtp_worker) for track in self.tracks.values()] if c is not None]
        if len(chunks) == 0:
            return None
        if len(chunks) == 1:
            return chunks[0]                                        # single active track: the pre-encoded payload survives
        # mixing needs linear PCM (:74-81): the payloads are dropped here and the codec encodes the mix
        audio = [_samples(c) for c in chunks]
        max_len = max(a.size(0) for a in audio)
        audio = [F.pad(a, (0, max_len - a.size(0)), "constant", 0) if a.size(0) < max_len else a for a in audio]
        return torch.sum(torch.stack(audio), dim=0) / len(self.tracks)
