"""The sound-dispatch end of a TTS request (mirror of TTSSndDispatch, /root/reference/Cluster/TTSSession.py:52-85): receives what
HelloSippyRTPipe.unbatch_and_dispatch hands out and forwards it to the RTP side's `soundout` (RTP/RTPOutputWorker.py:72-82).

The reference wraps every tensor into an AudioChunk (:81-82).  With the B200 tail the engine can hand over a ready G711AudioChunk
(audio + the G.711 bytes of the same samples, request attribute `dispatch_chunk`), which is forwarded as it is, so that the
payload-aware muxer (Core/OutputMuxer.py here) and G711Codec.encode can skip the CPU encode for single-track calls.
The Ray session bookkeeping around it (TTSSession, :87-140) is control plane and is not rebuilt."""
from __future__ import annotations

from time import monotonic
from typing import Callable, Optional

from infernos_b200.Core.AStreamMarkers import ASMarkerGeneric, ASMarkerNewSent, ASMarkerSentDoneCB
from infernos_b200.Core.AudioChunk import AudioChunk


class TTSSndDispatch:
    debug = False
    cancelled: bool = False

    def __init__(self, soundout: Callable, output_sr: int, done_cb: Optional[Callable] = None, cleanup_cb: Optional[Callable] = None):
        self.soundout, self.output_sr, self.done_cb, self.cleanup_cb = soundout, output_sr, done_cb, cleanup_cb

    def _end_marker(self):
        return ASMarkerNewSent() if self.done_cb is None else ASMarkerSentDoneCB(self.done_cb, sync=True)

    def cancel(self):
        self.cancelled = True
        self.soundout(chunk=self._end_marker())
        if self.cleanup_cb is not None:
            self.cleanup_cb()

    def sound_dispatch(self, chunk):
        """chunk: 1-D tensor (the reference contract), an AudioChunk / G711AudioChunk (pre-encoded dispatch), a marker, or None = end of sentence."""
        if self.cancelled:
            return
        do_cleanup = False
        if chunk is None:
            if self.debug:
                print(f"{monotonic():4.3f}: TTSSndDispatch.sound_dispatch {self.done_cb=}")
            chunk = self._end_marker()
            do_cleanup = True
        elif isinstance(chunk, AudioChunk):
            assert chunk.audio.size(0) > 0
        elif not isinstance(chunk, ASMarkerGeneric):
            assert chunk.size(0) > 0
            chunk = AudioChunk(chunk, self.output_sr)
        self.soundout(chunk=chunk)
        if do_cleanup and self.cleanup_cb is not None:
            self.cleanup_cb()
