"""Audio-stream markers (mirror of /root/reference/Core/AStreamMarkers.py:7-30, without the Ray dependency): objects that travel
through the same queue as the audio chunks and are acted on by the output thread when the audio before them has been sent."""
from __future__ import annotations

from time import monotonic


class ASMarkerGeneric:
    track_id: int
    debug: bool = False

    def __init__(self, track_id: int = 0):
        self.track_id = track_id


class ASMarkerNewSent(ASMarkerGeneric):
    # runs in the context of the output worker thread (RTP/RTPOutputWorker.py consume_audio -> OutputMuxer.idle)
    def on_proc(self, tro_self, *args):
        pass


class ASMarkerSentDoneCB(ASMarkerNewSent):
    debug = False

    def __init__(self, done_cb, sync: bool = False, **kwargs):
        super().__init__(**kwargs)
        self.done_cb = done_cb
        self.sync = sync

    def on_proc(self, tro_self, *args):
        if self.debug:
            print(f"{monotonic():4.3f}: ASMarkerSentDoneCB.on_proc")
        x = self.done_cb()
        if self.sync and hasattr(x, "result"):       # the reference waits on a Ray future here (ray.get); any future-like works
            x.result()
