"""TTS worker thread with the reference's API (/root/reference/Cluster/InfernTTSWorker.py:56-105):
InfernTTSWorker(lang, output_sr, device=None), .infer(wi) / .start() / .stop() / .process_batch(wis),
.get_voice / .get_rand_voice / .get_rand_voice_id, attributes max_batch_size, tts_engine, output_sr.

The reference caps a batch at 8 because its eager tail is launch-bound; the B200 tail wants thousands of
windows in flight, so max_batch_size defaults to the engine's slot pool.

continuous=True replaces the reference's run-a-batch-to-completion loop (:83-92) by continuous batching:
between two engine calls the worker admits whatever has been queued as a new cohort, every cohort advances by one call, and
the tail runs once over all live sessions (HelloSippyRTPipe.infer_many).  A request therefore waits at most one call
(16 decoder steps, 0.512 s of audio) instead of the rest of the longest sentence in flight.
"""
from __future__ import annotations

from typing import List

import torch

from infernos_b200.Cluster.InfernBatchedWorker import InfernBatchedWorker
from infernos_b200.Core.InfernWrkThread import RTPWrkTRun
from infernos_b200.HelloSippyTTSRT.HelloSippyRTPipe import HelloSippyPipeState, HelloSippyPipeStateBatched, HelloSippyPlayRequest, HelloSippyRTPipe

# language -> HF model table of the reference (InfernTTSWorker.py:37-45); resolving these needs the hub
lang2model = {
    "en": {},
    "it": {"model": "Sandiago21/speecht5_finetuned_voxpopuli_it"},
    "es": {"model": "Sandiago21/speecht5_finetuned_facebook_voxpopuli_spanish"},
    "fr": {"model": "Sandiago21/speecht5_finetuned_facebook_voxpopuli_french"},
    "de": {"model": "JFuellem/speecht5_finetuned_voxpopuli_de"},
    "pt": {"model": "evertonaleixo/speecht5_finetuned_fleurs_ptbr"},
    "ru": {"model": "zaebee/speecht5_tts_common_ru"},
    "ja": {"model": "esnya/japanese_speecht5_tts"},
}


def get_torch_hw() -> str:
    if torch.cuda.is_available():
        return "cuda"
    raise AttributeError("Could not find CUDA devices (infernos_b200 is CUDA-only)")


class InfernTTSWorker(InfernBatchedWorker):
    max_batch_size: int = 8
    debug = False
    tts_engine: HelloSippyRTPipe
    output_sr: int

    continuous: bool = False
    async_dispatch: bool = False

    def __init__(self, lang, output_sr, device=None, continuous: bool = False, async_dispatch: bool = False, **engine_kwa):
        """async_dispatch (continuous mode): the dispatch callbacks of a call run on a second thread, in order, while this thread already
        drives the next engine call -- the reference overlaps generation and dispatch the same way in its load test
        (HelloSippyTTSRT/HelloSippyRTPipeTest.py:126-161: separate executors); the GPU no longer idles while Python slices and hands out audio."""
        super().__init__()
        self.continuous = continuous
        self.async_dispatch = async_dispatch and continuous
        if device is None:
            device = get_torch_hw()
        kwa = dict(lang2model[lang])
        kwa.update(engine_kwa)
        want_batch = kwa.pop("max_batch_size", None)          # a worker knob, not a SpeechT5Config override
        self.tts_engine = HelloSippyRTPipe(device, output_sr=output_sr, **kwa)
        self.output_sr = output_sr
        pool = self.tts_engine.tail.max_sessions
        self.max_batch_size = pool if want_batch is None else max(1, min(int(want_batch), pool))

    def _new_cohort(self, wis: List[HelloSippyPlayRequest]):
        """Builds the batch state of newly admitted requests.  Anything that goes wrong here (tokenisation, a missing script
        entry, an exhausted slot pool) ends those requests' sentences (`dispatch(None)`) and leaves the worker running: the
        sessions already in flight must not be stranded by a bad newcomer."""
        try:
            return HelloSippyPipeStateBatched([HelloSippyPipeState(self.tts_engine, r) for r in wis], self.tts_engine)
        except Exception as e:
            print(f"InfernTTSWorker: could not admit {len(wis)} request(s): {e!r}")
            for r in wis:
                try:
                    r.dispatch(None)
                except Exception:
                    pass
            return None

    def process_batch(self, wis: List[HelloSippyPlayRequest]):
        state = self._new_cohort(wis)
        if state is None:
            return
        while True:
            try:
                self.tts_engine.infer(state)
            except RuntimeError as e:
                self.handle_runtime_error(e, state, wis)
                raise
            if not self.tts_engine.unbatch_and_dispatch(state):
                break

    def run(self):
        if not self.continuous:
            return super().run()
        self.thread_started()
        cohorts: List[HelloSippyPipeStateBatched] = []
        dq, dthread = None, None
        if self.async_dispatch:
            import queue
            import threading
            dq = queue.Queue()

            def dispatcher():
                while True:
                    job = dq.get()
                    if job is None:
                        return
                    try:
                        job()
                    except Exception as e:          # a listener's callback must not take the worker down
                        print(f"InfernTTSWorker: dispatch callback failed: {e!r}")
            dthread = threading.Thread(target=dispatcher, daemon=True)
            dthread.start()
        try:
            self._run_continuous(cohorts, dq)
        finally:
            if dq is not None:
                dq.put(None)
                dthread.join()

    def _run_continuous(self, cohorts, dq):
        while self.get_state() == RTPWrkTRun:
            in_flight = sum(len(c.dispatch) for c in cohorts)
            room = self.max_batch_size - in_flight
            wis = self.next_batch(block=not cohorts, limit=room) if room > 0 else []
            if wis is None:
                break
            if wis:
                for wi in wis:
                    cb = getattr(wi, "_proc_start_cb", None)
                    if cb is not None:
                        cb()
                cohort = self._new_cohort(wis)
                if cohort is not None:
                    cohorts.append(cohort)
            if not cohorts:
                continue
            try:
                self.tts_engine.infer_many(cohorts)
            except RuntimeError as e:
                for c in cohorts:
                    self.handle_runtime_error(e, c, [])
                raise
            alive = []
            for c in cohorts:
                if dq is None:
                    more = self.tts_engine.unbatch_and_dispatch(c)
                else:
                    deliver, more = self.tts_engine.unbatch_prepare(c)
                    dq.put(deliver)
                if more:
                    alive.append(c)
                else:
                    c.release()                       # its pre_frames slots go back to the pool at once
            cohorts = alive

    def handle_runtime_error(self, e, state, wis: List[HelloSippyPlayRequest]):
        print(f"InfernTTSWorker.handle_runtime_error: {e}")
        for d in state.dispatch:                      # end every still-open sentence so callers do not hang
            if d is not None:
                try:
                    d(None)
                except Exception:
                    pass

    def get_voice(self, *args):
        return self.tts_engine.get_voice(*args)

    def get_rand_voice(self):
        return self.tts_engine.get_rand_voice()

    def get_rand_voice_id(self):
        return self.tts_engine.get_rand_voice_id()
