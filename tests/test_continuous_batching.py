"""Continuous batching: host-side scheduler logic on the CPU with a stand-in tail, and on the GPU the
property that matters: a session's audio does not depend on which other sessions share the pass or when they were admitted.

Reference behaviour being replaced: InfernTTSWorker.process_batch runs a batch to completion before admitting new requests
(/root/reference/Cluster/InfernTTSWorker.py:83-92); HelloSippyPipeStateBatched.mergein is disabled
(/root/reference/HelloSippyTTSRT/HelloSippyRTPipeTest.py:145)."""
import threading
import time
import uuid

import pytest
import torch

from infernos_b200.Cluster.InfernBatchedWorker import InfernBatchedWorker
from infernos_b200.HelloSippyTTSRT.HelloSippyRTPipe import (HelloSippyPipeState, HelloSippyPipeStateBatched, HelloSippyPlayRequest, HelloSippyRTPipe,
                                                            ScriptedFrontend)


class FakeTail:
    """Row-independent, per-slot-stateful stand-in for TTSTail.tail on the CPU: audio depends on the session's own mel and on
    how many calls its slot has seen (as pre_frames continuity would), never on its neighbours."""

    def __init__(self, max_sessions):
        self.max_sessions = max_sessions
        self.calls = [0] * max_sessions
        self.batch_sizes = []

    def reset_sessions(self, slots):
        for s in slots:
            self.calls[s] = 0

    def tail(self, slots, mel, law=0, **kw):
        B, n, _ = mel.shape
        self.batch_sizes.append(B)
        per_frame = mel.sum(dim=2) * 0.01                                    # (B, n)
        audio = per_frame.repeat_interleave(128, dim=1) + torch.arange(n * 128).float() * 1e-5
        for b, s in enumerate(slots.tolist()):
            audio[b] += self.calls[s]
            self.calls[s] += 1
        return (audio * 7).to(torch.uint8), audio


def cpu_pipe(frontend, max_sessions=8):
    pp = HelloSippyRTPipe.__new__(HelloSippyRTPipe)          # the real constructor insists on a CUDA device
    pp.cuda_lock = threading.Lock()
    pp.cleanup_text = None
    pp.device = torch.device("cpu")
    pp.output_sr, pp.law, pp.fused = 8000, 0, True
    pp.speaker_embeddings = [torch.zeros(1, 512)]
    pp.frontend, pp.reduction_factor = frontend, frontend.reduction_factor
    pp.tail = FakeTail(max_sessions)
    pp.vocoder = pp.chunker = pp.resampler = None
    pp._free = list(range(max_sessions - 1, -1, -1))
    pp._slot_lock = threading.Lock()
    return pp


def plans():
    g = torch.Generator().manual_seed(5)
    # text -> (mel plan, decoder step at which the stop probability fires)
    return {"short": (torch.randn(200, 80, generator=g), 40), "medium": (torch.randn(200, 80, generator=g), 30),
            "long": (torch.randn(200, 80, generator=g), 55), "tiny": (torch.randn(200, 80, generator=g), 5)}


def collector():
    chunks, events = [], []

    def cb(chunk):
        events.append("end" if chunk is None else "chunk")
        if chunk is not None:
            chunks.append(chunk.clone())
    return chunks, events, cb


def solo(text, make_pipe):
    pp = make_pipe(ScriptedFrontend(plans(), maxlen=100))
    chunks, events, cb = collector()
    st = HelloSippyPipeStateBatched([HelloSippyPipeState(pp, HelloSippyPlayRequest(uuid.uuid4(), text, pp.get_voice(0), cb))], pp)
    while True:
        pp.infer(st)
        if not pp.unbatch_and_dispatch(st):
            break
    assert events[-1] == "end" and events.count("end") == 1
    return torch.cat(chunks)


def staggered(make_pipe, schedule):
    """schedule: {call number: [texts admitted before that call]} -> {text: audio}, batch sizes the tail saw."""
    pp = make_pipe(ScriptedFrontend(plans(), maxlen=100))
    out, cohorts, call = {}, [], 0
    ends = {}
    while True:
        for text in schedule.get(call, []):
            chunks, events, cb = collector()
            out[text], ends[text] = chunks, events
        new = [HelloSippyPlayRequest(uuid.uuid4(), t, pp.get_voice(0), _cb_of(out, ends, t)) for t in schedule.get(call, [])]
        if new:
            cohorts.append(HelloSippyPipeStateBatched([HelloSippyPipeState(pp, r) for r in new], pp))
        if not cohorts and call > max(schedule):
            break
        pp.infer_many(cohorts)
        alive = []
        for c in cohorts:
            if pp.unbatch_and_dispatch(c):
                alive.append(c)
            else:
                c.release()
        cohorts = alive
        call += 1
    assert all(e.count("end") == 1 and e[-1] == "end" for e in ends.values())
    assert sorted(pp._free) == list(range(len(pp._free)))          # every slot came back
    return {t: torch.cat(c) for t, c in out.items()}, pp


def _cb_of(out, ends, text):
    def cb(chunk):
        ends[text].append("end" if chunk is None else "chunk")
        if chunk is not None:
            out[text].append(chunk.clone())
    return cb


def test_infer_many_gives_every_session_its_solo_audio_cpu():
    want = {t: solo(t, cpu_pipe) for t in ("short", "medium", "long")}
    got, pp = staggered(cpu_pipe, {0: ["long", "short"], 2: ["medium"]})
    for t in want:
        assert got[t].shape == want[t].shape, t
        assert torch.equal(got[t], want[t]), t
    # the tail ran once per call over the union of live sessions: 2, 2, then 3 with the late cohort; ended sessions drop out
    assert pp.tail.batch_sizes[:3] == [2, 2, 3]
    assert pp.tail.batch_sizes[-1] == 1 and max(pp.tail.batch_sizes) == 3


def test_slot_pool_is_bounded_and_reused_cpu():
    pp = cpu_pipe(ScriptedFrontend(plans(), maxlen=100), max_sessions=2)
    mk = lambda t: HelloSippyPipeState(pp, HelloSippyPlayRequest(uuid.uuid4(), t, pp.get_voice(0), lambda c: None))
    a = HelloSippyPipeStateBatched([mk("short"), mk("long")], pp)
    with pytest.raises(RuntimeError, match="out of session slots"):
        HelloSippyPipeStateBatched([mk("medium")], pp)
    a.release()
    a.release()                                                     # idempotent
    b = HelloSippyPipeStateBatched([mk("medium")], pp)
    assert len(pp._free) == 1 and b.slots_host[0] in (0, 1)


class _Collect(InfernBatchedWorker):
    max_batch_size = 3

    def process_batch(self, wis):
        pass


def test_next_batch_polling_and_limit():
    w = _Collect()
    assert w.next_batch(block=False) == []
    for i in range(5):
        w.infer(i)
    assert w.next_batch(block=False, limit=2) == [0, 1]
    assert w.next_batch() == [2, 3, 4]
    w.infer(None)
    assert w.next_batch() is None


def test_worker_admits_requests_while_a_sentence_is_in_flight_cpu():
    from infernos_b200.Cluster.InfernTTSWorker import InfernTTSWorker
    w = InfernTTSWorker.__new__(InfernTTSWorker)
    InfernBatchedWorker.__init__(w)
    w.continuous, w.output_sr, w.max_batch_size = True, 8000, 4
    w.tts_engine = cpu_pipe(ScriptedFrontend(plans(), maxlen=100))
    log, lock = [], threading.Lock()
    first_long = threading.Event()
    done = threading.Event()

    def cb(name):
        def f(chunk):
            with lock:
                log.append((name, chunk is None))
                if name == "long" and chunk is not None:
                    first_long.set()
                if sum(1 for n, e in log if e) == 2:
                    done.set()
            time.sleep(0.002)
        return f
    w.infer(HelloSippyPlayRequest(uuid.uuid4(), "long", w.get_voice(0), cb("long")))
    w.start()
    assert first_long.wait(10)
    w.infer(HelloSippyPlayRequest(uuid.uuid4(), "tiny", w.get_voice(0), cb("short")))      # a one-call sentence, queued mid-flight
    assert done.wait(10)
    w.stop()
    names = [n for n, _ in log]
    # the late request was served before the long sentence finished, not after it
    assert names.index("short") < len(names) - 1 - names[::-1].index("long")
    assert ("short", True) in log and ("long", True) in log
    assert log.index(("short", True)) < log.index(("long", True))
    assert sorted(w.tts_engine._free) == list(range(8))


def _run_worker(async_dispatch):
    from infernos_b200.Cluster.InfernTTSWorker import InfernTTSWorker
    w = InfernTTSWorker.__new__(InfernTTSWorker)
    InfernBatchedWorker.__init__(w)
    w.continuous, w.async_dispatch, w.output_sr, w.max_batch_size = True, async_dispatch, 8000, 4
    w.tts_engine = cpu_pipe(ScriptedFrontend(plans(), maxlen=100))
    got, order, lock, done = {}, [], threading.Lock(), threading.Event()

    def cb(name):
        def f(chunk):
            with lock:
                order.append((name, threading.current_thread().name))
                got.setdefault(name, []).append(None if chunk is None else chunk.clone())
                if sum(1 for v in got.values() if v[-1] is None) == 3:
                    done.set()
        return f
    for t in ("long", "short", "medium"):
        w.infer(HelloSippyPlayRequest(uuid.uuid4(), t, w.get_voice(0), cb(t)))
    w.start()
    assert done.wait(20)
    w.stop()
    assert sorted(w.tts_engine._free) == list(range(8))
    return got, order


def test_async_dispatch_delivers_the_same_chunks_in_the_same_order_cpu():
    """InfernTTSWorker(async_dispatch=True): callbacks run on the dispatcher thread while the worker drives the next call (the reference's
    load test overlaps the two on separate executors, HelloSippyRTPipeTest.py:126-161); per session the chunks and their order must not change."""
    sync, _ = _run_worker(False)
    asyn, order = _run_worker(True)
    assert set(sync) == set(asyn)
    for t in sync:
        assert len(sync[t]) == len(asyn[t]) and asyn[t][-1] is None and all(c is not None for c in asyn[t][:-1]), t
        assert torch.equal(torch.cat(sync[t][:-1]), torch.cat(asyn[t][:-1])), t
    assert len({th for _, th in order}) == 1                              # every callback came from the one dispatcher thread, in queue order


@pytest.mark.gpu
def test_continuous_batching_is_session_independent_on_the_gpu():
    from infernos_b200 import synth

    def gpu_pipe(frontend, max_sessions=8):
        return HelloSippyRTPipe("cuda:0", output_sr=8000, frontend=frontend, vocoder_state_dict=synth.hifigan_state_dict(),
                                chunker_state_dict=synth.chunker_state_dict(), mode="fp32", max_sessions=max_sessions)
    want = {t: solo(t, gpu_pipe) for t in ("short", "long")}
    got, _ = staggered(gpu_pipe, {0: ["long"], 1: ["short"], 3: ["medium"]})
    for t in want:
        assert got[t].shape == want[t].shape
        assert (got[t] - want[t]).abs().max().item() < 1e-5, t
