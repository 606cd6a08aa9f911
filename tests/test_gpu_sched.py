"""GPU: the latency-bounded scheduler (b2_sched_*) must produce, for every session, exactly the bytes the plain fused tail produces for
the same chunk sequence — whatever the arrival pattern, the sub-batch boundaries, the graph buckets (padding sessions) and the
number of sub-batches in flight."""
import threading
import time

import numpy as np
import pytest
import torch

from infernos_b200 import synth

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def sds():
    return synth.hifigan_state_dict(), synth.chunker_state_dict()


def _reference_bytes(tail, mel, nframes):
    """Per-session bytes of consecutive chunks through the plain device entry (one call per chunk index, all sessions)."""
    S, total, _ = mel.shape
    slots = torch.arange(S, dtype=torch.int32)
    tail.reset_sessions(slots.tolist())
    out = []
    for c in range(total // nframes):
        g, _ = tail.tail(slots.cuda(), mel[:, c * nframes:(c + 1) * nframes].contiguous().cuda(), want_audio=False)
        out.append(g.cpu())
    tail.poll_errors()
    return torch.stack(out, dim=1)                       # (S, chunks, nframes*128)


class _Poller(threading.Thread):
    """Completions are collected on their own thread, as in a deployment: staging buffers only return to the pool through poll(), so a
    caller that submits and polls on one thread can block itself (submit then fails after 20 s instead of hanging)."""

    def __init__(self, sched, want):
        super().__init__(daemon=True)
        self.sched, self.want, self.got, self.n, self.err = sched, want, {}, 0, None

    def run(self):
        t0 = time.time()
        try:
            while self.n < self.want and time.time() - t0 < 120:
                recs, by = self.sched.poll(timeout_ms=50)
                for r in recs:
                    self.got.setdefault(int(r.tag), []).append((r.t_enqueue_ns, r.t_launch_ns, r.t_done_ns, int(r.slot), int(r.batch_sessions),
                                                                by[r.g711_offset:r.g711_offset + r.nbytes].clone()))
                self.n += len(recs)
        except Exception as e:          # surfaced by the test's assertions
            self.err = e


@pytest.mark.parametrize("mode,use_graphs,depth", [("fp32", True, 2), ("bf16", True, 3), ("bf16", False, 2)])
def test_scheduler_bytes_equal_the_plain_tail(sds, mode, use_graphs, depth):
    from infernos_b200.engine import TailScheduler, TTSTail
    S, nframes, chunks = 37, 8, 5
    mel = synth.synth_mel(S, nframes * chunks, seed=91)
    tail = TTSTail("cuda:0", sds[0], sds[1], mode=mode, max_sessions=S, max_windows=64)
    try:
        ref = _reference_bytes(tail, mel, nframes)
        tail.reset_sessions(list(range(S)))
        sched = TailScheduler(tail, nframes=nframes, depth=depth, use_graphs=use_graphs, max_batch=24)       # 37 sessions never fit one sub-batch
        poller = _Poller(sched, S * chunks)
        poller.start()
        rng = np.random.default_rng(3)
        sent = 0
        for c in range(chunks):
            order = rng.permutation(S)
            # ragged arrivals: groups of random size, a few of them back to back, then a pause
            i = 0
            while i < S:
                k = int(rng.integers(1, 12))
                idx = torch.from_numpy(order[i:i + k].astype(np.int64))
                tags = (idx * 1000 + c).to(torch.int64)
                sched.submit(idx.to(torch.int32), mel[idx, c * nframes:(c + 1) * nframes].contiguous(), tags=tags)
                sent += idx.numel()
                i += k
                if rng.random() < 0.3:
                    time.sleep(0.002)
        sched.flush()
        poller.join(130)
        assert poller.err is None and not poller.is_alive()
        got, n = poller.got, poller.n
        st = sched.stats()
        sched.close()
        assert n == sent == S * chunks and st["sessions"] == sent and st["max_sub_batch"] <= 24
        assert (st["graph_launches"] == st["sub_batches"]) if use_graphs else (st["graph_launches"] == 0)
        for s in range(S):
            for c in range(chunks):
                (te, tl, td, slot, bs, by), = got[s * 1000 + c]
                assert slot == s and te <= tl <= td and 1 <= bs <= 24
                assert torch.equal(by, ref[s, c]), (s, c)
    finally:
        tail.close()


def test_scheduler_same_session_twice_in_flight_keeps_order(sds):
    """Two chunks of ONE session submitted back to back must not share a sub-batch (the second needs the first's pre_frames)."""
    from infernos_b200.engine import TailScheduler, TTSTail
    S, nframes = 3, 8
    mel = synth.synth_mel(S, nframes * 4, seed=92)
    tail = TTSTail("cuda:0", sds[0], sds[1], mode="fp32", max_sessions=S, max_windows=16)
    try:
        ref = _reference_bytes(tail, mel, nframes)
        tail.reset_sessions(list(range(S)))
        sched = TailScheduler(tail, nframes=nframes, depth=2)
        slots = torch.tensor([0, 1, 2, 0, 1, 2, 0, 0, 1, 1, 2, 2], dtype=torch.int32)
        cidx = [0, 0, 0, 1, 1, 1, 2, 3, 2, 3, 2, 3]
        m = torch.stack([mel[int(s), c * nframes:(c + 1) * nframes] for s, c in zip(slots.tolist(), cidx)]).contiguous()
        tags = torch.tensor([int(s) * 1000 + c for s, c in zip(slots.tolist(), cidx)], dtype=torch.int64)
        poller = _Poller(sched, 12)
        poller.start()
        sched.submit(slots, m, tags=tags)                       # six sub-batches: more than there are staging buffers
        sched.flush()
        poller.join(130)
        assert poller.err is None and poller.n == 12
        got = poller.got
        assert sched.stats()["sub_batches"] >= 6
        sched.close()
        for s, c in zip(slots.tolist(), cidx):
            assert torch.equal(got[s * 1000 + c][0][5], ref[s, c])
    finally:
        tail.close()


def test_prebuild_leaves_no_graph_to_build_on_the_serving_path(sds):
    """b2_sched_prebuild captures every bucket up front: afterwards sub-batches of sizes never seen before launch existing graphs, and the
    bytes are the plain tail's."""
    from infernos_b200.engine import TailScheduler, TTSTail
    S, nframes = 40, 8
    mel = synth.synth_mel(S, nframes * 2, seed=93)
    tail = TTSTail("cuda:0", sds[0], sds[1], mode="bf16", max_sessions=S, max_windows=64)
    try:
        ref = _reference_bytes(tail, mel, nframes)
        tail.reset_sessions(list(range(S)))
        sched = TailScheduler(tail, nframes=nframes, depth=2, max_batch=48)
        sched.prebuild(0)
        built = sched.stats()["graphs_built"]
        assert built >= 6 * 2                                    # buckets 8, 16, .., 48 for at least two staging buffers
        tail.reset_sessions(list(range(S)))
        for c, sizes in ((0, (5, 17, 18)), (1, (33, 7))):        # sub-batch sizes in different buckets
            i = 0
            for n in sizes:
                poller = _Poller(sched, n)
                poller.start()
                sl = torch.arange(i, i + n, dtype=torch.int32)
                sched.submit(sl, mel[i:i + n, c * nframes:(c + 1) * nframes].contiguous(), tags=sl.to(torch.int64) * 1000 + c)
                sched.flush()
                poller.join(60)
                assert poller.err is None and poller.n == n
                for s_ in range(i, i + n):
                    assert torch.equal(poller.got[s_ * 1000 + c][0][5], ref[s_, c])
                i += n
        assert sched.stats()["graphs_built"] == built            # nothing was captured while serving
        sched.close()
    finally:
        tail.close()


def test_scheduler_rejects_bad_slots_and_survives_threads(sds):
    from infernos_b200.engine import TailScheduler, TTSTail
    S, nframes = 16, 8
    tail = TTSTail("cuda:0", sds[0], sds[1], mode="bf16", max_sessions=S, max_windows=32)
    try:
        sched = TailScheduler(tail, nframes=nframes)
        with pytest.raises(RuntimeError, match="outside the pool"):
            sched.submit(torch.tensor([S], dtype=torch.int32), synth.synth_mel(1, nframes))
        mel = synth.synth_mel(S, nframes, seed=93)

        def worker(lo, hi):
            for s in range(lo, hi):
                sched.submit(torch.tensor([s], dtype=torch.int32), mel[s:s + 1].contiguous())
        poller = _Poller(sched, S)
        poller.start()
        ths = [threading.Thread(target=worker, args=(i * 4, i * 4 + 4)) for i in range(4)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        sched.flush()
        poller.join(130)
        assert poller.err is None and poller.n == S and sorted(poller.got) == list(range(S))
        sched.close()
    finally:
        tail.close()


def test_submit_without_a_poller_fails_instead_of_hanging(sds, monkeypatch):
    """Back-pressure with nobody polling: submit gives up with an error (the wait is 20 s in the library; the test only checks the path by
    filling every staging buffer and then polling, which must let a blocked submit through)."""
    from infernos_b200.engine import TailScheduler, TTSTail
    tail = TTSTail("cuda:0", sds[0], sds[1], mode="bf16", max_sessions=8, max_windows=8)
    try:
        sched = TailScheduler(tail, nframes=8, depth=1, max_batch=1)
        mel = synth.synth_mel(8, 8, seed=94)
        done = []

        def feeder():
            for s in range(8):                                    # eight one-session sub-batches, three staging buffers
                sched.submit(torch.tensor([s], dtype=torch.int32), mel[s:s + 1].contiguous())
            done.append(1)
        th = threading.Thread(target=feeder, daemon=True)
        th.start()
        time.sleep(1.0)
        assert not done                                           # blocked on back-pressure: nothing has been polled yet
        poller = _Poller(sched, 8)
        poller.start()
        th.join(30)
        poller.join(30)
        assert done == [1] and poller.n == 8
        sched.close()
    finally:
        tail.close()
