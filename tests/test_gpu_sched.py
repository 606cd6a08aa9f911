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


def _drain(sched, want, got, deadline_s=60):
    t0 = time.time()
    n = 0
    while True:
        recs, by = sched.poll(timeout_ms=50 if want else 0)
        for r in recs:
            got.setdefault(int(r.tag), []).append((r.t_enqueue_ns, r.t_launch_ns, r.t_done_ns, int(r.slot), int(r.batch_sessions),
                                                   by[r.g711_offset:r.g711_offset + r.nbytes].clone()))
        n += len(recs)
        if n >= want or time.time() - t0 >= deadline_s:
            return n


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
        got = {}
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
            _drain(sched, 0, got, deadline_s=0)          # a non-blocking poll in between keeps buffers moving
        sched.flush()
        n = sum(len(v) for v in got.values())
        n += _drain(sched, sent - n, got)
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
        sched.submit(slots, m, tags=tags)
        sched.flush()
        got = {}
        assert _drain(sched, 12, got) == 12
        sched.close()
        for s, c in zip(slots.tolist(), cidx):
            assert torch.equal(got[s * 1000 + c][0][5], ref[s, c])
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
        ths = [threading.Thread(target=worker, args=(i * 4, i * 4 + 4)) for i in range(4)]
        [t.start() for t in ths]
        [t.join() for t in ths]
        sched.flush()
        got = {}
        assert _drain(sched, S, got) == S and sorted(got) == list(range(S))
        sched.close()
    finally:
        tail.close()
