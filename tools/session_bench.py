#!/usr/bin/env python
"""The reference's own load test, on this library: /root/reference/HelloSippyTTSRT/HelloSippyRTPipeTest.py:180-236 plays N prompts through
the engine and prints, per session, time_to_first_frame, time_to_last_frame, number_of_frames and
    rtr = (time_to_last_frame - time_to_first_frame) / (number_of_frames / 8000)        (:233-235; < 1 = faster than real time)

Here: N requests go through InfernTTSWorker (continuous batching) -> HelloSippyRTPipe with the GPU front half (B200Frontend: AR decoder in
the library) -> post-net + tail in the library -> dispatch callbacks (pre-encoded G711AudioChunk), and the same per-session figures are
collected.  Text encoding is per-sentence glue outside the path: synthetic encoder states stand in for it (there is no checkpoint
offline); sentences run to `steps` decoder steps (the random model's stop probability stays low, so maxlen ends them).

    python tools/session_bench.py [--sessions 50 1000] [--steps 96] [--mode bf16]
"""
import argparse
import json
import os
import sys
import threading
import time
import uuid

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def run(n, steps, mode="bf16", enc_len=48, device="cuda:0"):
    import numpy as np
    import torch
    from infernos_b200 import synth
    from infernos_b200.Cluster.InfernTTSWorker import InfernTTSWorker
    from infernos_b200.Core.AudioChunk import G711AudioChunk
    from infernos_b200.engine import TTSDecoder
    from infernos_b200.HelloSippyTTSRT.HelloSippyRTPipe import B200Frontend, HelloSippyPlayRequest

    enc_bank = synth.synth_encoder_states(16, enc_len, seed=41)

    def tokenizer(text):
        return torch.full((1, enc_len), int(text.split("#")[1]) % 16, dtype=torch.long)

    def encoder(ids, mask):
        return enc_bank[ids[:, 0]]

    dec = TTSDecoder(device, synth.decoder_state_dict(), mode=mode, max_sessions=n, max_rows=min(n, 1024), max_steps=steps + 40, max_enc_len=enc_len)
    fe = B200Frontend(dec, tokenizer, encoder)
    w = InfernTTSWorker("en", 8000, device=device, continuous=True, frontend=fe, vocoder_state_dict=synth.hifigan_state_dict(),
                        chunker_state_dict=synth.chunker_state_dict(), postnet_state_dict=synth.postnet_state_dict(), mode=mode,
                        max_sessions=n, max_windows=min(4 * n, 4096), speaker_embeddings=[synth.synth_speakers(1, seed=42)])
    w.tts_engine.maxlenratio = 2.0 * steps / enc_len          # maxlen = steps decoder steps (:117)
    res = [None] * n
    left = [n]
    done = threading.Event()
    t0 = [0.0]

    class Sess:
        def __init__(self, i):
            self.i, self.first, self.frames, self.bytes = i, None, 0, 0

        def __call__(self, chunk):
            now = time.monotonic()
            if chunk is None:
                res[self.i] = (self.first - t0[0], now - t0[0], self.frames, self.bytes)
                left[0] -= 1
                if left[0] == 0:
                    done.set()
                return
            assert isinstance(chunk, G711AudioChunk) and len(chunk.payload) == chunk.audio.size(0)
            if self.first is None:
                self.first = now
            self.frames += chunk.audio.size(0)
            self.bytes += len(chunk.payload)
    w.start()
    # warm-up sentence (graphs, lazily built kernels), not reported
    warm = threading.Event()
    w.infer(HelloSippyPlayRequest(uuid.uuid4(), "warm #0", w.get_voice(0), lambda c: warm.set() if c is None else None, pre_encoded=True))
    assert warm.wait(300)
    t0[0] = time.monotonic()
    for i in range(n):
        w.infer(HelloSippyPlayRequest(uuid.uuid4(), f"prompt #{i}", w.get_voice(0), Sess(i), pre_encoded=True))
    ok = done.wait(600)
    wall = time.monotonic() - t0[0]
    w.stop()
    dec.close()
    assert ok, "sessions did not finish"
    a = np.array(res, dtype=np.float64)
    ttff, ttlf, frames = a[:, 0], a[:, 1], a[:, 2]
    rtr = (ttlf - ttff) / (frames / 8000.0)
    return {"sessions": n, "decoder_steps_per_sentence": steps, "audio_s_per_session": round(float(frames.mean()) / 8000.0, 3),
            "time_to_first_frame_s": {"p50": round(float(np.percentile(ttff, 50)), 4), "p99": round(float(np.percentile(ttff, 99)), 4)},
            "time_to_last_frame_s": {"p50": round(float(np.percentile(ttlf, 50)), 4), "p99": round(float(np.percentile(ttlf, 99)), 4)},
            "rtr": {"p50": round(float(np.percentile(rtr, 50)), 4), "max": round(float(rtr.max()), 4)},
            "all_faster_than_real_time": bool(rtr.max() < 1.0), "wall_s": round(wall, 3),
            "audio_s_per_wall_s": round(float(frames.sum()) / 8000.0 / wall, 1), "payload_bytes_equal_samples": bool((a[:, 3] == a[:, 2]).all()),
            "mode": mode, "what": "reference load test shape (HelloSippyRTPipeTest.py:180-236): all requests queued at t=0, per-session TTFF / TTLF / "
                                  "rtr = (TTLF - TTFF) / audio seconds; GPU decoder + post-net + tail + G.711 behind InfernTTSWorker(continuous=True)"}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sessions", type=int, nargs="*", default=[50, 1000])
    ap.add_argument("--steps", type=int, default=96)
    ap.add_argument("--mode", default="bf16")
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    rows = []
    for n in a.sessions:
        r = run(n, a.steps, a.mode)
        rows.append(r)
        print(json.dumps(r), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
