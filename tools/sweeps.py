#!/usr/bin/env python
"""Config C4 (chunk-size sweep, p99 per-chunk latency) and C5 (stand-alone codec sweep) of BASELINE.json.

    python tools/sweeps.py [--out gpurun_out/sweeps_r1.json]

C4: 256 concurrent sessions, n in {2,4,8,16,32} mel frames per call, >= 1000 timed calls each.  Latency is wall-clock from
the host call (mel in pinned host memory) to the G.711 bytes being resident in pinned host memory (b2_tts_tail_host).
n = 8,16,32 run the full tail (1,2,4 windows of n/.. 12 frames); n = 2,4 vocode an (n+4)-frame window and trim (the chunker
only exists for 12-frame windows) through the three reference callables.
C5: fused resample+encode, encode, decode, decode+upsample for 1k..100k streams x one 20 ms packet / one 100 ms quantum,
CUDA-event timed, achieved GB/s of algorithmic bytes against the measured HBM peak.
"""
from __future__ import annotations

import argparse
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import numpy as np
import torch

from infernos_b200 import engine, synth
from infernos_b200.engine import TTSTail


def pct(xs, q):
    xs = sorted(xs)
    return xs[min(len(xs) - 1, int(q * len(xs)))]


def chunk_sweep(sessions=256, calls=1000, mode="bf16"):
    dev = torch.device("cuda:0")
    tail = TTSTail(dev, synth.hifigan_state_dict(), synth.chunker_state_dict(), mode=mode, max_sessions=sessions, max_windows=sessions * 4)
    rows = []
    for n in (2, 4, 8, 16, 32):
        mel = synth.synth_mel(sessions, n, seed=n).pin_memory()
        slots = torch.arange(sessions, dtype=torch.int32).pin_memory()
        out = torch.empty(sessions, n * 128, dtype=torch.uint8).pin_memory()
        pre = torch.zeros(sessions, 4, 80, device=dev)

        if n >= 8:
            def call():
                tail.tail_host(slots, mel, out)
        else:
            def call():
                nonlocal pre
                m = mel.to(dev, non_blocking=True)
                win = torch.cat((pre, m), dim=1)                  # (B, n+4, 80): one window, 2 frames of context each side
                pre = win[:, -4:, :]
                audio = tail.vocoder(win.contiguous())[:, 512:-512].contiguous()
                g = engine.resample_g711_encode(audio)
                out.copy_(g, non_blocking=True)
                torch.cuda.synchronize()
        for _ in range(20):
            call()
        torch.cuda.synchronize()
        lat = []
        t_all = time.perf_counter()
        for _ in range(calls):
            t0 = time.perf_counter()
            call()
            lat.append((time.perf_counter() - t0) * 1e3)
        wall = time.perf_counter() - t_all
        audio_s = sessions * n * 0.016 * calls
        rows.append({"frames_per_call": n, "windows_per_call": max(1, n // 8), "chunker": n >= 8, "sessions": sessions, "calls": calls,
                     "lat_ms_p50": round(pct(lat, 0.5), 3), "lat_ms_p99": round(pct(lat, 0.99), 3), "lat_ms_max": round(max(lat), 3),
                     "streams_rtf1": round(audio_s / wall, 1), "chunk_audio_ms": n * 16})
        print(rows[-1], flush=True)
    tail.close()
    return rows


def timed(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def codec_sweep(hbm_gbs):
    dev = torch.device("cuda:0")
    rows = []
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)      # > 126 MB L2
    for streams in (1_000, 3_000, 10_000, 30_000, 100_000):
        for name, L16 in (("20ms_packet", 320), ("100ms_quantum", 1600)):
            x16 = synth.synth_audio(streams, L16).to(dev)
            x8 = x16[:, ::2].contiguous()
            pcm = engine.f32_to_pcm16(x8)
            codes = engine.g711_encode(x8)
            nout = streams * (L16 // 2)
            cases = {
                "resample+encode (fp32 16k -> ulaw 8k)": (lambda: engine.resample_g711_encode(x16), 9.0 * nout),
                "encode fp32 -> ulaw": (lambda: engine.g711_encode(x8), 5.0 * nout),
                "encode int16 -> ulaw": (lambda: engine.g711_encode(pcm), 3.0 * nout),
                "decode ulaw -> fp32 8k": (lambda: engine.g711_decode(codes), 5.0 * nout),
                "decode+upsample ulaw -> fp32 16k": (lambda: engine.g711_decode_upsample(codes), 9.0 * nout),
            }
            for cname, (fn, nbytes) in cases.items():
                def with_flush():
                    flush.zero_()
                    fn()
                t_flush = timed(lambda: flush.zero_(), 10, 2)
                ms = timed(with_flush, 10, 2) - t_flush           # cold-L2 time of the kernel (+ its output allocation)
                ms_hot = timed(fn, 20, 3)
                rows.append({"streams": streams, "unit": name, "op": cname, "algorithmic_MB": round(nbytes / 1e6, 3),
                             "ms_cold_l2": round(ms, 4), "ms_hot_l2": round(ms_hot, 4),
                             "GBps_cold": round(nbytes / max(ms, 1e-6) / 1e6, 1), "frac_of_hbm_peak_cold": round(nbytes / max(ms, 1e-6) / 1e6 / hbm_gbs, 4)})
        print(f"codec sweep: {streams} streams done", flush=True)
    return rows


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "sweeps_r1.json"))
    ap.add_argument("--calls", type=int, default=1000)
    ap.add_argument("--skip-chunk", action="store_true")
    ap.add_argument("--skip-codec", action="store_true")
    args = ap.parse_args()
    peaks = {"hbm_gbs": 6446.6}
    pj = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(pj):
        peaks = json.load(open(pj))
    res = {"gpu": torch.cuda.get_device_name(0), "hbm_gbs_peak": peaks["hbm_gbs"]}
    if not args.skip_chunk:
        res["c4_chunk_sweep_256_sessions"] = chunk_sweep(calls=args.calls)
    if not args.skip_codec:
        res["c5_codec_sweep"] = codec_sweep(peaks["hbm_gbs"])
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(res, f, indent=1)
    print("wrote", args.out)


if __name__ == "__main__":
    main()
