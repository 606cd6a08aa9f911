"""p99 chunk latency of the serving loop over (sessions, pipeline depth): python tools/latency_sweep.py [--sessions 5000 10000 ...]
Each row is bench.run_latency: staggered real-time arrivals of 8-frame chunks, nominal arrival -> bytes in pinned host memory."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import bench  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--sessions", type=int, nargs="*", default=[5000, 10000, 15000, 20000])
    ap.add_argument("--depths", type=int, nargs="*", default=[1, 2])
    ap.add_argument("--seconds", type=float, default=3.0)
    ap.add_argument("--max-windows", type=int, default=2048)
    ap.add_argument("--max-batch", type=int, nargs="*", default=[0])
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    import torch
    dev = torch.device("cuda", 0)
    rows = []
    for n in a.sessions:
        for d in a.depths:
            for mb in a.max_batch:
                r = bench.run_latency(dev, n, a.seconds, depth=d, max_windows=a.max_windows, max_batch=mb)
                r.pop("definition", None)
                rows.append(r)
                print(json.dumps(r), flush=True)
    if a.out:
        with open(a.out, "w") as f:
            json.dump(rows, f, indent=1)


if __name__ == "__main__":
    main()
