#!/usr/bin/env python
"""Per-launch roofline table of one TTS step from an ncu launch list (profiles/r1d_launches_bench_1024sessions.csv):
algorithmic FLOPs and HBM bytes of every launch, the time each bound alone would take -- tensor pipe at the SHAPE's MMA rate
(cycles per tcgen05.mma M=128 K=16: 32 + N/4 for N <= 128, N/2 for N = 256; tools/mma_rate.cu) at the SM clock, HBM at the measured
copy bandwidth -- and the measured time.  CPU only; writes a markdown table."""
import csv
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SMS, GHZ = 148, 1.965
peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))) if os.path.exists(os.path.join(ROOT, "MEASURED_PEAKS.json")) else {}
HBM = peaks.get("hbm_gbs", 6446.6) * 1e9
TF_SUST = peaks.get("bf16_tflops_sustained", 1413.7) * 1e12


def mma_cycles(n):
    return 32 + n / 4 if n <= 128 else n / 2


def tensor_ms(rows, n_tile, n_tiles, k_elems):
    """rows of M (all windows, incl. padding to 128), MMAs of N = n_tile, K = k_elems per output row"""
    mmas = (rows / 128) * n_tiles * (k_elems / 16)
    return mmas * mma_cycles(n_tile) / (SMS * GHZ * 1e9) * 1e3


def main():
    src = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r1d_launches_bench_1024sessions.csv")
    rows = [r for r in csv.reader(l for l in open(src) if not l.startswith("=="))]
    hdr = rows[0]
    iN, iV = hdr.index("Kernel Name"), hdr.index("Metric Value")
    launches = [(r[iN], int(r[iV]) / 1e6) for r in rows[1:] if len(r) > iV and r[iV].isdigit()]
    W = 4096
    C = [256, 128, 64, 32]
    T = [48, 192, 768, 3072]
    KS = [3, 7, 11]
    out = []

    def add(name, ms, flop, hbm_bytes, t_ms, note=""):
        bound = max(t_ms, hbm_bytes / HBM * 1e3)
        out.append((name, ms, flop, hbm_bytes, t_ms, hbm_bytes / HBM * 1e3, bound / ms if ms else 0.0, note))

    it = iter(launches)
    nxt = lambda: next(it)
    n, ms = nxt()                                            # resample+G.711 of the step
    add("resample+G.711 (4 MB)", ms, 0, 9 * 1024 * 4096, 0, "launch-latency bound")
    n, ms = nxt()
    add("build_windows", ms, 0, W * 12 * 80 * 4 * 3, 0)
    n, ms = nxt()                                            # conv_pre 80(->128) -> 512, k7, T=12, 9 windows per tile
    rows_m = W / 9 * 128
    add("conv_pre", ms, 2 * W * 12 * 80 * 512 * 7, W * 12 * (128 * 2 + 512 * 2), tensor_ms(rows_m, 256, 2, 7 * 128))
    n, ms = nxt()                                            # upsampler 0: 512 -> 4x256
    add("upsampler 0", ms, 2 * W * 48 * 512 * 256 * 2, W * (12 * 512 * 2 + 48 * 256 * 4), tensor_ms(rows_m, 128, 8, 3 * 512), "A tile re-loaded for each of 8 N tiles")
    for k in KS:                                             # stage 0, conv by conv, two 48-row windows per 128-row tile
        for d in range(3):
            for conv in (1, 2):
                n, ms = nxt()
                el = W * 48 * 256
                by = el * (2 + 2) if conv == 1 else el * (2 + 4 + 4 + 2)
                add(f"stage 0 k={k} pair {d} conv{conv}", ms, 2 * el * 256 * k, by, tensor_ms(W / 2 * 128, 256, 1, k * 256), "weights streamed from L2 at 64 B/cycle/SM")
    for i in (1, 2, 3):
        n, ms = nxt()                                        # upsampler i
        cin, cout = C[i - 1], C[i]
        rows_in = W * T[i - 1]
        nt, ntiles = (256, 4 * cout // 256) if 4 * cout >= 256 else (4 * cout, 1)
        padrows = rows_in if T[i - 1] >= 128 else W / (129 // (T[i - 1] + 1)) * 128
        add(f"upsampler {i}", ms, 2 * W * T[i] * cin * cout * 2, rows_in * cin * 2 + W * T[i] * cout * 4, tensor_ms(padrows, nt, ntiles, 3 * cin))
        for k in KS:
            n, ms = nxt()
            el = W * T[i] * C[i]
            H = (k - 1) // 2 * 12
            slab = 256 if C[i] == 128 else 512
            if C[i] == 128:
                rows_c = W * 256                             # whole 192-row window in a 256-row slab
            else:
                V = slab - 2 * H
                rows_c = W * -(-T[i] // V) * slab
            add(f"fused ResBlock C={C[i]} k={k}", ms, 6 * 2 * el * C[i] * k, el * (4 + 4 + 4), 6 * tensor_ms(rows_c, C[i], 1, k * C[i]),
                "halo rows recomputed" if C[i] < 128 else "one CTA per SM")
    rest = list(it)
    add("chunker (7 launches)", sum(m for _, m in rest), 2 * W * 12.84e6, W * (3072 * 4 * 2 + 2048 * 4 + 12 * 192 * 6 + 48 * 128 * 4 + 192 * 64 * 16), 0, "0.8 % of the FLOPs")
    tot_ms = sum(o[1] for o in out)
    tot_fl = sum(o[2] for o in out)
    lines = ["| launch | measured ms | TFLOP | TFLOP/s | % of sustained bf16 peak | tensor pipe at the shape's MMA rate, ms | HBM GB | HBM ms | bound / measured | note |", "|---|---|---|---|---|---|---|---|---|---|"]
    groups = {}
    for o in out:
        key = o[0]
        if key.startswith("stage 0"):
            key = " ".join(o[0].split()[:3]) + " " + o[0].split()[-1] + " (x3)"
        g = groups.setdefault(key, [0, 0, 0, 0, 0, o[7]])
        for j in range(5):
            g[j] += o[1 + j]
    for key, g in groups.items():
        ms, fl, by, tms, hms = g[:5]
        bound = max(tms, hms)
        lines.append(f"| {key} | {ms:.3f} | {fl / 1e12:.3f} | {fl / ms / 1e9:.0f} | {100 * fl / (ms / 1e3) / TF_SUST:.1f} | {tms:.3f} | {by / 1e9:.2f} | {hms:.3f} | {bound / ms:.2f} | {g[5]} |")
    lines.append(f"| **step** | {tot_ms:.2f} | {tot_fl / 1e12:.2f} | {tot_fl / tot_ms / 1e9:.0f} | {100 * tot_fl / (tot_ms / 1e3) / TF_SUST:.1f} | {sum(o[4] for o in out):.2f} | {sum(o[3] for o in out) / 1e9:.1f} | {sum(o[5] for o in out):.2f} | | |")
    text = "\n".join(lines)
    print(text)
    with open(os.path.join(ROOT, "profiles", "r1d_step_model.md"), "w") as f:
        f.write("# Per-launch roofline table of one 1,024-session step (final build)\n\nGenerated by `python tools/step_model.py` from `r1d_launches_bench_1024sessions.csv` "
                f"(ncu launch list: cold-cache, serialised).  Tensor-pipe time is at the SHAPE's MMA rate (M=128, K=16: 32 + N/4 cycles for N <= 128, N/2 for N = 256) on {SMS} SMs at "
                f"{GHZ} GHz, padded rows included; HBM time at the measured {HBM / 1e9:.0f} GB/s; `bound / measured` = max of the two over the measured time.\n\n" + text + "\n")


if __name__ == "__main__":
    main()
