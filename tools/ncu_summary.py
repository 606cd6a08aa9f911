"""Selected metrics of an `ncu -i X.ncu-rep --page raw --csv` export -> the small CSV committed under profiles/, and (with --traffic) the DRAM
bytes per fused-ResBlock launch that bench.py reports as roofline.traffic.

    python tools/ncu_summary.py gpurun_out/r3_tail_full_raw.csv profiles/r3_ncu_full_tail_128sessions.csv --traffic profiles/r3_resblock_traffic.json --scale 8
"""
import argparse
import csv
import json

COLS = ["Kernel Name", "launch__grid_size", "launch__registers_per_thread", "gpu__time_duration.sum",
        "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_sector_hit_rate.pct", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "launch__cluster_size"]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("raw")
    ap.add_argument("out")
    ap.add_argument("--traffic")
    ap.add_argument("--scale", type=float, default=8.0, help="windows of the bench step / windows of the captured call")
    ap.add_argument("--note", default="")
    a = ap.parse_args()
    rows = [r for r in csv.reader(open(a.raw, newline="")) if r]
    hi = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
    head, units, body = rows[hi], rows[hi + 1], rows[hi + 2:]
    idx = [head.index(c) for c in COLS if c in head]
    with open(a.out, "w", newline="") as f:
        w = csv.writer(f)
        w.writerow([head[i] for i in idx])
        w.writerow([units[i] for i in idx])
        for r in body:
            name = r[head.index("Kernel Name")]
            name = name.split("(")[0].replace("void ", "").replace("b2::", "")
            w.writerow([name if i == head.index("Kernel Name") else r[i] for i in idx])
    if a.traffic:
        ki, ri, wi = head.index("Kernel Name"), head.index("dram__bytes_read.sum"), head.index("dram__bytes_write.sum")

        def to_bytes(v, u):
            v = float(v.replace(",", ""))
            return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
        tot, n = 0.0, 0
        for r in body:
            if "k_resblock" in r[ki]:
                tot += to_bytes(r[ri], units[ri]) + to_bytes(r[wi], units[wi])
                n += 1
        json.dump({"k_resblock_bytes_per_launch": int(tot * a.scale / max(n, 1)), "launches_captured": n, "scale": a.scale,
                   "note": a.note or f"dram__bytes_read.sum + dram__bytes_write.sum of the {n} fused-ResBlock launches of one warm call, ncu --set full --clock-control none, "
                                      f"scaled x{a.scale:g} to the bench's 4,096 windows and averaged per launch ({a.out})"}, open(a.traffic, "w"), indent=1)


if __name__ == "__main__":
    main()
