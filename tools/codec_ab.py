import os, sys, json, subprocess
sys.path.insert(0, os.environ.get("GRAFT_REPO_ROOT", "/root/repo"))
code = r'''
import sys, json, torch
sys.path.insert(0, %r)
from infernos_b200 import engine
out = {}
for streams, L, name in ((100000, 1600, "100k x 100 ms"), (100000, 320, "100k x 20 ms"), (20000, 8192, "20k x 8192")):
    x = (torch.rand(streams, L, device="cuda") * 2 - 1) * 0.9
    for _ in range(3): engine.resample_g711_encode(x)
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize(); a.record()
    for _ in range(20): engine.resample_g711_encode(x)
    b.record(); torch.cuda.synchronize()
    ms = a.elapsed_time(b) / 20
    out[name] = {"ms": round(ms, 4), "GBps": round(9.0 * streams * (L // 2) / (ms / 1e3) / 1e9, 1)}
    del x
print(json.dumps(out))
''' % os.environ.get("GRAFT_REPO_ROOT", "/root/repo")
res = {}
for k in ("0", "1"):
    e = dict(os.environ); e["B2_RS_KERNEL"] = k
    r = subprocess.run([sys.executable, "-c", code], env=e, capture_output=True, text=True, timeout=300)
    res["B2_RS_KERNEL=" + k] = json.loads(r.stdout.strip().splitlines()[-1]) if r.returncode == 0 else {"error": r.stderr[-500:]}
print(json.dumps(res, indent=1))
