#!/usr/bin/env python
"""Post-net device time at the bench's shape (1,024 sessions x 32 frames), alone and inside the fused tail call, and the distance
between the two precision modes.  (Parity against the oracle and the golden vectors is tests/test_gpu_postnet.py's job: tools stay clear of the oracle directory.)"""
import os, sys, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infernos_b200 import synth
from infernos_b200.engine import TTSTail


def snr_db(ref, x):
    ref, x = ref.double(), x.double()
    return float(10 * torch.log10((ref ** 2).sum() / ((ref - x) ** 2).sum()))



def timed(fn, iters=10, warm=3):
    for _ in range(warm): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


out = {}
y_by_mode = {}
sds = synth.hifigan_state_dict(), synth.chunker_state_dict(), synth.postnet_state_dict()
S = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
for mode in ("fp32", "bf16"):
    t = TTSTail("cuda:0", sds[0], sds[1], mode=mode, max_sessions=S, max_windows=(4 * S if mode == "bf16" else 256), postnet_sd=sds[2])
    mel = synth.synth_mel(64, 32, seed=5)
    y = t.postnet(mel.cuda()).cpu()
    y_by_mode[mode] = y
    r = {}
    if mode == "bf16":      # the fp32 mode (1e-5 from the reference module, tests) as the yardstick for the tensor-core mode
        ref = y_by_mode["fp32"]
        r = {"max_abs_vs_fp32_mode": float((y - ref).abs().max()), "snr_db_vs_fp32_mode": snr_db(ref, y), "snr_db_layers_only": snr_db(ref - mel, y - mel)}
    big = synth.synth_mel(S, 32, seed=6).cuda()
    r["postnet_ms_%d_sessions" % S] = round(timed(lambda: t.postnet(big)), 4)
    if mode == "bf16":
        slots = torch.arange(S, dtype=torch.int32).cuda()
        r["tail_ms"] = round(timed(lambda: t.tail(slots, big)), 3)
        r["tail_with_postnet_ms"] = round(timed(lambda: t.tail(slots, big, apply_postnet=True)), 3)
        r["tail_ms_again"] = round(timed(lambda: t.tail(slots, big)), 3)
        for name, ap in (("classes_tail", False), ("classes_tail_with_postnet", True)):
            t.profile_begin()
            t.tail(slots, big, apply_postnet=ap)
            ms, n = t.profile_end()
            r[name] = {k: (round(v, 3), n[k]) for k, v in ms.items()}
    r["gflop"] = round(S * 32 * (80 * 256 + 3 * 256 * 256 + 256 * 80) * 5 * 2 / 1e9, 2)
    out[mode] = r
    t.close()
print(json.dumps(out))
