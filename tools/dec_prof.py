#!/usr/bin/env python
"""Runs the GPU decoder (b2_dec_*) for ncu captures: `sessions` sentences, `warm` steps to grow the KV cache, then ONE more step launched
kernel by kernel (graphs off) — its 74 launches are the last 74 of the process.   python tools/dec_prof.py [sessions] [warm_steps]"""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infernos_b200 import synth
from infernos_b200.engine import TTSDecoder
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
warm = int(sys.argv[2]) if len(sys.argv) > 2 else 96
dec = TTSDecoder("cuda:0", synth.decoder_state_dict(), mode="bf16", max_sessions=n, max_rows=min(n, 1024), max_steps=warm + 8, max_enc_len=64)
dec.set_graphs(False)
slots = torch.arange(n, dtype=torch.int32).cuda()
enc = synth.synth_encoder_states(min(n, 64), 64, seed=5).cuda().repeat((n + 63) // 64, 1, 1)[:n].contiguous()
dec.start(slots, enc, None, synth.synth_speakers(n, seed=6).cuda())
done = 0
while done < warm:
    k = min(16, warm - done)
    dec.steps(slots, k)
    done += k
torch.cuda.synchronize()
print(f"--- profiled step at position {warm}", file=sys.stderr, flush=True)
torch.cuda.profiler.start()              # ncu --profile-from-start off: only this step is captured
dec.steps(slots, 1)
torch.cuda.synchronize()
torch.cuda.profiler.stop()
dec.poll_errors()
print("ok")
