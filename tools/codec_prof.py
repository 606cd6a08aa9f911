#!/usr/bin/env python
"""Runs the memory-bound codec kernels at config-5 scale (100k streams x 100 ms) a few times, for ncu captures."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infernos_b200 import engine, synth
streams, L = 100_000, 1600
x = (torch.rand(streams, L, device="cuda") * 2 - 1) * 0.9
codes = engine.g711_encode(x[:, ::2].contiguous())
for _ in range(3):
    engine.resample_g711_encode(x)
    engine.g711_encode(x)
    engine.g711_decode(codes)
    engine.g711_decode_upsample(codes)
torch.cuda.synchronize()
print("ok")
