"""Phase timestamps of the fused ResBlock kernel (B2_RB_DBG=1): one warm-up call, one reported call."""
import os, sys
if "--nodbg" not in sys.argv:
    os.environ.setdefault("B2_RB_DBG", "1")
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from infernos_b200 import synth
from infernos_b200.engine import TTSTail
n = int(sys.argv[1]) if len(sys.argv) > 1 and sys.argv[1].isdigit() else 256
t = TTSTail("cuda:0", synth.hifigan_state_dict(), synth.chunker_state_dict(), mode="bf16", max_sessions=n, max_windows=4 * n)
mel = synth.synth_mel(n, 32, seed=1).cuda()
slots = torch.arange(n, dtype=torch.int32).cuda()
for i in range(2):
    print(f"--- call {i}", file=sys.stderr, flush=True)
    if i == 1:
        torch.cuda.profiler.start()          # ncu --profile-from-start off: only the second (warm) call is captured
    t.tail(slots, mel)
    torch.cuda.synchronize()
torch.cuda.profiler.stop()
