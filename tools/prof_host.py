import sys, cProfile, pstats, io, torch
from pathlib import Path
sys.path.insert(0, str(Path(__file__).resolve().parent.parent))
from mixgrpo_b200 import sampling_utils as su, ops, coefs
from mixgrpo_b200._cabi import SRC_NOISE
dev = torch.device("cuda:0")
sig = su.sd3_time_shift(3.0, torch.linspace(1, 0, 26)).to(dev)
B = 12
x = torch.randn(B, 4096, 64, device=dev); v = torch.randn(B, 4096, 64, device=dev).bfloat16(); e = torch.randn(B, 4096, 64, device=dev).bfloat16()
fn = lambda: su.flow_grpo_step(v, x, 0.7, sig, 9, None, noise=e)
for _ in range(50): fn()
torch.cuda.synchronize()
pr = cProfile.Profile(); pr.enable()
for _ in range(2000): fn()
pr.disable(); torch.cuda.synchronize()
s = io.StringIO(); pstats.Stats(pr, stream=s).sort_stats("tottime").print_stats(22); print(s.getvalue()[:6000])
