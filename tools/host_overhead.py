import sys, time, torch
sys.path.insert(0, "/root/repo")
from mixgrpo_b200 import sampling_utils as su, ops, coefs
from mixgrpo_b200._cabi import SRC_NOISE
dev = torch.device("cuda:0")
sig = su.sd3_time_shift(3.0, torch.linspace(1, 0, 26)).to(dev)
for B in (1, 12):
    x = torch.randn(B, 4096, 64, device=dev); v = torch.randn(B, 4096, 64, device=dev).bfloat16(); e = torch.randn(B, 4096, 64, device=dev).bfloat16()
    for name, fn in (("flow_grpo_step rollout (drop-in)", lambda: su.flow_grpo_step(v, x, 0.7, sig, 9, None, noise=e)),
                     ("flow_grpo_step rollout, own randn", lambda: su.flow_grpo_step(v, x, 0.7, sig, 9, None)),
                     ("flow_grpo_step train (no grad)", lambda: su.flow_grpo_step(v, x, 0.7, sig, 9, x)),):
        for _ in range(20): fn()
        torch.cuda.synchronize(); t0 = time.perf_counter()
        for _ in range(300): fn()
        t1 = time.perf_counter(); torch.cuda.synchronize(); t2 = time.perf_counter()
        print(f"B={B:2d} {name:36s} host {1e6*(t1-t0)/300:7.1f} us/call   total {1e6*(t2-t0)/300:7.1f} us/call")
    k, _ = coefs.flow(sig, 9, 0.7, "ref_cuda", True)
    fn = lambda: ops.fused_step(ops.FLOW, v, x, k, src=SRC_NOISE, noise=e, want_x0=True, round_like_torch=True)
    for _ in range(20): fn()
    torch.cuda.synchronize(); t0 = time.perf_counter()
    for _ in range(300): fn()
    t1 = time.perf_counter(); torch.cuda.synchronize()
    print(f"B={B:2d} {'ops.fused_step':36s} host {1e6*(t1-t0)/300:7.1f} us/call")
