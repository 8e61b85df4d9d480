"""A C++ program with no torch and no Python (examples/c_abi_demo.cpp) drives the C ABI directly — cudaMalloc'ed buffers, its
own stream — and must produce the same bits as the Python layer on the same synthetic inputs."""
import shutil
import subprocess
from pathlib import Path

import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu
ROOT = Path(__file__).resolve().parent.parent


def _pattern(E, salt):
    i = np.arange(E, dtype=np.uint64)
    return (((i * np.uint64(2654435761) + np.uint64(salt)) % np.uint64(2048)).astype(np.float32) / np.float32(1024.0) - np.float32(1.0))


def _fnv(t):
    h = 1469598103934665603
    for a in t.contiguous().view(torch.int32).cpu().numpy().astype(np.uint32).ravel().tolist():
        h = ((h ^ a) * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def test_cpp_client_matches_python_layer(tmp_path):
    from mixgrpo_b200 import _build, coefs, ops
    from mixgrpo_b200._cabi import SRC_NOISE
    nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(nvcc).exists():
        pytest.skip("nvcc not available on this box")
    lib = _build.build()
    exe = tmp_path / "c_abi_demo"
    r = subprocess.run([nvcc, "-std=c++17", "-Wno-deprecated-gpu-targets", "-I", str(ROOT / "include"), str(ROOT / "examples" / "c_abi_demo.cpp"), "-o", str(exe),
                        "-L", str(lib.parent), "-lmixgrpo_b200", "-Xlinker", f"-rpath={lib.parent}"], capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    B, S = 3, 40                                            # n = 2560: one full tile and a partial one per sample
    n = S * 64
    sig = torch.linspace(1, 0, 26)
    sig = (3.0 * sig) / (1 + 2.0 * sig)
    k, _ = coefs.flow(sig, 9, 0.7, "ref_cuda", True)
    floats = [k.two_var, k.log_scale, k.log_norm] + [k.c[i] for i in range(16)]
    r = subprocess.run([str(exe), str(B), str(n)] + [float(f).hex() for f in floats], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    out = dict()
    logps = []
    for line in r.stdout.splitlines():
        p = line.split()
        if p[0] == "logp":
            logps.append(float.fromhex(p[2]))
        else:
            out[p[0]] = p[1]
    dev = torch.device("cuda:0")
    E = B * n
    x = torch.from_numpy(_pattern(E, 1)).view(B, S, 64).to(dev)
    v = torch.from_numpy(_pattern(E, 7)).view(B, S, 64).to(dev).bfloat16()
    e = torch.from_numpy(_pattern(E, 13)).view(B, S, 64).to(dev).bfloat16()
    xn, x0, lp, _ = ops.fused_step(ops.FLOW, v, x, k, src=SRC_NOISE, noise=e, want_x0=True, round_like_torch=True)
    assert [float(t) for t in lp.cpu()] == logps
    assert int(out["x_next_fnv"], 16) == _fnv(xn) and int(out["x0_fnv"], 16) == _fnv(x0)
