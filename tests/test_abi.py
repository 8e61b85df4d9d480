"""C-ABI library: loads on a CPU-only box, exports every symbol include/mixgrpo_b200.h declares, and
rejects bad arguments before touching CUDA (no compute calls without a GPU)."""
import ctypes
import re
from pathlib import Path

import pytest

ROOT = Path(__file__).resolve().parent.parent
HEADER = (ROOT / "include" / "mixgrpo_b200.h").read_text()


def _declared():
    body = re.sub(r"/\*.*?\*/", "", HEADER, flags=re.S)
    return sorted(set(re.findall(r"\b(mixgrpo_[a-z0-9_]+)\s*\(", body)))


def test_header_and_binding_agree():
    from mixgrpo_b200 import _cabi
    assert _declared() == sorted(_cabi.SIGNATURES), "ctypes binding and header disagree"


def test_library_builds_loads_and_exports_all_symbols():
    from mixgrpo_b200 import _build, _cabi
    lib_path = _build.build()
    h = ctypes.CDLL(str(lib_path))
    for name in _declared():
        assert hasattr(h, name), f"{name} not exported"
    lib = _cabi.lib()
    assert lib.mixgrpo_abi_version() == _cabi.ABI_VERSION
    assert b"sm_100a" in lib.mixgrpo_build_info()
    assert lib.mixgrpo_error_string(0) == b"success"
    assert b"invalid" in lib.mixgrpo_error_string(-1)


def test_sass_is_sm100a_with_256bit_accesses():
    """The shipped cubin targets sm_100a and the step kernel uses LDG.E.256 / STG.E.256 (Blackwell-only widths)."""
    import shutil
    import subprocess
    from mixgrpo_b200 import _build
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    lib = str(_build.build())
    elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
    assert "sm_100a" in elf
    usage = subprocess.run(["cuobjdump", "-res-usage", lib], capture_output=True, text=True).stdout
    names = [ln.split()[1].rstrip(":") for ln in usage.splitlines() if ln.strip().startswith("Function") and "step_kernelILi0E13__nv_bfloat16S1_Li0E" in ln]
    assert names, "no flow/bf16/SDE step_kernel instantiation in the library"
    sass = subprocess.run(["cuobjdump", "-sass", "-fun", names[0], lib], capture_output=True, text=True).stdout
    assert ".256" in sass and "LDG" in sass and "STG" in sass


def test_host_side_argument_validation_needs_no_gpu():
    from mixgrpo_b200 import _cabi
    lib = _cabi.lib()
    k = _cabi.StepCoefs()
    assert lib.mixgrpo_step_workspace_bytes(12, 4096 * 64) == 512
    assert lib.mixgrpo_step_workspace_bytes(0, 10) == 0
    assert lib.mixgrpo_step_workspace_bytes(100, 10) == 3328
    # a deferred launch spreads a sample's arrivals over 8 such records
    assert lib.mixgrpo_deferred_workspace_bytes(12, 4096 * 64) == 12 * 256
    assert lib.mixgrpo_deferred_workspace_bytes(0, 10) == 0
    assert lib.mixgrpo_deferred_workspace_bytes(3, 10) == 768
    # null pointers / bad sizes / bad enums are rejected with MIXGRPO_EINVAL before any launch
    assert lib.mixgrpo_flow_step(None, 1, None, 0, None, None, 0, None, 0, None, None, None, None, 0, 1, 8, ctypes.byref(k), 0, 0, None, None) == -1
    assert lib.mixgrpo_dpm_step(1, 1, 1, 8, None, None, None, 4, None, 8, None, None, None, None, 0, 1, 8, ctypes.byref(k), 2, 0, None, None) == -1
    assert lib.mixgrpo_logprob_bwd(7, 1, 1, 1, 8, 1, 8, 1, 1, 1, 8, ctypes.byref(k), 0, None) == -1
    assert lib.mixgrpo_group_advantages(None, None, 1, 4, 4, 0, 1, None, 0, None, None) == -1
    assert lib.mixgrpo_grpo_loss(None, None, None, 1, 1e-4, 5.0, 0.0, 12.0, None, None, None, None) == -1
    assert lib.mixgrpo_pack_latents(None, None, 0, 1, 16, 8, 8, None) == -1
    assert lib.mixgrpo_set_tuning(99, 1) == -1
    # peer exchange: region sizing is pure host arithmetic; bad ranks / worlds / capacities never reach CUDA
    assert lib.mixgrpo_peer_region_bytes(8, 4096) == 256 + 2 * 8 * 4096 * 8 + 2 * 8 * 256 * 8
    assert lib.mixgrpo_peer_region_bytes(17, 16) == 0 and lib.mixgrpo_peer_region_bytes(0, 16) == 0
    regs = (ctypes.c_void_p * 2)(0x1000, 0x2000)
    assert lib.mixgrpo_peer_gather_advantages(None, 0, 2, 64, 1, None, 1, 4, 4, 0, 0, None, 1, None) == -1
    assert lib.mixgrpo_peer_gather_advantages(regs, 2, 2, 64, 1, None, 1, 4, 4, 0, 0, None, 1, None) == -1       # rank >= world
    assert lib.mixgrpo_peer_gather_advantages(regs, 0, 2, 64, 1, None, 3, 4, 4, 0, 0, None, 1, None) == -1       # 3 models, no weights
    assert lib.mixgrpo_peer_gather_advantages(regs, 0, 2, 8, 1, 1, 3, 4, 4, 0, 0, None, 1, None) == -3           # 12 floats > cap 8
    assert lib.mixgrpo_peer_gather_advantages(regs, 0, 2, 64, 1, None, 1, 4, 4, 0, 7, None, 1, None) == -1       # bad mode
    assert lib.mixgrpo_peer_allreduce(regs, 0, 2, 64, 1, 257, 1, None) == -3
    assert lib.mixgrpo_peer_allreduce(regs, 0, 2, 64, None, 4, 1, None) == -1
    assert lib.mixgrpo_peer_region_open(None, None) == -1
    old = lib.mixgrpo_set_tuning(2, 1234)
    assert lib.mixgrpo_set_tuning(2, old) == 1234
    # CTA-shape knobs (ABI v6): key 6 = deferred step launches as 128-thread CTAs (0 never | 1 auto | 2 always), key 7 = CTA size of the
    # backward kernels (128 | 256), key 8 = read-only launch counter; a CTA cap above 2047 per sample no longer fits the packed word
    assert lib.mixgrpo_set_tuning(6, 3) == -1 and lib.mixgrpo_set_tuning(6, -1) == -1
    assert lib.mixgrpo_set_tuning(6, 2) == 1 and lib.mixgrpo_set_tuning(6, 1) == 2
    assert lib.mixgrpo_set_tuning(7, 64) == -1 and lib.mixgrpo_set_tuning(7, 128) == 256 and lib.mixgrpo_set_tuning(7, 256) == 128
    assert lib.mixgrpo_set_tuning(8, 0) >= 0
    assert lib.mixgrpo_set_tuning(0, 2048) == -1 and lib.mixgrpo_set_tuning(0, 2047) == 2047
    # a deferred launch needs the eight-records-per-sample block: the plain workspace size is refused before any launch
    assert lib.mixgrpo_flow_step(1, 1, 1, 8, None, None, 0, 1, 8, None, None, None, 1, 512, 12, 8, ctypes.byref(k), 2, _cabi.FLAG_DEFER_LOGP, None, None) == -3


def test_python_layer_refuses_cpu_tensors():
    import torch
    from mixgrpo_b200 import grpo, ops, sampling_utils as su
    x = torch.zeros(1, 4, 64)
    sig = torch.linspace(1, 0, 5)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        su.flow_grpo_step(x.bfloat16(), x, 0.7, sig, 1, x)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        su.dance_grpo_step(x.bfloat16(), x, 0.7, sig, 1, x, True, True)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        grpo.compute_group_advantages(torch.zeros(4), 4)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        grpo.grpo_loss(torch.zeros(2), torch.zeros(2), torch.zeros(2), 1e-4, 5.0, 0.0, 1, 1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        ops.pack_latents(torch.zeros(1, 16, 4, 4), 1, 16, 4, 4)
    from mixgrpo_b200.peer import PeerExchange
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        PeerExchange(device=torch.device("cpu"))


def test_product_never_imports_oracle():
    """The product package must not route through the oracle (or any CPU fallback)."""
    for f in (ROOT / "mixgrpo_b200").rglob("*.py"):
        src = f.read_text()
        assert "import oracle" not in src and "from oracle" not in src, f


def test_header_is_plain_c_and_the_library_is_callable_from_c(tmp_path):
    """include/mixgrpo_b200.h is valid C99 (no torch, no C++), and a C program can load the library and call it."""
    import shutil
    import subprocess
    from mixgrpo_b200 import _build, _cabi
    if not shutil.which("gcc"):
        pytest.skip("gcc not available")
    lib = _build.build()
    src = tmp_path / "probe.c"
    src.write_text(r'''
#include <dlfcn.h>
#include <stdio.h>
#include "mixgrpo_b200.h"
int main(int argc, char** argv) {
  void* h = dlopen(argv[1], RTLD_NOW | RTLD_LOCAL);
  if (!h) { fprintf(stderr, "%s\n", dlerror()); return 2; }
  int (*abi)(void) = (int (*)(void))dlsym(h, "mixgrpo_abi_version");
  int64_t (*ws)(int64_t, int64_t) = (int64_t (*)(int64_t, int64_t))dlsym(h, "mixgrpo_step_workspace_bytes");
  int64_t (*rb)(int, int64_t) = (int64_t (*)(int, int64_t))dlsym(h, "mixgrpo_peer_region_bytes");
  int (*step)(const void*, int, const float*, int64_t, const void*, const float*, int64_t, float*, int64_t, float*, float*, float*, void*,
              int64_t, int64_t, int64_t, const mixgrpo_step_coefs*, int, unsigned, void*, const mixgrpo_step_ext*) = dlsym(h, "mixgrpo_flow_step");
  mixgrpo_step_coefs k = {0};
  if (!abi || !ws || !rb || !step) return 3;
  printf("%d %lld %lld %d\n", abi(), (long long)ws(12, 262144), (long long)rb(8, 36),
         step(NULL, MIXGRPO_BF16, NULL, 0, NULL, NULL, 0, NULL, 0, NULL, NULL, NULL, NULL, 0, 1, 8, &k, MIXGRPO_SRC_NOISE, 0u, NULL, NULL));
  return abi() == MIXGRPO_ABI_VERSION ? 0 : 4;
}
''')
    exe = tmp_path / "probe"
    r = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic-errors", "-Wno-pedantic", "-I", str(ROOT / "include"), str(src), "-o", str(exe), "-ldl"],
                       capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    r = subprocess.run([str(exe), str(lib)], capture_output=True, text=True)
    assert r.returncode == 0, (r.stdout, r.stderr)
    abi, ws, rb, rc = r.stdout.split()
    assert int(abi) == _cabi.ABI_VERSION and int(ws) == 512 and int(rb) == 256 + 2 * 8 * 36 * 8 + 2 * 8 * 256 * 8 and int(rc) == -1


def test_missing_library_fails_loudly_instead_of_falling_back(monkeypatch, tmp_path):
    """No .so and no way to build it: loading raises (there is nothing else to route to)."""
    from mixgrpo_b200 import _build, _cabi

    def no_nvcc(*a, **k):
        raise RuntimeError("mixgrpo_b200: nvcc not found; cannot build the CUDA library")

    monkeypatch.setattr(_cabi, "_lib", None)
    monkeypatch.setattr(_build, "LIB", tmp_path / "nope" / "libmixgrpo_b200.so")
    monkeypatch.setattr(_build, "STAMP", tmp_path / "nope" / "libmixgrpo_b200.stamp")
    monkeypatch.setattr(_build, "build", no_nvcc)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        _cabi.lib()
    import mixgrpo_b200
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        mixgrpo_b200.load_library()
