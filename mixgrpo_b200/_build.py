"""Builds ``mixgrpo_b200/_lib/libmixgrpo_b200.so`` from ``csrc/*.cu`` with nvcc for sm_100a.

In-tree, explicit nvcc (no JIT cache): the built ``.so`` is git-ignored but travels with the
gpurun snapshot.  A stamp file (hash of sources + flags) makes the build idempotent; a file lock plus
write-to-temp-then-rename makes it safe when several ranks of one job find the stamp stale at the same time
(eight ranks once linked over each other's output: "file too short").
"""
from __future__ import annotations

import contextlib
import fcntl
import hashlib
import os
import shutil
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor
from pathlib import Path

PKG = Path(__file__).resolve().parent
CSRC = PKG / "csrc"
INCLUDE = PKG.parent / "include"
LIBDIR = PKG / "_lib"
LIB = LIBDIR / "libmixgrpo_b200.so"
STAMP = LIBDIR / "libmixgrpo_b200.stamp"
OBJDIR = PKG.parent / "build" / "obj"

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17",
    "-Xcompiler", "-fPIC", "-Xcompiler", "-fvisibility=hidden", "--expt-relaxed-constexpr",
]


def _nvcc() -> str:
    cand = os.environ.get("NVCC") or shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not Path(cand).exists():
        raise RuntimeError("mixgrpo_b200: nvcc not found; cannot build the CUDA library")
    return cand


def _sources():
    return sorted(CSRC.glob("*.cu"))


def source_hash() -> str:
    h = hashlib.sha256()
    for f in sorted(list(CSRC.glob("*.cu")) + list(CSRC.glob("*.cuh")) + [INCLUDE / "mixgrpo_b200.h"]):
        h.update(f.name.encode())
        h.update(f.read_bytes())
    h.update(" ".join(NVCC_FLAGS).encode())
    return h.hexdigest()


def is_fresh() -> bool:
    return LIB.exists() and STAMP.exists() and STAMP.read_text().strip() == source_hash()


@contextlib.contextmanager
def _build_lock():
    """One builder at a time per tree (processes of one torchrun job share it)."""
    LIBDIR.mkdir(parents=True, exist_ok=True)
    with open(LIBDIR / ".build.lock", "w") as fh:
        fcntl.flock(fh, fcntl.LOCK_EX)
        try:
            yield
        finally:
            fcntl.flock(fh, fcntl.LOCK_UN)


def _publish(tmp: Path, dst: Path) -> None:
    os.replace(tmp, dst)                                 # atomic: a concurrent dlopen sees the old file or the new one, never half


def build(force: bool = False, verbose: bool = False) -> Path:
    if not force and is_fresh():
        return LIB
    with _build_lock():
        if not force and is_fresh():                     # another rank built it while this one waited
            return LIB
        return _build_locked(verbose)


def _build_locked(verbose: bool) -> Path:
    nvcc = _nvcc()
    OBJDIR.mkdir(parents=True, exist_ok=True)
    LIBDIR.mkdir(parents=True, exist_ok=True)

    def compile_one(src: Path) -> Path:
        obj = OBJDIR / (src.stem + ".o")
        cmd = [nvcc, *NVCC_FLAGS, "-I", str(INCLUDE), "-c", str(src), "-o", str(obj)]
        if verbose:
            print(" ".join(cmd), file=sys.stderr)
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed for {src.name}:\n{r.stdout}\n{r.stderr}")
        return obj

    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        objs = list(ex.map(compile_one, _sources()))
    tmp = LIBDIR / f".libmixgrpo_b200.{os.getpid()}.so.tmp"
    cmd = [nvcc, "-shared", "-o", str(tmp), *map(str, objs), "-gencode", "arch=compute_100a,code=sm_100a",
           "-cudart", "static"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    if STAMP.exists():
        STAMP.unlink()                                   # never a fresh-looking stamp next to a library of other sources
    _publish(tmp, LIB)
    tmp_stamp = LIBDIR / f".stamp.{os.getpid()}.tmp"
    tmp_stamp.write_text(source_hash() + "\n")
    _publish(tmp_stamp, STAMP)
    return LIB


# ------------------------------------------------------------------------------------------ compiled Python binding
BIND_SRC = PKG / "csrc_bind" / "torch_bind.cpp"
BIND_LIB = LIBDIR / "_torchbind.so"
BIND_STAMP = LIBDIR / "_torchbind.stamp"


def _bind_hash() -> str:
    import torch
    h = hashlib.sha256()
    h.update(BIND_SRC.read_bytes())
    h.update((INCLUDE / "mixgrpo_b200.h").read_bytes())
    h.update(torch.__version__.encode())
    h.update(sys.version.encode())
    return h.hexdigest()


def binding_is_fresh() -> bool:
    return BIND_LIB.exists() and BIND_STAMP.exists() and BIND_STAMP.read_text().strip() == _bind_hash()


def build_binding(force: bool = False, verbose: bool = False) -> Path:
    """g++ -shared of csrc_bind/torch_bind.cpp against libtorch / libtorch_python and the in-tree libmixgrpo_b200.so
    (rpath $ORIGIN): the pybind11 + torch::autograd layer over the SAME C ABI the ctypes binding calls.  In-tree and
    stamp-guarded like the CUDA library (no JIT cache: the built .so travels with the gpurun snapshot)."""
    if not force and binding_is_fresh():
        return BIND_LIB
    build()                                              # the library it links against
    with _build_lock():
        if not force and binding_is_fresh():
            return BIND_LIB
        return _build_binding_locked(verbose)


def _build_binding_locked(verbose: bool) -> Path:
    import sysconfig

    import torch
    from torch.utils import cpp_extension as ce
    cxx = os.environ.get("CXX") or shutil.which("g++") or "g++"
    tmp = LIBDIR / f"._torchbind.{os.getpid()}.so.tmp"
    cuda_home = Path(_nvcc()).resolve().parent.parent
    inc = [*ce.include_paths(), sysconfig.get_paths()["include"], str(cuda_home / "include"), str(INCLUDE)]
    torch_lib = str(Path(torch.__file__).resolve().parent / "lib")
    cmd = [cxx, "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-DTORCH_EXTENSION_NAME=_torchbind", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", *[f"-I{i}" for i in inc], str(BIND_SRC), "-o", str(tmp),
           f"-L{LIBDIR}", "-lmixgrpo_b200", f"-L{torch_lib}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
           "-Wl,-rpath,$ORIGIN", f"-Wl,-rpath,{torch_lib}"]
    if verbose:
        print(" ".join(cmd), file=sys.stderr)
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"g++ failed for {BIND_SRC.name}:\n{r.stdout}\n{r.stderr}")
    if BIND_STAMP.exists():
        BIND_STAMP.unlink()
    _publish(tmp, BIND_LIB)
    tmp_stamp = LIBDIR / f".bindstamp.{os.getpid()}.tmp"
    tmp_stamp.write_text(_bind_hash() + "\n")
    _publish(tmp_stamp, BIND_STAMP)
    return BIND_LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose=True))
    print(build_binding(force="--force" in sys.argv, verbose=True))
