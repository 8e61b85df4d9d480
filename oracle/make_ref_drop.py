"""TEST / BENCH INFRASTRUCTURE ONLY — refreshes the git-ignored ``baseline/_ref`` drop of the reference's path files.

The reference (zqqqqz2000/MixGRPO) is pure Python, so there is nothing to compile: "building" the real reference for the
GPU box means placing the few files of THIS path — unmodified, where ``oracle/ref_loader.py`` / ``oracle/ref_extract.py``
execute them from — under ``baseline/_ref/`` (git-ignored, so no reference source ever enters the history; NOT
gpurun-ignored, so it travels with the snapshot like a built ``.so``).  ``/root/reference`` does not exist on the GPU box;
with the drop, ``bench.py --impl reference`` and the ``cpu_baseline`` leg time the reference's own code there
(``cpu_baseline.kind == "reference"``); without it they fall back to the oracle restatement (``"port"``).

    python -m oracle.make_ref_drop            # called by __graft_entry__.build() when /root/reference is present
"""
from __future__ import annotations

import filecmp
import shutil
from pathlib import Path

REPO = Path(__file__).resolve().parent.parent
SRC = Path("/root/reference")
DST = REPO / "baseline" / "_ref"
FILES = (
    "fastvideo/utils/sampling_utils.py",            # the sampler / log-prob operators (SU)
    "fastvideo/utils/grpo_states.py",               # the sliding-window scheduler
    "fastvideo/models/reward_model/utils.py",       # balance_pos_neg
    "fastvideo/train_grpo_flux.py",                 # never imported: ref_extract.py cuts TR:440-501 / TR:560-583 out with ast
)


def refresh(verbose: bool = False) -> int:
    """Copy the path's files when the reference tree is here; returns how many files the drop now holds."""
    if not SRC.is_dir():
        return sum((DST / f).is_file() for f in FILES)
    n = 0
    for f in FILES:
        s, d = SRC / f, DST / f
        if not s.is_file():
            continue
        d.parent.mkdir(parents=True, exist_ok=True)
        if not d.is_file() or not filecmp.cmp(s, d, shallow=False):
            shutil.copyfile(s, d)
            if verbose:
                print(f"[ref-drop] {f}")
        n += 1
    return n


if __name__ == "__main__":
    print(f"[ref-drop] {refresh(verbose=True)} file(s) under {DST}")
