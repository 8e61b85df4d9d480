"""mixgrpo_b200 — B200-native (sm_100a) implementation of MixGRPO's data-parallel rollout and
policy-update hot path: fused sampler step + transition log-prob, group-relative advantages and the
clipped-ratio GRPO loss, behind the reference's own operator names.

    from mixgrpo_b200.sampling_utils import flow_grpo_step, dance_grpo_step, dpm_step, run_sample_step, sd3_time_shift
    from mixgrpo_b200.grpo import compute_group_advantages, grpo_loss, gather_rewards
    from mixgrpo_b200.grpo_states import GRPOTrainingStates

Importing the package does not load CUDA; the first operator call loads (and if necessary builds)
``_lib/libmixgrpo_b200.so`` and fails loudly when that is impossible.  There is no CPU fallback.
"""
from . import _build, _cabi  # noqa: F401

__version__ = "0.1.0"


def library_path() -> str:
    """Path of the C-ABI shared library (include/mixgrpo_b200.h)."""
    return _cabi.library_path()


def load_library():
    """Force-load the CUDA library now (raises RuntimeError if it is missing and cannot be built)."""
    return _cabi.lib()
