"""mixgrpo_b200.grpo_states.GRPOTrainingStates vs the reference class (live when present) and vs golden traces
recorded from the reference (tests/golden/grpo_states_traces.json, tools/make_golden.py)."""
import json
from pathlib import Path

import pytest

from mixgrpo_b200.grpo_states import GRPOTrainingStates

TRACES = json.loads((Path(__file__).parent / "golden" / "grpo_states_traces.json").read_text())


@pytest.mark.parametrize("case", range(len(TRACES)))
def test_against_golden_traces(case):
    cfg, trace = TRACES[case]["config"], TRACES[case]["trace"]
    s = GRPOTrainingStates(**cfg)
    for it, want in enumerate(trace):
        assert [int(x) for x in s.get_current_timesteps()] == want["t"], (cfg, it)
        assert bool(s.is_training_complete()) == want["done"]
        if cfg.get("sample_strategy") == "random":
            s.update_iteration(seed=1000 + it)
        else:
            s.update_iteration()


def test_against_live_reference():
    from oracle import ref_loader
    st = ref_loader.load_states()
    if st is None:
        pytest.skip("reference tree not present")
    import itertools
    for strat, ov, step, rb, ipg, gs in itertools.product(("progressive", "decay", "exp_decay"), (False, True), (0, 1, 3), (False, True), (1, 4, 25), (1, 4)):
        kw = dict(iters_per_group=ipg, group_size=gs, max_timesteps=23, sample_strategy=strat, prog_overlap=ov, prog_overlap_step=step, roll_back=rb)
        a, b = GRPOTrainingStates(**kw), st.GRPOTrainingStates(**kw)
        for _ in range(200):
            assert a.get_current_timesteps() == b.get_current_timesteps()
            assert a.is_training_complete() == b.is_training_complete()
            assert (a.cur_timestep, a.cur_iter_in_group) == (b.cur_timestep, b.cur_iter_in_group)
            a.update_iteration()
            b.update_iteration()
    with pytest.raises(ValueError):
        GRPOTrainingStates(1, 1, 5, sample_strategy="nope").update_iteration()


def test_shipping_config_window_sequence():
    """finetune_flux_grpo_MixGRPO.sh: steps 25 -> max_timesteps 23, window 4, stride 1, 25 iterations per window."""
    s = GRPOTrainingStates(iters_per_group=25, group_size=4, max_timesteps=23, prog_overlap=True, prog_overlap_step=1)
    seen = []
    for _ in range(25 * 30):
        seen.append(tuple(s.get_current_timesteps()))
        s.update_iteration()
    assert seen[0] == (0, 1, 2, 3) and seen[24] == (0, 1, 2, 3) and seen[25] == (1, 2, 3, 4)
    assert seen[25 * 19] == (19, 20, 21, 22) and seen[25 * 20] == (20, 21, 22)     # clipped at max_timesteps
    assert seen[-1] == ()                                                          # parked at max_timesteps
