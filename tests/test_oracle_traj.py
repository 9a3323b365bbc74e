"""CPU: the test-side ports of the reference's self-play bookkeeping (helpers.RefGameHistory, helpers.ref_select_action:
core/game.py:73-204, core/utils.py:280-295, core/selfplay_worker.py:29-39) against fixtures written by the reference's
own Python (tests/golden/make_golden.py) — so the ports the GPU tests compare the kernels with are themselves pinned."""
import numpy as np
import pytest

from helpers import RefGameHistory, golden_files, load_golden, ref_select_action, unpack_traj_golden


@pytest.mark.parametrize("name", golden_files("traj_"))
def test_ports_reproduce_reference_trajectories(name):
    g = load_golden(name)
    obs, eps = unpack_traj_golden(g)
    stack = int(g["stack"])
    done_eps, gh, legal = [], None, None
    for t in range(len(g["action"])):
        if g["action"][t] < 0:
            gh = RefGameHistory(stack)
            gh.init([obs[t]] * stack, g["legal"][t].astype(np.float64))
            legal = g["legal"][t]
            continue
        action, _, mutated = ref_select_action(g["visits"][t].tolist(), 1, True, legal, 0.0)
        assert action == int(g["action"][t])
        gh.store_search_stats(mutated, float(g["value"][t]))
        gh.append(action, obs[t], int(g["reward"][t]), g["legal"][t].astype(np.float64))
        legal = g["legal"][t]
        if g["done"][t]:
            gh.game_over()
            gh.put()
            done_eps.append(gh)
    assert len(done_eps) == len(eps) > 0
    for gh, want in zip(done_eps, eps):
        assert (gh.child_visits == want["vis"]).all() and (gh.root_values == want["root"]).all()
        assert (gh.actions == want["a"]).all() and (gh.rewards == want["r"]).all()
        assert (gh.obs_history == want["o"]).all() and (gh.legal_actions == want["la"]).all()
