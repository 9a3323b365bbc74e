"""GPU: MCTS.run_multi (device-resident loop, CUDA-graph replay, and the compat path around a model that
only speaks the reference's NetworkOutput contract) against the CPU oracle fed the same network outputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu

N, A, S = 96, 20, 24


def _setup(amp="none", small=False):
    from hanabizero_b200.hanabi_env import HanabiVecEnv
    from hanabizero_b200.mcts import SearchConfig
    from hanabizero_b200.model import MuZeroNet, MuZeroNetFull
    dev = torch.device("cuda")
    torch.manual_seed(0)
    name, a = ("Hanabi-Small", 11) if small else ("Hanabi-Full", 20)
    env = HanabiVecEnv(N, name, np.arange(N))
    g, _, legal = env.reset_all()
    model = (MuZeroNet if small else MuZeroNetFull)(env.global_dim, a).randomize_heads().to(dev).eval()
    with torch.no_grad():
        _, logits, hidden = model.initial_inference_device(g)
    noise = np.random.default_rng(3).dirichlet([0.3] * a, N).astype(np.float32)
    return dev, model, logits, hidden, legal, noise, SearchConfig(num_simulations=S, amp_type=amp), a


def _record(model):
    rec = []
    orig = model.recurrent_inference_device

    def wrapped(h, act):
        out = orig(h, act)
        if h.shape[0] == N:   # skip the 2-row dtype probe MCTS makes when it builds its workspace
            rec.append(tuple(o.detach().clone() for o in out[:3]) + (h.detach().clone(), act.detach().clone()))
        return out

    model.recurrent_inference_device = wrapped
    return rec, orig


def _oracle_replay(rec, cfg, noise, logits, legal, a):
    from oracle import loader as L
    cpu = L.oracle_tree(N, a, S, cfg.value_delta_max)
    cpu.prepare(0.25, noise, np.zeros(N, np.float32), logits.float().cpu().numpy(), legal.cpu().numpy().astype(np.int32))
    trace = []
    for x, (v, rw, lg, _, _) in enumerate(rec, start=1):
        trace.append(cpu.traverse(cfg.pb_c_base, cfg.pb_c_init, cfg.discount))
        cpu.backprop(x, cfg.discount, rw.float().cpu().numpy(), v.float().cpu().numpy(),
                     np.nan_to_num(lg.float().cpu().numpy(), nan=0.0))
    return cpu, trace


@pytest.mark.parametrize("amp,small", [("none", False), ("torch_amp", False), ("none", True)])
def test_run_multi_matches_oracle_given_same_network_outputs(amp, small):
    from hanabizero_b200 import cytree
    from hanabizero_b200.mcts import MCTS
    dev, model, logits, hidden, legal, noise, cfg, a = _setup(amp, small)
    rec, orig = _record(model)
    roots = cytree.Roots(N, a, S)
    roots.prepare(0.25, noise, [0.0] * N, logits, legal)
    MCTS(cfg).run_multi(roots, model, hidden, use_graph=False)
    model.recurrent_inference_device = orig
    assert len(rec) == S - 1                               # the last iteration is skipped (core/mcts.py:25-26)
    cpu, trace = _oracle_replay(rec, cfg, noise, logits, legal, a)
    visits, values = roots.get_stats_tensors()
    ov, oval, _ = cpu.stats()
    assert (visits.cpu().numpy() == ov).all()
    np.testing.assert_allclose(values.cpu().numpy(), oval, rtol=1e-5)
    # the batches handed to the network are the parents' hidden states and the last actions
    pool = [hidden.to(rec[0][3].dtype)] + [None] * S
    for x, ((v, rw, lg, h, act), (ix, iy, la)) in enumerate(zip(rec, trace), start=1):
        assert (act.view(-1).cpu().numpy() == la).all()
        want = torch.stack([pool[int(i)][int(j)] for i, j in zip(ix, iy)])
        assert torch.equal(h, want)
        with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=amp == "torch_amp"):
            pool[x] = orig(h, act)[3]


def test_graph_replay_equals_eager_and_is_repeatable():
    from hanabizero_b200 import cytree
    from hanabizero_b200.mcts import MCTS
    dev, model, logits, hidden, legal, noise, cfg, a = _setup()
    mcts = MCTS(cfg)
    outs = []
    for it in range(4):     # 0: eager, 1: capture + replay, 2-3: replay; fresh Roots each time like selfplay
        roots = cytree.Roots(N, a, S)
        nz = noise if it != 3 else np.roll(noise, 1, axis=0)
        roots.prepare(0.25, nz, [0.0] * N, logits, legal)
        mcts.run_multi(roots, model, hidden)
        v, val = roots.get_stats_tensors()
        outs.append((v.cpu().clone(), val.cpu().clone(), roots.export(S)["visits"].cpu().clone()))
        del roots
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][0], outs[2][0])
    assert torch.equal(outs[0][1], outs[1][1]) and torch.equal(outs[0][2], outs[2][2])
    assert not torch.equal(outs[0][0], outs[3][0])         # different noise -> the replay really re-ran the search
    assert (outs[3][0].sum(1) == S - 1).all()


def test_compat_path_with_reference_style_model():
    """A model exposing only recurrent_inference -> NetworkOutput of numpy arrays (core/model.py:74-84)."""
    from hanabizero_b200 import cytree
    from hanabizero_b200.mcts import MCTS
    from hanabizero_b200.model import NetworkOutput
    dev, model, logits, hidden, legal, noise, cfg, a = _setup()

    class RefStyle:
        def __init__(self, m):
            self.m = m

        def eval(self):
            return self

        def recurrent_inference(self, h, act):
            v, r, lg, st = self.m.recurrent_inference_device(h, act)
            return NetworkOutput(v.view(-1, 1).cpu().numpy(), r.view(-1, 1).cpu().numpy(), lg.cpu().numpy(), st.cpu().numpy())

    r1, r2 = cytree.Roots(N, a, S), cytree.Roots(N, a, S)
    for r in (r1, r2):
        r.prepare(0.25, noise, [0.0] * N, logits, legal)
    MCTS(cfg).run_multi(r1, RefStyle(model), hidden.cpu().numpy())
    MCTS(cfg).search(r2, model, hidden, use_graph=False)
    assert r1.get_distributions() == r2.get_distributions()
    np.testing.assert_allclose(r1.get_values(), r2.get_values(), rtol=1e-5)
