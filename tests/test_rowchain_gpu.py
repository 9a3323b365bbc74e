"""GPU: the row-block resident network executor (hz_rowchain_*, hanabizero_b200/csrc/hz_rowchain.cu) computes the
same recurrent_inference (/root/reference/core/model.py:74-84) as the seven-launch cuBLASLt chain and as the nn.Module.
Floating point (fp16 storage, fp32 accumulation): the tolerance is stated per assertion; the tree step that consumes
these outputs stays bit-exact given its inputs."""
import numpy as np
import pytest
import torch

pytestmark = pytest.mark.gpu


def _net(seed=1, bn_stats=True):
    from hanabizero_b200.model import MuZeroNetFull
    torch.manual_seed(seed)
    net = MuZeroNetFull(785, 20).randomize_heads(std=0.05).to("cuda")
    if bn_stats:
        for m in net.modules():
            if isinstance(m, torch.nn.BatchNorm1d):
                m.running_mean.normal_(0, 0.2); m.running_var.uniform_(0.5, 1.5)
                m.weight.data.uniform_(0.5, 1.5); m.bias.data.normal_(0, 0.2)
    return net.eval()


def _fill(ch, plan, seed):
    g = torch.Generator(device="cuda").manual_seed(seed)
    ch.x0.zero_()
    ch.x0[:, :plan.F].copy_((torch.randn(ch.n, plan.F, device="cuda", generator=g).clamp_min(0) * 0.7).half())
    a = torch.randint(0, plan.A, (ch.n, 1), device="cuda", generator=g)
    ch.x0[:, plan.F:].scatter_(1, a, 1.0)
    return a


@pytest.mark.parametrize("n", [1, 128, 300, 1000, 4096, 19000])
def test_row_executor_matches_library_chain(n):
    """Every output of the one-launch executor against the library chain on the same buffers: row counts below, at and
    across the 128-row block size, one CTA looping over several row blocks (19000 rows > 148 blocks), rows without an
    action.  Both round fp32 sums to fp16 after every layer; they differ by summation order only: <= 2 fp16 ulp of the
    largest activation (asserted as 3e-3 absolute on values of magnitude ~4), and both sit equally close to float64."""
    from hanabizero_b200.plan import BoundChain
    net = _net()
    plan = net.recurrent_plan(torch.float16)
    ch = BoundChain(plan, n)
    assert ch.rows_supported()
    _fill(ch, plan, 5 + n)
    if n >= 300:
        ch.x0[7, plan.F:] = 0          # a row without any action column set: fc1 sees the state alone
    st = torch.cuda.current_stream().cuda_stream
    ch.bind_state(ch.state)
    ch.run(st)
    s_lib, o_lib = ch.state.clone(), ch.out.clone()
    other = torch.zeros_like(ch.state)
    ch.out.zero_()
    ch.set_executor("rows")
    ch.bind_state(other)               # the search loop re-points the state output at pool[x] every simulation
    ch.run(st)
    torch.cuda.synchronize()
    assert torch.count_nonzero(ch.state - s_lib) == 0, "the unbound state buffer must be left alone"
    torch.testing.assert_close(other.float(), s_lib.float(), rtol=0, atol=3e-3)
    torch.testing.assert_close(ch.out.float(), o_lib.float(), rtol=0, atol=1e-3)
    # float64 evaluation of the same folded weights: the two executors are equally far from it
    w = {k: v.double() for k, v in plan._w.items()}
    x, H = ch.x0.double(), plan.H
    relu = torch.relu
    y1 = relu(x @ w["W1"].T + w["b1"]); y2 = relu(y1 @ w["W2"].T + w["b2"]); s = relu(y2 @ w["W3"].T + w["b3"] + x[:, :plan.F])
    h1 = relu(s @ w["Wh1"].T + w["bh1"])
    a1 = relu(h1[:, :H] @ w["WB2"][0].T + w["bB2"][0]); v2 = relu(h1[:, H:2 * H] @ w["WB2"][1].T + w["bB2"][1])
    r2 = relu(h1[:, 2 * H:] @ w["WB2"][2].T + w["bB2"][2]); a2 = relu(a1 @ w["Wa2"].T + w["ba2"] + h1[:, :H])
    o64 = torch.stack([t @ w["WB3"][i].T + w["bB3"][i] for i, t in enumerate((v2, r2, a2))])
    e_rows, e_lib = (ch.out.double() - o64).abs().max().item(), (o_lib.double() - o64).abs().max().item()
    assert e_rows <= 2 * e_lib + 1e-4, (e_rows, e_lib)
    s_rows, s_l = (other.double() - s).abs().max().item(), (s_lib.double() - s).abs().max().item()
    assert s_rows <= 2 * s_l + 1e-3, (s_rows, s_l)
    # switching back restores the library path on the same buffers
    ch.set_executor("library")
    ch.bind_state(ch.state)
    ch.state.zero_()
    ch.run(st)
    torch.cuda.synchronize()
    assert torch.equal(ch.state, s_lib)


def test_row_executor_follows_weight_updates_and_graph_replay():
    """The executor reads the plan's fixed weight buffers: refresh() after a parameter update is seen without
    re-creating anything, also by a CUDA graph captured before the update."""
    from hanabizero_b200.plan import BoundChain
    net = _net(seed=3)
    plan = net.recurrent_plan(torch.float16)
    ch = BoundChain(plan, 384)
    _fill(ch, plan, 9)
    ch.set_executor("rows")
    ch.bind_state(ch.state)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        ch.run(side.cuda_stream)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g, stream=side):
        ch.run(torch.cuda.current_stream().cuda_stream)
    g.replay(); torch.cuda.synchronize()
    before = ch.out.clone()
    with torch.no_grad():
        for p in net.parameters():
            p.mul_(1.05)
    assert plan.refresh()
    g.replay(); torch.cuda.synchronize()
    after_rows = ch.out.clone()
    assert not torch.equal(before, after_rows)
    ch.set_executor("library")
    ch.run(torch.cuda.current_stream().cuda_stream)
    torch.cuda.synchronize()
    torch.testing.assert_close(after_rows.float(), ch.out.float(), rtol=0, atol=1e-3)


def test_search_with_row_executor_selects_the_same_actions():
    """A whole search (graph replay) with executor='rows' against the same search on the library chain: same visit
    totals, >= 0.99 of the root actions identical (the trees are bit-exact functions of the network outputs, which
    differ in the last fp16 bits), root values within 1e-2."""
    from hanabizero_b200 import cytree
    from hanabizero_b200.mcts import MCTS
    from hanabizero_b200.model import MuZeroNetFull
    N, A, S = 1024, 20, 30

    class Cfg:
        pb_c_base, pb_c_init, discount, value_delta_max, num_simulations, amp_type = 19652, 1.25, 0.997, 0.01, S, "torch_amp"

    torch.manual_seed(0)
    model = MuZeroNetFull(785 * 4, A).randomize_heads(seed=0).to("cuda").eval()
    rng = np.random.default_rng(4)
    obs = torch.from_numpy((rng.random((N, 785 * 4)) < 0.2).astype(np.float32)).cuda()
    with torch.no_grad():
        _, logits, hidden = model.initial_inference_device(obs)
    legal = torch.from_numpy((rng.random((N, A)) < 0.7).astype(np.int32)).cuda()
    legal[:, 0] = 1
    noise = torch.from_numpy(rng.dirichlet([0.3] * A, N).astype(np.float32)).cuda()
    stats = {}
    for ex in ("library", "rows"):
        mcts = MCTS(Cfg)
        for _ in range(3):     # eager, capture, replay
            roots = cytree.Roots(N, A, S)
            roots.prepare(0.25, noise, torch.zeros(N, device="cuda"), logits.float(), legal)
            mcts.run_multi(roots, model, hidden, executor=ex)
        v, val = roots.get_stats_tensors()
        stats[ex] = (v.clone(), val.clone())
    (v0, val0), (v1, val1) = stats["library"], stats["rows"]
    assert int(v1.sum()) == N * (S - 1) and torch.equal(v0.sum(1), v1.sum(1))
    agree = (v0.argmax(1) == v1.argmax(1)).float().mean().item()
    assert agree >= 0.99, agree
    assert (val0 - val1).abs().max().item() < 1e-2
