#!/usr/bin/env python
"""Benchmark of the HanabiZero self-play hot path (BASELINE.json metric: MCTS simulations/s and
Hanabi env steps/s on 1/2/4/8 B200 next to the reference's host-CPU implementation).

    python bench.py --gpus N --steps K --warmup W            # our arm (one rank per GPU under torchrun)
    python bench.py --impl reference --gpus N --steps K ...  # the reference's CPU path on the host cores

A "step" is one batched search: Roots.prepare + MCTS.run_multi (num_simulations-1 simulations, the
PyTorch network included, random-init weights with re-drawn heads) + root statistics, over
`--trees` Hanabi-Full trees per GPU (weak scaling: the root batch is sharded, one model replica per
GPU, one NCCL all_gather of the final statistics per search).  `value` = simulations/s of the whole
job with inputs resident in HBM; `e2e` = the same through the list/numpy drop-in API with pinned host
inputs copied in and statistics copied out inside the timed region.  The Hanabi env is timed the
same way and reported in the `env` object.  One JSON line is printed by rank 0.
"""
import argparse
import json
import multiprocessing as mp
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import numpy as np  # noqa: E402

CONST = dict(pb_c_base=19652, pb_c_init=1.25, discount=0.999, delta=0.006, frac=0.25)
F_HIDDEN = 512


def parse_args():
    p = argparse.ArgumentParser()
    p.add_argument("--gpus", type=int, default=1)
    p.add_argument("--steps", type=int, default=20)
    p.add_argument("--warmup", type=int, default=3)
    p.add_argument("--impl", default="ours", choices=["ours", "reference"])
    p.add_argument("--trees", type=int, default=4096, help="trees (and games) per GPU")
    p.add_argument("--sims", type=int, default=50)
    p.add_argument("--stack", type=int, default=4)
    p.add_argument("--amp", default="torch_amp", choices=["torch_amp", "none"])
    p.add_argument("--mdp", default="global", choices=["global", "local"], help="observation fed to the network")
    p.add_argument("--env-steps", type=int, default=200, help="env steps per timed env pass")
    p.add_argument("--no-cpu-baseline", action="store_true")
    p.add_argument("--no-graph", action="store_true")
    return p.parse_args()


def b_sim(A, D, s, F):
    """Algorithmic bytes of one simulation of one tree."""
    return D * (16 * A + 32) + 22 * A + 20 * (D + 1) + 12 * s + 32 + 8 * F


B_ENV_STEP = 2 * 192 + 4 * 785 + 4 * 20 + 12  # state r/w + fp32 global obs + legal + r/d/s


# ======================================================================================================
# CPU reference arm (test infrastructure: the ONLY place besides tests/smoke that executes oracle/)
# ======================================================================================================
def _ref_tree_worker(args):
    """One reference actor: the loop of core/mcts.py:24-55 around the reference's own cytree module
    (oracle/_ref, stock build) with pre-generated network outputs instead of the GPU model —
    Python-list marshalling, host hidden-state gather and .tolist() included, as the reference pays."""
    n, A, S, seed, reps = args
    import importlib.util
    ref_dir = os.path.join(ROOT, "oracle", "_ref")
    so = [f for f in os.listdir(ref_dir) if f.startswith("cytree.") and f.endswith(".so")] if os.path.isdir(ref_dir) else []
    kind = "reference"
    if so:
        spec = importlib.util.spec_from_file_location("cytree", os.path.join(ROOT, "oracle", "_ref", so[0]))
        tree = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(tree)
    else:
        tree, kind = None, "port"
    rng = np.random.default_rng(seed)
    logits0 = rng.standard_normal((n, A)).astype(np.float32)
    noise = rng.dirichlet([0.3] * A, n).astype(np.float32)
    mask = np.ones((n, A), np.float64)
    hidden_roots = rng.standard_normal((n, F_HIDDEN)).astype(np.float32)
    sim_hidden = rng.standard_normal((n, F_HIDDEN)).astype(np.float32)
    sim_out = [(rng.standard_normal((n, 1)).astype(np.float32), rng.standard_normal((n, 1)).astype(np.float32),
                rng.standard_normal((n, A)).astype(np.float32)) for _ in range(4)]
    t0 = time.perf_counter()
    for _ in range(reps):
        if tree is not None:
            roots = tree.Roots(n, A, S)
            roots.prepare(CONST["frac"], noise.tolist(), [0.0] * n, logits0.tolist(), [m for m in mask])
            pool = [hidden_roots]
            mm = tree.MinMaxStatsList(n)
            mm.set_delta(CONST["delta"])
            for x in range(1, S):
                results = tree.ResultsWrapper(n)
                ix, iy, la = tree.multi_traverse(roots, CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"], mm, results)
                hs = np.asarray([pool[a][b] for a, b in zip(ix, iy)])       # core/mcts.py:31-33
                la = np.asarray(la)
                value, reward, pol = sim_out[x % 4]
                reward_pool = reward.reshape(-1).tolist()
                value_pool = value.reshape(-1).tolist()
                pol = pol.copy()
                pol[np.isnan(pol)] = 0.0
                pool.append(sim_hidden)
                tree.multi_back_propagate(x, CONST["discount"], reward_pool, value_pool, pol.tolist(), mm, results)
            roots.get_distributions(); roots.get_values()
        else:
            from oracle import loader as L
            eng = L.oracle_tree(n, A, S, CONST["delta"])
            eng.prepare(CONST["frac"], noise, np.zeros(n, np.float32), logits0, mask.astype(np.int32))
            for x in range(1, S):
                eng.traverse(CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"])
                value, reward, pol = sim_out[x % 4]
                eng.backprop(x, CONST["discount"], reward.reshape(-1), value.reshape(-1), pol)
            eng.stats()
    return time.perf_counter() - t0, kind


def _ref_env_worker(args):
    n_steps, seed = args
    from oracle import loader as L
    kind = "reference" if L.have_ref() else "port"
    g = L.ref_hanabi(0, seed) if kind == "reference" else L.oracle_hanabi(0, seed)
    t0 = time.perf_counter()
    g.play(n_steps, seed + 1)
    return time.perf_counter() - t0, kind


def cpu_reference(trees, A, S, cores, reps=1, env_steps_per_core=20000):
    """Reference CPU path on `cores` processes (mirrors num_actors: the reference's only parallelism)."""
    ctx = mp.get_context("spawn")
    per = [trees // cores + (1 if i < trees % cores else 0) for i in range(cores)]
    per = [p for p in per if p > 0]
    with ctx.Pool(len(per)) as pool:
        t0 = time.perf_counter()
        res = pool.map(_ref_tree_worker, [(p, A, S, 100 + i, reps) for i, p in enumerate(per)])
        wall = time.perf_counter() - t0
        wall = max(max(r[0] for r in res), 1e-9)
        sims_s = trees * (S - 1) * reps / wall
        eres = pool.map(_ref_env_worker, [(env_steps_per_core, i) for i in range(len(per))])
        env_s = env_steps_per_core * len(per) / max(r[0] for r in eres)
    return dict(sims_per_s=sims_s, env_steps_per_s=env_s, kind=res[0][1], env_kind=eres[0][1], cores=len(per),
                wall_s=wall)


def reference_with_model(trees, A, S):
    """Informational: the loop of core/mcts.py:24-55 exactly as the reference runs it — its own cytree on the
    host, the PyTorch network on the GPU, Python-list marshalling, host hidden-state gather, H2D of the batch
    and D2H of the outputs EVERY simulation (one actor process).  Returns simulations/s or None."""
    try:
        import importlib.util
        import torch
        if not torch.cuda.is_available():
            return None
        ref_dir = os.path.join(ROOT, "oracle", "_ref")
        so = [f for f in os.listdir(ref_dir) if f.startswith("cytree.") and f.endswith(".so")] if os.path.isdir(ref_dir) else []
        if not so:
            return None
        spec = importlib.util.spec_from_file_location("cytree", os.path.join(ROOT, "oracle", "_ref", so[0]))
        tree = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(tree)
        from hanabizero_b200.model import MuZeroNetFull
        torch.manual_seed(0)
        model = MuZeroNetFull(785 * 4, A).randomize_heads(seed=0).cuda().eval()
        rng = np.random.default_rng(0)
        obs = torch.from_numpy((rng.random((trees, 785 * 4)) < 0.2).astype(np.float32)).cuda()
        best = None
        for rep in range(2):
            with torch.no_grad():
                with torch.autocast("cuda", dtype=torch.float16):
                    out = model.initial_inference(obs)
                torch.cuda.synchronize()
                t0 = time.perf_counter()
                roots = tree.Roots(trees, A, S)
                noises = [rng.dirichlet([0.3] * A).astype(np.float32).tolist() for _ in range(trees)]
                roots.prepare(CONST["frac"], noises, out.reward, out.policy_logits.tolist(), [np.ones(A) for _ in range(trees)])
                pool = [out.hidden_state]
                mm = tree.MinMaxStatsList(trees)
                mm.set_delta(CONST["delta"])
                for x in range(1, S):
                    results = tree.ResultsWrapper(trees)
                    ix, iy, la = tree.multi_traverse(roots, CONST["pb_c_base"], CONST["pb_c_init"], CONST["discount"], mm, results)
                    hs = torch.from_numpy(np.asarray([pool[a][b] for a, b in zip(ix, iy)])).to("cuda")
                    la = torch.from_numpy(np.asarray(la)).to("cuda").unsqueeze(1).long()
                    with torch.autocast("cuda", dtype=torch.float16):
                        no = model.recurrent_inference(hs.float(), la)
                    pol = no.policy_logits
                    pol[np.isnan(pol)] = 0.0
                    pool.append(no.hidden_state)
                    tree.multi_back_propagate(x, CONST["discount"], no.reward.reshape(-1).tolist(),
                                              no.value.reshape(-1).tolist(), pol.tolist(), mm, results)
                roots.get_distributions(); roots.get_values()
                dt = time.perf_counter() - t0
            best = dt if best is None else min(best, dt)
        return trees * (S - 1) / best
    except Exception as e:  # informational only
        return f"unavailable: {type(e).__name__}: {e}"


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    A, S = 20, args.sims
    trees = args.trees * args.gpus
    times = []
    out = None
    for i in range(args.warmup + args.steps):
        out = cpu_reference(trees, A, S, cores, reps=1, env_steps_per_core=5000)
        if i >= args.warmup:
            times.append(out)
    val = statistics.mean(t["sims_per_s"] for t in times)
    env = statistics.mean(t["env_steps_per_s"] for t in times)
    sample = (f"{trees} Hanabi-Full trees x {S - 1} simulations per step split over {out['cores']} processes; "
              "reference cytree driven like core/mcts.py with pre-generated network outputs (no model time); "
              f"env: {5000 * out['cores']} steps of reference libhanabi (apply+deal+2x observe/encode), random legal play")
    line = {
        "impl": "reference", "metric": "mcts_simulations_per_sec", "value": val, "unit": "simulations/s",
        "n_gpus": args.gpus, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": 1e3 * trees * (S - 1) / val, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"Hanabi-Full 2p, {args.trees} trees/GPU x {S} simulations ({S - 1} executed)",
                   "trees_total": trees, "actions": A},
        "cpu_baseline": {"value": val, "unit": "simulations/s", "cores": out["cores"], "kind": out["kind"], "sample": sample},
        "e2e": {"value": val, "unit": "simulations/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "env": {"metric": "hanabi_env_steps_per_sec", "value": env, "unit": "steps/s", "kind": out["env_kind"],
                "cores": out["cores"]},
        "gpu_launches": 0,
        "reference_with_model_on_gpu": {
            "value": reference_with_model(min(args.trees, 1024), A, S), "unit": "simulations/s",
            "what": "informational: one reference actor, its own cytree on the host + the PyTorch network on cuda:0 with "
                    "per-simulation H2D/D2H and list marshalling exactly as core/mcts.py:24-55 does, "
                    f"{min(args.trees, 1024)} trees x {S - 1} simulations"},
    }
    print(json.dumps(line), flush=True)


# ======================================================================================================
# our arm
# ======================================================================================================
class ClockSampler(threading.Thread):
    """One long-running `nvidia-smi -lms 100` for the whole timed phase (the recipe's clocks line)."""
    FIELDS = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
              "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
              "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.proc = index, [], None

    def run(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.FIELDS}",
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            for line in self.proc.stdout:
                parts = [x.strip() for x in line.strip().split(",")]
                if len(parts) >= 7:
                    self.samples.append(parts)
        except Exception:
            pass

    def stop(self):
        if self.proc is not None:
            try:
                self.proc.terminate()
            except Exception:
                pass
        self.join(timeout=3)
        sm = [float(s[0]) for s in self.samples if s[0].replace(".", "").isdigit()]
        reasons = set()
        for s in self.samples:
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), s[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None,
                "sm_max_mhz": float(self.samples[0][1]) if self.samples else None,
                "samples": len(self.samples), "reasons": sorted(reasons)}


def ncu_traffic_per_launch():
    """dram__bytes_read.sum + dram__bytes_write.sum per launch of the fused search-step kernel, from the
    committed `ncu --set full` capture (profiles/); None if the summary is missing."""
    import csv
    path = os.path.join(ROOT, "profiles", "r01c_ncu_full_k_search_step.csv")
    try:
        rows = list(csv.reader(open(path)))
        hdr, units = rows[0], rows[1]
        ir, iw, ik = hdr.index("dram__bytes_read.sum"), hdr.index("dram__bytes_write.sum"), hdr.index("Kernel Name")
        scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
        vals = [float(r[ir]) * scale[units[ir]] + float(r[iw]) * scale[units[iw]] for r in rows[2:] if "1, 1" in r[ik]]
        return sum(vals) / len(vals) if vals else None
    except Exception:
        return None


def run_ours(args):
    import torch
    import torch.distributed as dist
    from hanabizero_b200 import _lib, cytree
    from hanabizero_b200.dist import gather_root_stats_equal
    from hanabizero_b200.hanabi_env import HanabiVecEnv
    from hanabizero_b200.mcts import MCTS, SearchConfig
    from hanabizero_b200.model import MuZeroNetFull

    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py needs a GPU: hanabizero_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    lib = _lib.load()

    N, A, S, F = args.trees, 20, args.sims, F_HIDDEN
    K, W = args.steps, max(args.warmup, 3)
    cfg = SearchConfig(num_simulations=S, amp_type=args.amp)
    torch.manual_seed(0)
    obs_dim = 785 if args.mdp == "global" else 660
    model = MuZeroNetFull(obs_dim * args.stack, A).randomize_heads(seed=0).to(dev).eval()

    # ---- synthetic roots: real Hanabi positions (reset + k random legal steps) -> initial inference ----
    env = HanabiVecEnv(N, "Hanabi-Full", np.arange(N) + rank * N, device=dev)
    g, loc, legal = env.reset_all()
    gen = torch.Generator(device=dev).manual_seed(1234 + rank)
    for _ in range(10):
        acts = torch.multinomial(legal, 1, generator=gen).view(-1).int()
        g, loc, legal, _, _, _ = env.step_all(acts, auto_reset=True)
    env.check()
    obs = (g if args.mdp == "global" else loc).repeat(1, args.stack)
    with torch.no_grad(), torch.autocast("cuda", dtype=torch.float16, enabled=args.amp == "torch_amp"):
        _, root_logits, root_hidden = model.initial_inference_device(obs)
    root_logits = root_logits.float().contiguous()
    root_hidden = root_hidden.contiguous()
    rng = np.random.default_rng(7 + rank)
    noise = torch.from_numpy(rng.dirichlet([0.3] * A, N).astype(np.float32)).to(dev)
    zeros_r = torch.zeros(N, device=dev)
    legal_i = legal.int().contiguous()
    packed = torch.empty(world * N, A + 1, dtype=torch.int32, device=dev) if world > 1 else None

    mcts = MCTS(cfg)
    launches_before, gemm_before = _lib.launch_count(), _lib.gemm_launch_count()

    def search_step(roots_holder):
        roots = cytree.Roots(N, A, S, device=dev)
        roots.prepare(CONST["frac"], noise, zeros_r, root_logits, legal_i)
        mcts.run_multi(roots, model, root_hidden, use_graph=not args.no_graph)
        visits, values = roots.get_stats_tensors()
        if world > 1:
            visits, values = gather_root_stats_equal(visits, values, packed)
        roots_holder[0] = roots
        return visits, values

    holder = [None]
    search_step(holder)                       # eager search (also the launch census)
    torch.cuda.synchronize()
    launches_per_search = _lib.launch_count() - launches_before
    gemm_per_search = _lib.gemm_launch_count() - gemm_before
    for _ in range(W):
        search_step(holder)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
        time.sleep(0.3)   # let nvidia-smi attach before the timed regions
    # ---- timed region 1: device-resident search ------------------------------------------------
    barrier()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(K):
        visits, values = search_step(holder)
    e1.record()
    barrier()
    ms = e0.elapsed_time(e1)
    t_ms = torch.tensor([ms], device=dev)
    if world > 1:
        dist.all_reduce(t_ms, op=dist.ReduceOp.MAX)
    ms_total = float(t_ms.item())
    sims_total = world * N * (S - 1) * K
    value = sims_total / (ms_total * 1e-3)
    assert int(visits[:N].sum().item()) == N * (S - 1), "search did not run the expected simulations"

    # ---- timed region 2: end to end through the list/numpy drop-in API with pinned host buffers ------
    h_noise, h_logits = noise.cpu().pin_memory(), root_logits.cpu().pin_memory()
    h_legal, h_hidden = legal_i.cpu().pin_memory(), root_hidden.cpu().pin_memory()
    h_reward = torch.zeros(N).pin_memory()
    h_visits = torch.empty(N, A, dtype=torch.int32).pin_memory()
    h_values = torch.empty(N).pin_memory()

    def e2e_step():
        roots = cytree.Roots(N, A, S, device=dev)
        roots.prepare(CONST["frac"], h_noise, h_reward, h_logits, h_legal)
        mcts.run_multi(roots, model, h_hidden, use_graph=not args.no_graph)
        v, val = roots.get_stats_tensors()
        h_visits.copy_(v, non_blocking=True)
        h_values.copy_(val, non_blocking=True)
        torch.cuda.current_stream().synchronize()
        holder[0] = roots
        return h_visits

    def timed_e2e(step_fn, finish=None, warm=2):
        for _ in range(warm):
            step_fn()
        if finish:
            finish()
        barrier()
        e0.record()
        for _ in range(K):
            step_fn()
        if finish:
            finish()
        e1.record()
        barrier()
        t = torch.tensor([max(e0.elapsed_time(e1), 0.0)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return sims_total / (float(t.item()) * 1e-3)

    e2e_serial = timed_e2e(e2e_step)
    # the same work through the double-buffered public API: the copies of neighbouring searches overlap the search
    from hanabizero_b200.mcts import SearchPipeline
    pipe = SearchPipeline(mcts, model, N, A, depth=2, device=dev)
    h_out = [(torch.empty(N, A, dtype=torch.int32).pin_memory(), torch.empty(N).pin_memory()) for _ in range(2)]
    turn = [0]

    def piped_step():
        hv, hval = h_out[turn[0] % 2]
        turn[0] += 1
        pipe.submit(CONST["frac"], h_noise, h_reward, h_logits, h_legal, h_hidden, hv, hval)

    e2e_value = timed_e2e(piped_step, pipe.drain, warm=6)   # each slot: one eager search, one capture, one replay
    assert int(h_out[0][0].sum().item()) == N * (S - 1) and torch.equal(h_out[0][0], h_out[1][0]), "pipelined search result"
    assert torch.equal(h_out[0][0], h_visits), "pipelined and serial searches must agree"
    h2d = sum(t.numel() * t.element_size() for t in (h_noise, h_logits, h_legal, h_hidden, h_reward))
    d2h = h_visits.numel() * 4 + h_values.numel() * 4

    # ---- timed region 3: Hanabi env steps (device-resident and e2e) -------------------------------
    T = args.env_steps
    acts_buf = torch.zeros(N, dtype=torch.int32, device=dev)

    def env_pass(steps, host):
        """host: None = device-resident; "f32" / "u8" = scalar-API style (actions come from the host, the
        observation and the legal mask go back to the host every step) with float32 or byte observations."""
        nonlocal legal
        for _ in range(steps):
            if host:
                acts_buf.copy_(h_acts, non_blocking=True)
            else:
                acts_buf.copy_(torch.argmax(legal * torch.rand_like(legal), dim=1))
            if host == "u8":
                env.step_all(acts_buf, auto_reset=True, want_local=False, out_global=g8, out_legal=legal8)
                h_obs8.copy_(g8_store, non_blocking=True)
                h_leg8.copy_(legal8, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                h_acts.copy_(torch.from_numpy(np.argmax(h_leg8.numpy() * host_rand, axis=1).astype(np.int32)))
                continue
            gg, _, legal, r, d, s = env.step_all(acts_buf, auto_reset=True, want_local=False)
            if host:
                h_obs.copy_(gg, non_blocking=True)
                h_leg.copy_(legal, non_blocking=True)
                torch.cuda.current_stream().synchronize()
                h_acts.copy_(torch.from_numpy(np.argmax(h_leg.numpy() * host_rand, axis=1).astype(np.int32)))

    gpad = (env.global_dim + 15) // 16 * 16
    g8_store = torch.zeros(N, gpad, dtype=torch.uint8, device=dev)     # 16-byte-multiple rows: word stores
    g8, legal8 = g8_store[:, :env.global_dim], torch.zeros(N, A, dtype=torch.uint8, device=dev)
    h_obs = torch.empty(N, env.global_dim).pin_memory()
    h_leg = torch.empty(N, A).pin_memory()
    h_obs8 = torch.empty(N, gpad, dtype=torch.uint8).pin_memory()
    h_leg8 = torch.empty(N, A, dtype=torch.uint8).pin_memory()
    h_acts = torch.zeros(N, dtype=torch.int32).pin_memory()
    host_rand = rng.random((N, A)).astype(np.float32) + 0.01

    def host_pick():
        h_leg.copy_(legal); torch.cuda.synchronize()
        h_acts.copy_(torch.from_numpy(np.argmax(h_leg.numpy() * host_rand, axis=1).astype(np.int32)))

    def timed_env(steps, host, warm):
        if host:
            host_pick()
        env_pass(warm, host)
        barrier()
        e0.record()
        env_pass(steps, host)
        e1.record()
        barrier()
        if host == "u8":       # the float legal mask of the device path is stale after byte steps
            env.observe()
        t = torch.tensor([e0.elapsed_time(e1)], device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return world * N * steps / (float(t.item()) * 1e-3)

    T_host = max(T // 4, 10)
    # device-resident: ten steps (action pick + fused step/auto-reset/observe launch) per CUDA graph, replayed — the
    # Python loop around five tiny launches per step would otherwise be what is timed
    env_pass(20, None)
    torch.cuda.synchronize()
    per_graph = 10
    env_graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(env_graph):
        env_pass(per_graph, None)
    legal = env.legal
    reps = max(T // per_graph, 1)
    env_graph.replay()
    barrier()
    e0.record()
    for _ in range(reps):
        env_graph.replay()
    e1.record()
    barrier()
    t_env = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(t_env, op=dist.ReduceOp.MAX)
    env_value = world * N * reps * per_graph / (float(t_env.item()) * 1e-3)
    T = reps * per_graph
    legal = env.legal
    env_e2e_f32 = timed_env(T_host, "f32", 5)
    legal = env.legal
    env_e2e = timed_env(T_host, "u8", 5)
    legal = env.legal
    env.check()
    # ---- timed region 4 (informational): whole self-play moves, device-resident ----
    from hanabizero_b200.selfplay import SelfPlayEngine
    eng = SelfPlayEngine(N, "Hanabi-Full", model, cfg, seeds=np.arange(N) + 7 * N * (rank + 1), mdp=args.mdp,
                         stack=args.stack, device=dev)
    eng.reset()
    for _ in range(3):
        eng.step()
    barrier()
    e0.record()
    n_moves = 5
    for _ in range(n_moves):
        eng.step()
    e1.record()
    barrier()
    sp_ms = torch.tensor([e0.elapsed_time(e1)], device=dev)
    if world > 1:
        dist.all_reduce(sp_ms, op=dist.ReduceOp.MAX)
    eng.env.check()
    selfplay_moves = world * N * n_moves / (float(sp_ms.item()) * 1e-3)
    clocks = sampler.stop() if rank == 0 else None

    # ---- roofline of the dominant tree kernel (fused backprop+traverse+gather), timed live ----------
    roof = None
    env_roof = None
    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback 6650 GB/s (B200_PROFILING.md)"
        # the production launch: hz_trees_search_step (decode + expand + backprop + min/max + traverse +
        # hand-off) on synthetic network outputs, alone on the stream, L2 flushed before every launch
        plan = model.recurrent_plan(torch.float16 if args.amp == "torch_amp" else torch.float32)
        ch = plan.chain(N)
        eb = ch.x0.element_size()
        ch.out.copy_(torch.randn_like(ch.out.float()).to(ch.out.dtype))
        ch.state.copy_(torch.rand_like(ch.state.float()).to(ch.state.dtype))
        roots = cytree.Roots(N, A, S, device=dev)
        roots.prepare(CONST["frac"], noise, zeros_r, root_logits, legal_i)
        mm = cytree.MinMaxStatsList(N); mm.set_delta(CONST["delta"])
        pool = torch.rand(S, N, F, device=dev).to(ch.x0.dtype)
        io = _lib.SearchIO()
        io.value_logits, io.ld_value = ch.value_logits.data_ptr(), ch.value_logits.stride(0)
        io.reward_logits, io.ld_reward = ch.reward_logits.data_ptr(), ch.reward_logits.stride(0)
        io.policy_logits, io.ld_policy = ch.policy_logits.data_ptr(), ch.policy_logits.stride(0)
        io.next_state, io.ld_state = ch.state.data_ptr(), ch.state.stride(0)
        io.support, io.support_width, io.support_delta = plan.support.data_ptr(), plan.n_support, plan.net.support_delta
        io.elem_bytes, io.sanitize_nan = eb, 1
        io.pool, io.state_cols = pool.data_ptr(), F
        io.out_batch, io.ld_batch, io.onehot_cols = ch.x0.data_ptr(), ch.x0.stride(0), plan.OH
        io.out_ix, io.out_action = None, None
        io.minmax, io.value_delta_max = mm.tensor(dev).data_ptr(), CONST["delta"]
        io.discount, io.pb_c_base, io.pb_c_init = CONST["discount"], CONST["pb_c_base"], CONST["pb_c_init"]
        st = torch.cuda.current_stream().cuda_stream
        ref = _lib.C.byref(io)
        _lib.check(lib.hz_trees_search_step(roots.handle, st, 0, 1, ref))
        evs = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(S)]
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
        depth_sum, depth_max = 0.0, []
        gen_out = torch.Generator(device=dev).manual_seed(99)
        outs = [torch.randn(ch.out.shape, device=dev, generator=gen_out).to(ch.out.dtype) for _ in range(8)]
        for x in range(1, S - 1):
            ch.out.copy_(outs[x % 8])   # fresh synthetic network outputs per simulation (realistic tree depths)
            flush.fill_(x & 1)  # evict L2: every launch starts cold, like inside a search whose working set exceeds L2
            evs[x][0].record()
            _lib.check(lib.hz_trees_search_step(roots.handle, st, x, 1, ref))
            evs[x][1].record()
            pl = roots.export(1)["path_len"].float()
            depth_sum += float(pl.mean().item()) - 1.0
            depth_max.append(int(pl.max().item()) - 1)
        torch.cuda.synchronize()
        durs = [evs[x][0].elapsed_time(evs[x][1]) * 1e-3 for x in range(1, S - 1)]
        # the same launches back to back inside a CUDA graph, no flush (what the search loop sees: 270 MB of
        # tree + pool per search is larger than L2, but the hot nodes of a tree stay L2-resident between sims)
        roots.prepare(CONST["frac"], noise, zeros_r, root_logits, legal_i)
        mm.clear()
        _lib.check(lib.hz_trees_search_step(roots.handle, st, 0, 1, ref))
        torch.cuda.synchronize()
        gr = torch.cuda.CUDAGraph()
        with torch.cuda.graph(gr):
            st_c = torch.cuda.current_stream().cuda_stream
            for x in range(1, S - 1):
                ch.out.copy_(outs[x % 8])
                _lib.check(lib.hz_trees_search_step(roots.handle, st_c, x, 1, ref))
        w0, w1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        w0.record()
        gr.replay()
        w1.record()
        torch.cuda.synchronize()
        _lib.check(lib.hz_trees_set_progress(roots.handle, S - 2))
        gc = torch.cuda.CUDAGraph()      # the same graph without the tree launches: the copies' own cost
        with torch.cuda.graph(gc):
            for x in range(1, S - 1):
                ch.out.copy_(outs[x % 8])
        c0, c1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        c0.record()
        gc.replay()
        c1.record()
        torch.cuda.synchronize()
        warm_us = (w0.elapsed_time(w1) - c0.elapsed_time(c1)) * 1e3 / (S - 2)
        n_l = len(durs)
        D = depth_sum / n_l
        s_mean = statistics.mean(range(1, S - 1))
        # SURVEY §8d per-simulation bytes (tree part with the hidden row in the model dtype) + what this fused
        # launch additionally replaces: reading both support-logit rows and copying the new state into the pool
        per_tree = b_sim(A, D, s_mean, F * eb / 4.0) + 2 * plan.n_support * eb + 2 * F * eb
        bytes_launch = N * per_tree
        achieved = bytes_launch / statistics.mean(durs) / 1e9
        roof = {"bound": "hbm", "kernel": "k_search_step<half,backprop,traverse> (hz_trees_search_step)",
                "achieved": achieved, "peak": peak, "unit": "GB/s", "frac": achieved / peak,
                "traffic": ncu_traffic_per_launch(),
                "peak_source": peak_src, "launch_us": 1e6 * statistics.mean(durs), "launch_us_in_graph_no_flush": warm_us,
                "algorithmic_bytes_per_launch": bytes_launch, "algorithmic_bytes_per_tree": per_tree,
                "mean_depth": D, "l2": "flushed before every timed launch (256 MiB write)",
                "per_sim_us_flushed": [round(1e6 * d, 1) for d in durs[::6]], "max_depth": depth_max[::6]}
        # env kernel
        evs2 = []
        for _ in range(30):
            flush.fill_(1)
            a0, a1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            acts_buf.copy_(torch.argmax(legal * torch.rand_like(legal), dim=1))
            a0.record()
            _, _, legal, _, _, _ = env.step_all(acts_buf, auto_reset=True, want_local=False)
            a1.record()
            evs2.append((a0, a1))
        torch.cuda.synchronize()
        d2 = statistics.mean(a.elapsed_time(b) for a, b in evs2) * 1e-3
        env_roof = {"bound": "hbm", "kernel": "k_env<step,observe>", "achieved": N * B_ENV_STEP / d2 / 1e9, "peak": peak,
                    "unit": "GB/s", "frac": N * B_ENV_STEP / d2 / 1e9 / peak, "traffic": None, "launch_us": d2 * 1e6}

    # ---- CPU baseline on the box's host cores (rank 0, N=1 only) ----------------------------------------
    cpu = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count() or 1
        one = cpu_reference(min(N, 1024), A, S, 1, reps=4, env_steps_per_core=100000)
        allc = cpu_reference(N, A, S, cores, reps=6, env_steps_per_core=100000)
        cpu = {"value": allc["sims_per_s"], "unit": "simulations/s", "cores": allc["cores"], "kind": allc["kind"],
               "value_1core": one["sims_per_s"],
               "env_steps_per_s": allc["env_steps_per_s"], "env_steps_per_s_1core": one["env_steps_per_s"],
               "env_kind": allc["env_kind"],
               "sample": (f"{N} Hanabi-Full trees x {S - 1} simulations x 6 over {allc['cores']} processes (and "
                          f"{min(N, 1024)} trees x 4 on 1 core): reference cytree driven like core/mcts.py with pre-generated "
                          "network outputs, no model time; env: 100000 steps/core of reference libhanabi in C++ "
                          "(about 20 core-seconds in total)")}

    if rank == 0:
        line = {
            "metric": "mcts_simulations_per_sec", "value": value, "unit": "simulations/s", "n_gpus": world,
            "steps": K, "warmup": W, "ms_per_step": ms_total / K, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"Hanabi-Full 2p {args.mdp} {'MDP' if args.mdp == 'global' else 'POMDP'}, {N} trees/GPU x {S} simulations ({S - 1} executed, "
                                   "as core/mcts.py:25-26), MuZeroNetFull random-init with re-drawn heads",
                       "trees_per_gpu": N, "trees_total": world * N, "actions": A, "simulations": S, "stack": args.stack,
                       "model_amp": args.amp, "cuda_graph": not args.no_graph, "sharding": f"roots x{world}",
                       "l2": "working set (tree nodes 67 MB + hidden pool >200 MB per search) exceeds the 126 MB L2"},
            "e2e": {"value": e2e_value, "unit": "simulations/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h,
                    "api": "SearchPipeline.submit/wait (hanabizero_b200/mcts.py): pinned host inputs in, root statistics out, "
                           "every search; two searches in flight so the copies overlap the neighbouring search",
                    "serial": {"value": e2e_serial, "api": "Roots.prepare + MCTS.run_multi + get_stats on host tensors, one "
                                                           "search at a time, a stream synchronise per search"}},
            "gpu_launches": int(launches_per_search * K),
            "gpu_launches_per_search": int(launches_per_search),
            "library_gemm_launches_per_search": int(gemm_per_search),
            "clocks": clocks, "roofline": roof, "cpu_baseline": cpu,
            "env": {"metric": "hanabi_env_steps_per_sec", "value": env_value, "unit": "steps/s",
                    "e2e": {"value": env_e2e, "unit": "steps/s", "h2d_bytes_per_step": 4 * N,
                            "d2h_bytes_per_step": N * (gpad + A), "obs_dtype": "u8",
                            "what": "actions from pinned host memory in, global observation + legal mask out as 0/1 "
                                    "bytes (the encoder's own value type, hz_envs_step_observe_u8), host picks the next "
                                    "action; one sync per step",
                            "f32": {"value": env_e2e_f32, "d2h_bytes_per_step": 4 * N * (env.global_dim + A)}},
                    "games_per_gpu": N, "steps_timed": T, "includes": "on-device random legal action pick (3 torch kernels) + "
                    "fused step/auto-reset/observe kernel, ten steps per CUDA graph", "roofline": env_roof},
            "selfplay": {"metric": "selfplay_moves_per_sec", "value": selfplay_moves, "unit": "moves/s",
                         "what": "frame stack -> representation+prediction -> Roots.prepare(Dirichlet) -> run_multi -> "
                                 "select_action -> env step with auto-reset, all on the device (SelfPlayEngine.step)",
                         "simulations_per_sec": selfplay_moves * (S - 1)},
        }
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()


if __name__ == "__main__":
    a = parse_args()
    if a.impl == "reference":
        run_reference(a)
    else:
        run_ours(a)
