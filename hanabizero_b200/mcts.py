"""Search driver: drop-in for the reference's core/mcts.py (MCTS.run_multi, :7-57; `search` is
the upstream-EfficientZero alias BASELINE.json uses).

The reference loop crosses the host twice per simulation (Python-list marshalling into C++, a
Python gather of hidden states, H2D of the batch, D2H + .tolist() of the network output).  Here a
simulation is, on the device and on one stream:

    [backprop of sim k-1 + traverse of sim k + gather of the parents' hidden states]  (1 launch)
    -> model.recurrent_inference_device (PyTorch, cuBLAS)  -> next hidden state into pool[k]

and the whole `num_simulations - 1` loop is captured in a CUDA graph after the first search with
a given (tree handle, model, shape), so a search is one graph launch.  Exactly like the reference
the last iteration is skipped (core/mcts.py:25-26): `num_simulations - 1` simulations run.

Models: anything exposing `recurrent_inference_device(hidden, action) -> (value[N], reward[N],
policy_logits[N, A], next_hidden[N, F])` on CUDA (hanabizero_b200.model does) gets the
device-resident path.  A model with only the reference's `recurrent_inference` (numpy or tensor
NetworkOutput) still works through the same kernels, paying that model's own host hops.
"""
import contextlib
import time
import weakref

import numpy as np
import torch

from . import _lib, cytree
from ._lib import check, ptr


class SearchConfig:
    """The hot-path constants of BaseMuZeroConfig (core/config.py:105-111) and the Hanabi configs
    (config/hanabi_control/__init__.py:25-28,143-146) for callers without a reference config."""

    def __init__(self, num_simulations=50, discount=0.999, value_delta_max=0.006, pb_c_base=19652,
                 pb_c_init=1.25, amp_type="none", root_dirichlet_alpha=0.3,
                 root_exploration_fraction=0.25):
        self.num_simulations = num_simulations
        self.discount = discount
        self.value_delta_max = value_delta_max
        self.pb_c_base = pb_c_base
        self.pb_c_init = pb_c_init
        self.amp_type = amp_type
        self.root_dirichlet_alpha = root_dirichlet_alpha
        self.root_exploration_fraction = root_exploration_fraction


class _Workspace:
    """Static device buffers of one (tree handle, model, shape): hidden-state pool [S, N, F], the
    gathered batch handed to the model, the min/max statistics, and the captured graph."""

    def __init__(self, roots, model, sims, feature, dtype):
        n, dev = roots.root_num, roots.device
        self.model_ref = weakref.ref(model)   # id(model) can be recycled: a hit must be for the same live object
        self.chain = None                     # the BoundChain a captured graph points into (kept alive with it)
        self.pool = torch.empty(sims, n, feature, dtype=dtype, device=dev)
        self.hidden = torch.empty(n, feature, dtype=dtype, device=dev)
        self.action64 = torch.zeros(n, 1, dtype=torch.int64, device=dev)
        self.ix = torch.empty(n, dtype=torch.int32, device=dev)
        self.la = torch.empty(n, dtype=torch.int32, device=dev)
        self.minmax = cytree.MinMaxStatsList(n)
        self.minmax.tensor(dev)
        self.graph = None
        self.searches = 0


class MCTS(object):
    max_cached_workspaces = 8

    def __init__(self, config, use_plan=True):
        self.config = config
        self.use_plan = use_plan
        self._ws = {}
        cytree.on_handle_destroyed(self._forget_handle)

    def _forget_handle(self, handle_value):
        """A cached tree batch was destroyed: graphs captured over its buffers must never be replayed."""
        for key in [k for k in self._ws if k[0] == handle_value]:
            del self._ws[key]

    # ---------------------------------------------------------------------------------------------
    def _autocast(self):
        if getattr(self.config, "amp_type", "none") == "torch_amp":  # core/mcts.py:38-40
            return torch.autocast("cuda", dtype=torch.float16)
        return contextlib.nullcontext()

    def _plan(self, model):
        """The model's fused eval-mode plan (hanabizero_b200/plan.py) if it offers one and use_plan is on."""
        if not self.use_plan or not hasattr(model, "recurrent_plan"):
            return None
        amp = getattr(self.config, "amp_type", "none") == "torch_amp"
        return model.recurrent_plan(torch.float16 if amp else torch.float32)

    def _workspace(self, roots, model, hidden_state_roots, gemm_sm_target=0, stage_limit=0, executor="library"):
        sims = int(self.config.num_simulations)
        key = (roots.handle.value, id(model), sims, getattr(self.config, "amp_type", "none"), self.use_plan,
               int(gemm_sm_target), int(stage_limit), executor)
        ws = self._ws.get(key)
        if ws is not None and ws.model_ref() is not model:
            ws = None   # a different model object that happens to live at a recycled address
        if ws is None:
            n, dev = roots.root_num, roots.device
            feature = int(hidden_state_roots.shape[-1])
            if self._plan(model) is not None:
                dtype = self._plan(model).dtype
            elif hasattr(model, "recurrent_inference_device"):
                with torch.no_grad(), self._autocast():
                    probe = model.recurrent_inference_device(
                        torch.zeros(2, feature, device=dev), torch.zeros(2, 1, dtype=torch.int64, device=dev))
                dtype = probe[3].dtype
            else:
                dtype = torch.float32
            if len(self._ws) >= self.max_cached_workspaces:
                self._ws.pop(next(iter(self._ws)))
            ws = self._ws[key] = _Workspace(roots, model, sims, feature, dtype)
            ws.gemm_sm_target, ws.stage_limit, ws.executor = int(gemm_sm_target), int(stage_limit), executor
        return ws

    def _simulate(self, roots, model, ws):
        """The device-resident loop; everything is enqueued on the current stream (capturable)."""
        cfg, lib, h = self.config, roots._lib, roots.handle
        sims = int(cfg.num_simulations)
        st = torch.cuda.current_stream(roots.device).cuda_stream
        mm = ws.minmax
        mm.set_delta(cfg.value_delta_max)
        mm.clear()
        mmp = ptr(mm.tensor(roots.device))
        row_bytes = ws.hidden.shape[1] * ws.hidden.element_size()
        base, init, disc, delta = int(cfg.pb_c_base), float(cfg.pb_c_init), float(cfg.discount), float(cfg.value_delta_max)
        if sims < 2:
            return
        check(lib.hz_trees_traverse(h, st, base, init, disc, mmp, delta, ptr(ws.ix), None, ptr(ws.la),
                                    ptr(ws.action64), ptr(ws.pool), ptr(ws.hidden), row_bytes))
        for x in range(1, sims):
            with self._autocast():
                value, reward, logits, state = model.recurrent_inference_device(ws.hidden, ws.action64)
            ws.pool[x].copy_(state)
            value = value.float().contiguous()
            reward = reward.float().contiguous()
            logits = logits.float().contiguous()
            if x < sims - 1:
                check(lib.hz_trees_backprop_traverse(
                    h, st, x, disc, ptr(reward), ptr(value), ptr(logits), 1, mmp, delta, base, init,
                    ptr(ws.ix), None, ptr(ws.la), ptr(ws.action64), ptr(ws.pool), ptr(ws.hidden), row_bytes))
            else:
                check(lib.hz_trees_backprop(h, st, x, disc, ptr(reward), ptr(value), ptr(logits), 1, mmp))

    def _simulate_chain(self, roots, plan, ws):
        """Device-resident loop on the fused plan: per simulation one GEMM chain (7 or 5 cuBLASLt
        launches) and ONE tree launch that decodes the raw value/reward logits, expands, back-propagates,
        stores the new hidden state in the pool, traverses and writes the next [hidden ‖ one-hot] batch."""
        cfg, lib, h = self.config, roots._lib, roots.handle
        sims = int(cfg.num_simulations)
        st = torch.cuda.current_stream(roots.device).cuda_stream
        mm = ws.minmax
        mm.set_delta(cfg.value_delta_max)
        mm.clear()
        if sims < 2:
            return
        if ws.chain is None or ws.chain.plan is not plan:
            # the workspace owns its chain (activation buffers + cuBLASLt plan): searches of different tree batches may
            # run on different streams at the same time (SearchPipeline), and a captured graph keeps its buffers alive
            from .plan import BoundChain
            ws.chain = BoundChain(plan, roots.root_num)
            if ws.gemm_sm_target:
                ws.chain.set_sm_target(ws.gemm_sm_target)
            if ws.executor == "rows" and ws.chain.rows_supported():
                ws.chain.set_executor("rows")
        ch = ws.chain
        io = _lib.SearchIO()
        io.value_logits, io.ld_value = ptr(ch.value_logits), ch.value_logits.stride(0)
        io.reward_logits, io.ld_reward = ptr(ch.reward_logits), ch.reward_logits.stride(0)
        io.policy_logits, io.ld_policy = ptr(ch.policy_logits), ch.policy_logits.stride(0)
        io.next_state, io.ld_state = None, 0      # the dynamics GEMM writes pool[x] itself (BoundChain.bind_state)
        io.support, io.support_width, io.support_delta = ptr(plan.support), plan.n_support, plan.net.support_delta
        io.elem_bytes, io.sanitize_nan = ch.x0.element_size(), 1
        io.pool, io.state_cols = ptr(ws.pool), plan.F
        io.out_batch, io.ld_batch, io.onehot_cols = ptr(ch.x0), ch.x0.stride(0), plan.OH
        io.out_ix, io.out_action = ptr(ws.ix), ptr(ws.la)
        io.minmax, io.value_delta_max = ptr(mm.tensor(roots.device)), float(cfg.value_delta_max)
        io.discount, io.pb_c_base, io.pb_c_init = float(cfg.discount), int(cfg.pb_c_base), float(cfg.pb_c_init)
        io.stage_limit = ws.stage_limit
        ref = _lib.C.byref(io)
        check(lib.hz_trees_search_step(h, st, 0, 1, ref))
        for x in range(1, sims):
            ch.bind_state(ws.pool[x])
            ch.run(st)
            check(lib.hz_trees_search_step(h, st, x, 1 if x < sims - 1 else 0, ref))

    def _simulate_compat(self, roots, model, ws):
        """Same kernels around a model that only speaks the reference's NetworkOutput contract."""
        cfg = self.config
        sims = int(cfg.num_simulations)
        mm = ws.minmax
        mm.set_delta(cfg.value_delta_max)
        mm.clear()
        for x in range(1, sims):
            results = cytree.ResultsWrapper(roots.root_num)
            cytree.multi_traverse(roots, cfg.pb_c_base, cfg.pb_c_init, cfg.discount, mm, results,
                                  as_tensor=True, pool=ws.pool, out_hidden=ws.hidden, out_action64=ws.action64)
            with self._autocast():
                out = model.recurrent_inference(ws.hidden.float(), ws.action64)
            dev = roots.device
            ws.pool[x].copy_(cytree.as_device(out.hidden_state, ws.pool.dtype, dev))
            cytree.multi_back_propagate(x, cfg.discount, _flat(out.reward), _flat(out.value),
                                        out.policy_logits, mm, results, sanitize_nan=True)

    # ---------------------------------------------------------------------------------------------
    def run_multi(self, roots, model, hidden_state_roots, use_graph=True, gemm_sm_target=0, stage_limit=0,
                  executor="library"):
        """core/mcts.py:11-57.  roots: cytree.Roots already prepared; hidden_state_roots: [N, F]
        numpy array or tensor (any device).  Mutates `roots` in place and returns None.
        gemm_sm_target > 0 sizes the network's library GEMMs for that many SMs instead of the whole device (for callers
        that keep several searches in flight on different streams: SearchPipeline sets it); stage_limit > 0 likewise
        shrinks the tree step's shared-memory staging so that it shares SMs with other searches' GEMMs
        (hz_search_io.stage_limit).  executor="rows" runs the network as one row-block resident launch per simulation
        (BoundChain.set_executor; fp16 Hanabi-Full plan, otherwise ignored).  None of them changes any result of the
        tree step; the SM target and the executor change network roundings in the last bits."""
        with torch.no_grad():
            if getattr(model, "training", True):
                model.eval()          # walks every submodule: only when something is still in training mode
            ws = self._workspace(roots, model, hidden_state_roots, gemm_sm_target, stage_limit, executor)
            if self._plan(model) is not None:
                self._plan(model).refresh()   # re-fold weights in place if the module was updated
            ws.pool[0].copy_(cytree.as_device(hidden_state_roots, ws.pool.dtype, roots.device))
            sims = int(self.config.num_simulations)
            plan = self._plan(model)
            if plan is not None:
                simulate = lambda: self._simulate_chain(roots, plan, ws)
            else:
                simulate = lambda: self._simulate(roots, model, ws)
            if plan is None and not hasattr(model, "recurrent_inference_device"):
                self._simulate_compat(roots, model, ws)
            elif not use_graph:
                simulate()
            elif ws.graph is None and ws.searches == 0:
                simulate()  # first search runs eagerly (also warms cuBLAS)
            else:
                if ws.graph is None:
                    torch.cuda.synchronize(roots.device)
                    graph = torch.cuda.CUDAGraph()
                    with torch.cuda.graph(graph):
                        simulate()
                    ws.graph = graph
                ws.graph.replay()
                # host-side progress of the handle is not touched by a replay
                check(roots._lib.hz_trees_set_progress(roots.handle, max(sims - 1, 0)))
            ws.searches += 1

    search = run_multi  # upstream EfficientZero name (BASELINE.json north_star)


def _flat(x):
    return x.reshape(-1) if isinstance(x, torch.Tensor) else np.asarray(x).reshape(-1)


STAGE_LIMIT_IN_FLIGHT = 4   # hz_search_io.stage_limit used when several searches share the GPU (profiles/r02_sm_target.md)


def gemm_sm_target_for(num_roots, in_flight, device):
    """SMs each search's library GEMMs should be sized for when `in_flight` independent searches of `num_roots` trees
    share a GPU.  cuBLASLt sizes a GEMM to fill the whole device, so the GEMMs of different streams queue behind one
    another; with small root batches it pays to size each for a share of the SMs so that they run side by side.
    Measured on B200 (profiles/r02_sm_target.md): 3n/64 SMs up to 56 (24 at 512 trees, 48 at 1024, 56 at 1536-2048);
    from ~3000 trees on the whole-device kernels are the best choice (returns 0)."""
    if in_flight <= 1 or num_roots >= 3072:
        return 0
    sms = torch.cuda.get_device_properties(device).multi_processor_count
    return int(max(16, min(3 * num_roots // 64, 56 * sms // 148)))


class SearchPipeline:
    """Searches kept in flight: `depth` slots, each with its own tree batch (and therefore its own captured search
    graph, network buffers and staging buffers) and its own compute stream.  Inputs are copied in on a copy stream and
    root statistics copied out on another, so for host-fed searches the copies of search i+1 / i-1 overlap search i —
    and, because a search is a dependent chain of short launches that leaves the GPU idle at every kernel boundary, the
    searches of different slots overlap each other: independent root batches (the reference's actors, each with its
    own p_mcts_num roots, core/selfplay_worker.py:93,102) fill one another's bubbles.

        pipe = SearchPipeline(MCTS(cfg), model, num_roots, num_actions)
        t = pipe.submit(fraction, noises, rewards, logits, legal, hidden_roots, out_visits, out_values)   # returns at once
        ...
        pipe.wait(t)          # out_visits / out_values now hold the result of that search

    Inputs are what Roots.prepare + MCTS.run_multi take (core/selfplay_worker.py:276-283): pinned host tensors
    (pageable memory works but serialises the copies) or device tensors; outputs likewise.  Device tensors that already
    have the staging buffers' type (float32 / int32 legal mask, contiguous, on this device) are read in place — keep
    them unchanged until wait(ticket).  `noises=None` selects prepare_no_noise.  `gather` (dist.AsyncStatsGather with
    depth >= this depth) additionally all-gathers every search's statistics over the ranks, off the compute streams;
    `gathered(t)` returns them.  Weights must not change while searches are in flight: call drain() before updating
    the module.

    With depth > 1 the network runs on the row-block resident executor (`executor`, hz_rowchain: one launch per
    simulation on n/128 SMs; profiles/r02_rowchain.md) where the plan supports it.  For the library chain two settings
    make the slots share the GPU instead of queueing (profiles/r02_sm_target.md): `gemm_sm_target`
    (default gemm_sm_target_for(num_roots, depth)) sizes each search's library GEMMs for a share of the SMs, and
    `stage_limit` (default 4) keeps the tree step's shared-memory footprint small enough to sit next to a GEMM CTA.
    Results per search are those of MCTS.run_multi with the same two settings (the SM target changes the rounding of
    the network outputs in the last bits, the staging limit changes nothing)."""

    def __init__(self, mcts, model, num_roots, num_actions, depth=8, device=None, gather=None, gemm_sm_target=None,
                 stage_limit=None, executor="auto"):
        self.mcts, self.model = mcts, model
        self.depth = int(depth)
        # "rows": the network as one row-block resident launch per simulation (BoundChain.set_executor) — a quarter of
        # the SMs per search at 4096 trees, which is what searches in flight want; "library": seven cuBLASLt launches,
        # the lower latency for a search that runs alone.  "auto" = rows whenever searches overlap (plans the
        # executor does not cover fall back to the library chain inside MCTS).
        self.executor = ("rows" if self.depth > 1 else "library") if executor == "auto" else executor
        self.n, self.a, self.depth = int(num_roots), int(num_actions), int(depth)
        self.device = next(model.parameters()).device if device is None else torch.device(device)
        dev, sims = self.device, int(mcts.config.num_simulations)
        if gemm_sm_target is None:
            gemm_sm_target = gemm_sm_target_for(self.n, self.depth, dev)
        self.gemm_sm_target = int(gemm_sm_target)
        # searches in flight: a small staging region lets the tree step share SMs with the other slots' GEMM CTAs
        self.stage_limit = (STAGE_LIMIT_IN_FLIGHT if self.depth > 1 else 0) if stage_limit is None else int(stage_limit)
        mcts.max_cached_workspaces = max(mcts.max_cached_workspaces, self.depth + 4)   # one workspace (graph) per slot
        if gather is not None and gather.depth < self.depth:
            raise ValueError("gather.depth must be at least the pipeline depth")
        self.gather = gather
        self.copy_in, self.copy_out = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
        self.slots = []
        for _ in range(self.depth):
            self.slots.append(dict(
                roots=cytree.Roots(self.n, self.a, sims, device=dev), stream=torch.cuda.Stream(dev),
                noise=torch.empty(self.n, self.a, device=dev), reward=torch.empty(self.n, device=dev),
                logits=torch.empty(self.n, self.a, device=dev), legal=torch.empty(self.n, self.a, dtype=torch.int32, device=dev),
                hidden=None, stats=None, visits=None, values=None, ticket=None,
                ev_in=torch.cuda.Event(), ev_done=torch.cuda.Event(), ev_out=torch.cuda.Event(), busy=False))
        for s in self.slots:
            # one int32 buffer [n * A visit counts | n value bit patterns]: the statistics kernel writes both parts in
            # place and the all-gather ships the buffer as it is (no packing kernels)
            s["stats"] = torch.empty(self.n * (self.a + 1), dtype=torch.int32, device=dev)
            s["visits"] = s["stats"][:self.n * self.a].view(self.n, self.a)
            s["values"] = s["stats"][self.n * self.a:].view(torch.float32)
        self._next = 0
        self.host_seconds, self.submitted = 0.0, 0

    def submit(self, fraction, noises, rewards, logits, legal, hidden_roots, out_visits=None, out_values=None):
        i = self._next
        self._next = (i + 1) % self.depth
        s = self.slots[i]
        if s["busy"]:
            s["ev_out"].synchronize()          # the slot's previous search has been read out
        s["busy"] = True
        t_host = time.perf_counter()
        hidden_roots = torch.as_tensor(hidden_roots)
        if s["hidden"] is None or s["hidden"].shape != hidden_roots.shape or s["hidden"].dtype != hidden_roots.dtype:
            s["hidden"] = torch.empty(hidden_roots.shape, dtype=hidden_roots.dtype, device=self.device)
        caller = torch.cuda.current_stream(self.device)
        compute = s["stream"]
        if self.gather is not None and s["ticket"] is not None and self.gather.stream is not None:
            # the slot's statistics buffer is the send buffer of its previous all-gather: the new search may only
            # overwrite it once that collective is done
            compute.wait_event(self.gather.slots[s["ticket"]]["done"])
        inputs = dict(noise=noises, reward=rewards, logits=logits, legal=legal, hidden=hidden_roots)
        if all(v is None or self._usable_in_place(v, s[k]) for k, v in inputs.items()):
            # device-resident inputs of the staging buffers' own type: read in place, no copies (the caller keeps them
            # unchanged until wait(ticket)); the slot's stream only has to follow whatever produced them
            compute.wait_stream(caller)
            src = inputs
        else:
            self.copy_in.wait_stream(caller)
            with torch.cuda.stream(self.copy_in):
                # the slot's staging buffers were last read by its previous search, which ev_out (waited above) follows
                if noises is not None:
                    s["noise"].copy_(torch.as_tensor(noises), non_blocking=True)
                s["reward"].copy_(torch.as_tensor(rewards), non_blocking=True)
                s["logits"].copy_(torch.as_tensor(logits), non_blocking=True)
                s["legal"].copy_(torch.as_tensor(legal), non_blocking=True)
                s["hidden"].copy_(hidden_roots, non_blocking=True)
                s["ev_in"].record(self.copy_in)
            compute.wait_event(s["ev_in"])
            src = s
        with torch.cuda.stream(compute):
            if noises is not None:
                s["roots"].prepare(fraction, src["noise"], src["reward"], src["logits"], src["legal"])
            else:
                s["roots"].prepare_no_noise(src["reward"], src["logits"], src["legal"])
            self.mcts.run_multi(s["roots"], self.model, src["hidden"], gemm_sm_target=self.gemm_sm_target,
                                stage_limit=self.stage_limit, executor=self.executor)
            check(s["roots"]._lib.hz_trees_root_stats(s["roots"].handle, compute.cuda_stream, ptr(s["visits"]),
                                                      ptr(s["values"])))
            s["ev_done"].record(compute)
            if self.gather is not None:
                s["ticket"] = self.gather.submit_flat(s["stats"])
        self.copy_out.wait_event(s["ev_done"])
        with torch.cuda.stream(self.copy_out):
            if out_visits is not None:
                out_visits.copy_(s["visits"], non_blocking=True)
            if out_values is not None:
                out_values.copy_(s["values"], non_blocking=True)
            s["ev_out"].record(self.copy_out)
        self.host_seconds += time.perf_counter() - t_host   # host time spent enqueueing (not waiting for the GPU)
        self.submitted += 1
        return i

    def _usable_in_place(self, x, like):
        return (isinstance(x, torch.Tensor) and x.device == like.device and x.dtype == like.dtype and x.shape == like.shape
                and x.is_contiguous())

    def wait(self, ticket):
        self.slots[ticket]["ev_out"].synchronize()

    def stats(self, ticket):
        """Device tensors (visits int32 [n, A], values float32 [n]) of a ticket's search; valid until the slot is
        submitted again.  The current stream waits for the search."""
        s = self.slots[ticket]
        torch.cuda.current_stream(self.device).wait_event(s["ev_done"])
        return s["visits"], s["values"]

    def gathered(self, ticket):
        """All ranks' statistics of a ticket's search (needs `gather`); the current stream waits for the collective."""
        return self.gather.result(self.slots[ticket]["ticket"])

    def drain(self):
        for s in self.slots:
            if s["busy"]:
                s["ev_out"].synchronize()
                s["stream"].synchronize()
                s["busy"] = False
        if self.gather is not None and self.gather.stream is not None:
            self.gather.stream.synchronize()
