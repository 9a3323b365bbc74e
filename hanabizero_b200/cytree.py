"""Drop-in for the reference's Cython module `core.ctree.cytree`
(/root/reference/core/ctree/cytree.pyx:17-101) on top of the sm_100a tree kernels.

Same names, argument order and meaning:
    Roots(root_num, action_num, tree_nodes)  .prepare / .prepare_no_noise / .get_trajectories /
        .get_distributions / .get_values / .clear / .num
    MinMaxStatsList(num).set_delta(value_delta_max)
    ResultsWrapper(num)
    multi_traverse(roots, pb_c_base, pb_c_init, discount, min_max_stats_lst, results)
    multi_back_propagate(hidden_state_index_x, discount, rewards, values, policies,
                         min_max_stats_lst, results)
plus the upstream-EfficientZero aliases batch_traverse / batch_back_propagate.

Inputs may be Python lists (what the reference takes), numpy arrays or torch tensors on any
device; results are Python lists like the reference's, or CUDA tensors with `as_tensor=True` /
the `*_tensor` accessors so that nothing has to leave the device.  All tree state lives in HBM
inside a `hz_trees` handle (include/hzb200.h); there is no CPU implementation behind this module.

Deviations from the reference:
  * ties between children within 1e-6 of the best score: the reference draws rand() % ties after reseeding libc's
    generator from the clock (cnode.cpp:367-369,409-411), which no one can reproduce.  Default here: the first element
    of the reference's tie list (= its rand() == 0 build, the contract of the bit-exact parity tests).  With
    `Roots.set_tie_break("random", seed)` the pick is uniform over the same tie list from a counter-based generator —
    use it whenever ties are systematic, e.g. with the reference's zero-initialised output heads, where every
    non-root node is an A-way tie and first-index would always descend action 0;
  * hidden_state_index_x passed to multi_back_propagate must be 1, 2, 3, ... in order, which is
    what core/mcts.py:52-55 passes; anything else raises instead of silently mis-indexing;
  * sizes are checked (the reference has no bounds checks: wrong sizes are undefined behaviour).
"""
import collections
import weakref

import numpy as np
import torch

from . import _lib
from ._lib import check, ptr

FLOAT_MAX = 1000000.0  # core/ctree/cminimax.h:7

# (device, num, actions, cap) -> [(hz_trees*, release event)]: handles are reused so that CUDA graphs captured against
# a handle stay valid when the caller builds a new Roots every move (as the reference's callers do).  The cache is
# bounded: beyond MAX_FREE_SHAPES distinct shapes (or MAX_FREE_PER_SHAPE idle handles of one shape) the least
# recently used handles are destroyed (hz_trees_destroy) and everything keyed on them is told to forget them.
_free_handles = collections.OrderedDict()
MAX_FREE_SHAPES = 6
MAX_FREE_PER_SHAPE = 4
_destroy_listeners = []   # weak references to callables(handle_value) — MCTS drops workspaces/graphs of a dead handle


def on_handle_destroyed(fn):
    _destroy_listeners.append(weakref.WeakMethod(fn) if hasattr(fn, "__self__") else weakref.ref(fn))


def _destroy_handle(lib, h):
    for ref in list(_destroy_listeners):
        fn = ref()
        if fn is None:
            _destroy_listeners.remove(ref)
        else:
            fn(h.value)
    lib.hz_trees_destroy(h)


def trim_free_handles(keep_shapes=None, keep_per_shape=None):
    """Destroy idle tree batches beyond the cache bounds (all of them with keep_shapes=0)."""
    keep_shapes = MAX_FREE_SHAPES if keep_shapes is None else keep_shapes
    keep_per_shape = MAX_FREE_PER_SHAPE if keep_per_shape is None else keep_per_shape
    lib = _lib.load()
    while len(_free_handles) > keep_shapes:
        _, idle = _free_handles.popitem(last=False)
        for h, _ in idle:
            _destroy_handle(lib, h)
    for idle in _free_handles.values():
        while len(idle) > keep_per_shape:
            h, _ = idle.pop(0)
            _destroy_handle(lib, h)


def _device_index(device):
    if device is None:
        return torch.cuda.current_device()
    return torch.device(device).index if torch.device(device).index is not None else torch.cuda.current_device()


def as_device(x, dtype, device, shape=None):
    """list / list of rows / numpy / torch (any device) -> contiguous CUDA tensor of `dtype`."""
    if isinstance(x, torch.Tensor):
        if x.dtype == dtype and x.device == device and x.is_contiguous() and not x.requires_grad:
            t = x       # already what the kernels take: no torch call at all
        else:
            t = x.detach().to(device=device, dtype=dtype, non_blocking=True)
    else:
        np_dtype = {torch.float32: np.float32, torch.int32: np.int32, torch.int64: np.int64,
                    torch.uint8: np.uint8}.get(dtype, np.float32)
        t = torch.from_numpy(np.ascontiguousarray(np.asarray(x), dtype=np_dtype)).to(device=device, dtype=dtype)
    t = t.contiguous()
    if shape is not None:
        if t.shape == tuple(shape):
            return t
        if t.numel() != int(np.prod(shape)):
            raise ValueError(f"expected {tuple(shape)} values, got tensor of shape {tuple(t.shape)}")
        t = t.view(*shape)
    return t


class MinMaxStatsList:
    """tools::CMinMaxStatsList (cminimax.h:26-36): per-tree {minimum, maximum}, device resident."""

    def __init__(self, num):
        self.num = int(num)
        self.value_delta_max = 0.0
        self._buf = None

    def set_delta(self, value_delta_max):
        self.value_delta_max = float(value_delta_max)

    def tensor(self, device):
        """float32 [num, 2] = (minimum, maximum), initialised to (+1e6, -1e6) (cminimax.cpp:5-9)."""
        if self._buf is None:
            self._buf = torch.empty(self.num, 2, dtype=torch.float32, device=device)
            self.clear()
        return self._buf

    def clear(self):
        if self._buf is not None:
            self._buf[:, 0] = FLOAT_MAX
            self._buf[:, 1] = -FLOAT_MAX


class ResultsWrapper:
    """tree::CSearchResults (cnode.h:62-73).  The search paths themselves stay in the tree batch's
    HBM buffers; this object carries the per-simulation outputs of the last traverse."""

    def __init__(self, num):
        self.num = int(num)
        self.roots = None
        self.hidden_state_index_x = None  # int32 [num] CUDA
        self.hidden_state_index_y = None
        self.last_actions = None


class Node:
    """Unused stub in the reference as well (cytree.pyx:73-85)."""

    def __init__(self, prior=0.0, action_num=0):
        self.prior, self.action_num = prior, action_num


class Roots:
    """tree::CRoots behind cytree.Roots (cytree.pyx:37-71)."""

    def __init__(self, root_num, action_num, tree_nodes, device=None):
        self.root_num = int(root_num)
        self.action_num = int(action_num)
        self.tree_nodes = int(tree_nodes)
        self.capacity = self.tree_nodes + 1  # the reference reserves action_num*(tree_nodes+2) nodes
        self.device_index = _device_index(device)
        self.device = torch.device("cuda", self.device_index)
        self._key = (self.device_index, self.root_num, self.action_num, self.capacity)
        self._lib = _lib.load()
        free = _free_handles.get(self._key)
        if free:
            _free_handles.move_to_end(self._key)
            self._h, released = free.pop()
            # whatever stream used the handle last must be done with it before this stream touches it
            torch.cuda.current_stream(self.device).wait_event(released)
            check(self._lib.hz_trees_set_tie_break(self._h, 0, 0, 0))
        else:
            h = _lib.C.c_void_p()
            check(self._lib.hz_trees_create(_lib.C.byref(h), self.device_index, self.root_num,
                                            self.action_num, self.capacity))
            self._h = h
        self._prepared = False
        self._keep = None

    def set_tie_break(self, mode="first", seed=0, tree_offset=0):
        """cselect_child's tie rule (cnode.cpp:367-369): "first" = element 0 of the tie list (the reference built with
        rand() == 0; default), "random" = uniform over the list, drawn from a counter-based generator keyed by
        (seed, tree_offset + tree, simulation, depth).  `tree_offset` = global index of this batch's first tree."""
        modes = {"first": 0, "random": 1}
        if mode not in modes:
            raise ValueError("tie-break mode must be 'first' or 'random'")
        check(self._lib.hz_trees_set_tie_break(self.handle, modes[mode], int(seed) & (2 ** 64 - 1), int(tree_offset)))

    # -- lifetime ------------------------------------------------------------------------------
    def release(self):
        """Return the HBM buffers to the per-shape free list (done automatically on __del__)."""
        if getattr(self, "_h", None) is not None:
            ev = torch.cuda.Event()
            ev.record(torch.cuda.current_stream(self.device))
            _free_handles.setdefault(self._key, []).append((self._h, ev))
            _free_handles.move_to_end(self._key)
            self._h = None
            trim_free_handles()

    def __del__(self):
        try:
            self.release()
        except Exception:
            pass

    @property
    def handle(self):
        if self._h is None:
            raise RuntimeError("Roots was cleared")
        return self._h

    @property
    def num(self):
        return self.root_num

    def _stream(self):
        return torch.cuda.current_stream(self.device).cuda_stream

    # -- cytree.Roots API --------------------------------------------------------------------------
    def prepare(self, root_exploration_fraction, noises, reward_pool, policy_logits_pool,
                stack_legal_action):
        """CRoots::prepare (cnode.cpp:247-253)."""
        self._prepare(float(root_exploration_fraction), noises, reward_pool, policy_logits_pool,
                      stack_legal_action)

    def prepare_no_noise(self, reward_pool, policy_logits_pool, stack_legal_action):
        """CRoots::prepare_no_noise (cnode.cpp:255-259)."""
        self._prepare(0.0, None, reward_pool, policy_logits_pool, stack_legal_action)

    def _prepare(self, frac, noises, reward_pool, logits, legal):
        n, a, dev = self.root_num, self.action_num, self.device
        nz = None if noises is None else as_device(noises, torch.float32, dev, (n, a))
        rw = as_device(reward_pool, torch.float32, dev, (n,))
        lg = as_device(logits, torch.float32, dev, (n, a))
        mk = as_device(legal, torch.int32, dev, (n, a))  # float 0/1 masks are truncated like Cython's int conversion
        check(self._lib.hz_trees_prepare(self.handle, self._stream(), frac, ptr(nz), ptr(rw), ptr(lg), ptr(mk)))
        self._keep = (nz, rw, lg, mk)  # keep inputs alive until the stream has consumed them
        self._prepared = True

    def get_distributions_tensor(self):
        """int32 [num, action_num] root child visit counts (CUDA)."""
        out = torch.empty(self.root_num, self.action_num, dtype=torch.int32, device=self.device)
        check(self._lib.hz_trees_root_stats(self.handle, self._stream(), ptr(out), None))
        return out

    def get_values_tensor(self):
        """float32 [num] root values (CUDA)."""
        out = torch.empty(self.root_num, dtype=torch.float32, device=self.device)
        check(self._lib.hz_trees_root_stats(self.handle, self._stream(), None, ptr(out)))
        return out

    def get_stats_tensors(self):
        visits = torch.empty(self.root_num, self.action_num, dtype=torch.int32, device=self.device)
        values = torch.empty(self.root_num, dtype=torch.float32, device=self.device)
        check(self._lib.hz_trees_root_stats(self.handle, self._stream(), ptr(visits), ptr(values)))
        return visits, values

    def get_distributions(self):
        """CRoots::get_distributions (cnode.cpp:276-284): list of lists ([] for unexpanded roots)."""
        if not self._prepared:
            return [[] for _ in range(self.root_num)]
        return self.get_distributions_tensor().cpu().tolist()

    def get_values(self):
        """CRoots::get_values (cnode.cpp:286-292)."""
        if not self._prepared:
            return [0.0] * self.root_num
        return self.get_values_tensor().cpu().tolist()

    def get_trajectories(self):
        """CRoots::get_trajectories (cnode.cpp:266-274): best-action chain of every root."""
        if not self._prepared:
            return [[] for _ in range(self.root_num)]
        max_len = self.capacity + 1
        out = torch.empty(self.root_num, max_len, dtype=torch.int32, device=self.device)
        check(self._lib.hz_trees_trajectories(self.handle, self._stream(), ptr(out), max_len))
        rows = out.cpu().numpy()
        return [row[row >= 0].tolist() for row in rows]

    def clear(self):
        """CRoots::clear (cnode.cpp:261-264)."""
        self._prepared = False
        self._keep = None
        self.release()

    # -- inspection for parity tests ------------------------------------------------------------------
    def export(self, cap=None):
        cap = self.capacity if cap is None else int(cap)
        n, a, dev = self.root_num, self.action_num, self.device
        reward = torch.empty(n, cap, dtype=torch.float32, device=dev)
        value_sum = torch.empty(n, cap, dtype=torch.float32, device=dev)
        visits = torch.empty(n, cap, dtype=torch.int32, device=dev)
        priors = torch.empty(n, a, dtype=torch.float32, device=dev)
        plen = torch.empty(n, dtype=torch.int32, device=dev)
        check(self._lib.hz_trees_export(self.handle, self._stream(), cap, ptr(reward), ptr(value_sum),
                                        ptr(visits), ptr(priors), ptr(plen)))
        return dict(reward=reward, value_sum=value_sum, visits=visits, root_priors=priors, path_len=plen)


def multi_traverse(roots, pb_c_base, pb_c_init, discount, min_max_stats_lst, results,
                   as_tensor=False, pool=None, out_hidden=None, out_action64=None):
    """cytree.multi_traverse (cytree.pyx:97-101 -> cmulti_traverse, cnode.cpp:407-441).

    Returns (hidden_state_index_x_lst, hidden_state_index_y_lst, last_actions): Python lists like
    the reference, or int32 CUDA tensors with as_tensor=True.  Optional fused hand-off: `pool`
    (CUDA tensor [S, num, F]) + `out_hidden` ([num, F]) gathers the parents' hidden states and
    `out_action64` (int64 [num] or [num, 1]) receives the actions, all inside the same kernel.
    """
    n, dev = roots.root_num, roots.device
    ix = torch.empty(n, dtype=torch.int32, device=dev)
    iy = torch.empty(n, dtype=torch.int32, device=dev)
    la = torch.empty(n, dtype=torch.int32, device=dev)
    mm = min_max_stats_lst.tensor(dev)
    if mm.shape[0] != n or results.num != n:
        raise ValueError("MinMaxStatsList / ResultsWrapper size does not match the roots")
    row_bytes = 0
    if pool is not None:
        if out_hidden is None or not pool.is_contiguous() or not out_hidden.is_contiguous():
            raise ValueError("pool and out_hidden must be contiguous CUDA tensors")
        row_bytes = out_hidden.shape[-1] * out_hidden.element_size()
    check(roots._lib.hz_trees_traverse(
        roots.handle, roots._stream(), int(pb_c_base), float(pb_c_init), float(discount), ptr(mm),
        float(min_max_stats_lst.value_delta_max), ptr(ix), ptr(iy), ptr(la), ptr(out_action64),
        ptr(pool), ptr(out_hidden), row_bytes))
    results.roots = roots
    results.hidden_state_index_x, results.hidden_state_index_y, results.last_actions = ix, iy, la
    if as_tensor:
        return ix, iy, la
    packed = torch.stack((ix, iy, la)).cpu().tolist()
    return packed[0], packed[1], packed[2]


def multi_back_propagate(hidden_state_index_x, discount, rewards, values, policies,
                         min_max_stats_lst, results, sanitize_nan=False):
    """cytree.multi_back_propagate (cytree.pyx:87-94 -> cmulti_back_propagate, cnode.cpp:337-344)."""
    roots = results.roots
    if roots is None:
        raise RuntimeError("multi_back_propagate: results carries no traverse (call multi_traverse first)")
    n, a, dev = roots.root_num, roots.action_num, roots.device
    rw = as_device(rewards, torch.float32, dev, (n,))
    vl = as_device(values, torch.float32, dev, (n,))
    pl = as_device(policies, torch.float32, dev, (n, a))
    mm = min_max_stats_lst.tensor(dev)
    check(roots._lib.hz_trees_backprop(roots.handle, roots._stream(), int(hidden_state_index_x),
                                       float(discount), ptr(rw), ptr(vl), ptr(pl),
                                       1 if sanitize_nan else 0, ptr(mm)))
    roots._keep = (rw, vl, pl)


# upstream EfficientZero names used by BASELINE.json's north_star wording
batch_traverse = multi_traverse
batch_back_propagate = multi_back_propagate
