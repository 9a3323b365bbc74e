"""Execution plan for `recurrent_inference` in eval mode: the same function as
HanabiMuZeroNet.recurrent_inference_device (dynamics + reward/value/policy heads + inverse
categorical transforms; /root/reference/core/model.py:74-84, config/hanabi_control/model.py), laid
out for a latency-bound batch so one simulation costs ~16 launches instead of ~70:

  * BatchNorm (running statistics) folded into the preceding Linear;
  * the one-hot action concat of `dynamics` (model.py:199-203, 301-305) replaced by a row lookup
    E[a] = W_fc1[:, F + a] + b added in the GEMM epilogue;
  * the first layer of the three heads run as one GEMM over the shared input;
  * bias + ReLU fused into the cuBLASLt epilogue (torch._addmm_activation) or, where a residual or
    the action row is involved, into one hz_bias_act launch; the next hidden state is written
    straight into its slot of the search's hidden-state pool;
  * value and reward decoded by one hz_support_decode launch.

The GEMMs are still torch/cuBLAS ("the network stays in PyTorch").  Folded weights live in fixed
buffers that `refresh()` rewrites in place whenever the module's parameters change, so CUDA graphs
captured over a plan stay valid across training updates.
"""
import torch

from . import _lib
from ._lib import check, ptr


def _fold(linear, bn):
    """Linear followed by eval-mode BatchNorm1d -> (W', b') in float32."""
    w, b = linear.weight.detach().float(), linear.bias.detach().float()
    if bn is None:
        return w, b
    scale = bn.weight.detach().float() / torch.sqrt(bn.running_var.detach().float() + bn.eps)
    return w * scale[:, None], (b - bn.running_mean.detach().float()) * scale + bn.bias.detach().float()


def _pad_rows(w, b, mult=8):
    out = w.shape[0]
    pad = (-out) % mult
    if pad:
        w = torch.cat((w, w.new_zeros(pad, w.shape[1])))
        b = torch.cat((b, b.new_zeros(pad)))
    return w, b


class RecurrentPlan:
    def __init__(self, net, dtype):
        if dtype not in (torch.float16, torch.float32):
            raise ValueError("plan dtype must be float16 or float32")
        self.net, self.dtype = net, dtype
        self.lib = _lib.load()
        self.F = net.feature_size
        self.A = net.action_space_n
        self.H = net.hidden_size
        self.full = net.full
        self.n_value = net._value_support.numel()
        self.n_reward = net._reward_support.numel()
        self._sig = None
        self._w = {}
        self.refresh(force=True)

    # -- weights ---------------------------------------------------------------------------------------
    def _signature(self):
        return tuple((t.data_ptr(), t._version) for t in list(self.net.parameters()) + list(self.net.buffers()))

    def _set(self, name, value):
        value = value.to(self.dtype).contiguous()
        cur = self._w.get(name)
        if cur is None:
            self._w[name] = value.clone()
        else:
            cur.copy_(value)

    @torch.no_grad()
    def refresh(self, force=False):
        """Re-fold the module's current parameters into the plan's fixed buffers (in place)."""
        sig = self._signature()
        if not force and sig == self._sig:
            return False
        net, F = self.net, self.F
        d = net._dynamics_state
        w1, b1 = _fold(d.fc1, d.bn1)
        self._set("W1", w1[:, :F])                           # [F, F] state part
        self._set("E1", w1[:, F:].t() + b1[None, :])         # [A, F] action row + bias
        for i, (fc, bn) in ((2, (d.fc2, d.bn2)), (3, (d.fc3, d.bn3))):
            w, b = _fold(fc, bn)
            self._set(f"W{i}", w)
            self._set(f"b{i}", b)
        v, r, p = net._prediction_value, net._dynamics_reward, net._prediction_actor
        heads1 = [_fold(h[0], h[1]) for h in (v, r, p)]
        self._set("Wh1", torch.cat([w for w, _ in heads1]))  # [3H, F]: value | reward | actor
        self._set("bh1", torch.cat([b for _, b in heads1]))
        if self.full:
            for name, (fc, bn) in (("v2", (v[3], v[4])), ("r2", (r[3], r[4])), ("a1", (p[3].fc1, p[3].bn1)),
                                   ("a2", (p[3].fc2, p[3].bn2))):
                w, b = _fold(fc, bn)
                self._set("W" + name, w)
                self._set("b" + name, b)
            finals = (("v", v[6]), ("r", r[6]), ("p", p[4]))
        else:
            finals = (("v", v[3]), ("r", r[3]), ("p", p[3]))
        for name, fc in finals:
            w, b = _pad_rows(*_fold(fc, None))
            self._set("Wf" + name, w)
            self._set("bf" + name, b)
        self._sig = sig
        return True

    # -- execution -------------------------------------------------------------------------------------
    def _epi(self, st, out, x, bias=None, residual=None, table=None, idx=None, relu=True):
        check(self.lib.hz_bias_act(st, ptr(out), out.stride(0), ptr(x), x.stride(0), ptr(bias), ptr(residual),
                                   0 if residual is None else residual.stride(0), ptr(table), ptr(idx),
                                   x.shape[0], x.shape[1], 1 if relu else 0, x.element_size()))
        return out

    @torch.no_grad()
    def run(self, hidden, action, out_state):
        """hidden [N, F] (plan dtype), action int64 [N] or [N, 1], out_state [N, F] (plan dtype, may be
        a slot of the hidden-state pool).  Returns (value [N], reward [N], policy_logits [N, A]) fp32."""
        w, H, n = self._w, self.H, hidden.shape[0]
        st = torch.cuda.current_stream(hidden.device).cuda_stream
        act = action.reshape(-1)
        addmm_act = torch._addmm_activation
        # dynamics
        y = torch.mm(hidden, w["W1"].t())
        if self.full:   # relu(bn1(fc1 [s ‖ a]))  ...  relu(bn3(fc3 .) + s)
            self._epi(st, y, y, table=w["E1"], idx=act)
            y = addmm_act(w["b2"], y, w["W2"].t())
            z = torch.mm(y, w["W3"].t())
            self._epi(st, out_state, z, bias=w["b3"], residual=hidden)
        else:           # relu(bn1(fc1 [s ‖ a]) + s)  ...  relu(bn3(fc3 .))
            self._epi(st, y, y, residual=hidden, table=w["E1"], idx=act)
            y = addmm_act(w["b2"], y, w["W2"].t())
            z = torch.mm(y, w["W3"].t())
            self._epi(st, out_state, z, bias=w["b3"])
        # heads: first layers share their input
        h1 = addmm_act(w["bh1"], out_state, w["Wh1"].t())       # [N, 3H] = value | reward | actor
        hv, hr, ha = h1[:, :H], h1[:, H:2 * H], h1[:, 2 * H:]
        if self.full:
            hv = addmm_act(w["bv2"], hv, w["Wv2"].t())
            hr = addmm_act(w["br2"], hr, w["Wr2"].t())
            a1 = addmm_act(w["ba1"], ha, w["Wa1"].t())
            a2 = torch.mm(a1, w["Wa2"].t())
            ha = self._epi(st, a2, a2, bias=w["ba2"], residual=ha)
        wide = w["Wfv"].shape[0]
        vr = torch.empty(2 * n, wide, dtype=self.dtype, device=hidden.device)
        torch.addmm(w["bfv"], hv, w["Wfv"].t(), out=vr[:n])
        torch.addmm(w["bfr"], hr, w["Wfr"].t(), out=vr[n:])
        logits = torch.addmm(w["bfp"], ha, w["Wfp"].t())[:, :self.A].float().contiguous()
        out = torch.empty(2 * n, dtype=torch.float32, device=hidden.device)
        if self.n_value == self.n_reward:
            check(self.lib.hz_support_decode(st, ptr(vr), vr.element_size(), ptr(self.net._value_support), ptr(out),
                                             2 * n, self.n_value, vr.stride(0), self.net.support_delta))
        else:
            check(self.lib.hz_support_decode(st, ptr(vr), vr.element_size(), ptr(self.net._value_support), ptr(out),
                                             n, self.n_value, vr.stride(0), self.net.support_delta))
            check(self.lib.hz_support_decode(st, ptr(vr[n:]), vr.element_size(), ptr(self.net._reward_support),
                                             ptr(out[n:]), n, self.n_reward, vr.stride(0), self.net.support_delta))
        return out[:n], out[n:], logits
