"""`ActorCritic` with the reference's interface (algo/ppo/actor_critic.py:36-128) on the tcgen05 GEMM path.

Parameters live in ONE flat fp32 buffer so that gradient all-reduce, clip_grad_norm_ and Adam are single
passes.  Each `nn.Linear` is a packed matrix `P[out_pad, ld]` = `[W | b | 0-pad]` with `ld = pitch(in + 1)` (whole 128-byte rows):
  * the forward GEMM reads `W` through a TMA tensor map with K extent `in` and takes the bias from column `in`;
  * the weight-gradient GEMM multiplies by the activations extended with a constant ones column, so the bias
    gradient is simply column `in` of the packed gradient — no separate reduction kernel;
  * rows are 16-byte aligned (615*4 and 1050*4 are not, SURVEY.md §7 hard part 2).
`state_dict()` / `load_state_dict()` use the reference's keys (`actor.{0,2,4,6}.{weight,bias}`,
`critic.{0,2,4,6}.{weight,bias}`, `std`), so checkpoints interchange.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, List, Optional

import torch

from .. import _lib
from .._lib import (GemmDesc, HB_EPI_ATOMIC_ADD, HB_EPI_BIAS, HB_EPI_BIAS_ELU, HB_EPI_ELU_BWD, HB_GEMM_3XTF32, HB_GEMM_TF32,
                    PPO_NUM_ACTIONS)


def pitch(n: int) -> int:
    """Row pitch (floats) of the buffers the TMA reads and writes here: whole 128-byte lines.  A 16-byte-aligned pitch is
    all the TMA asks for, but every 128-byte box row then straddles two lines and five 32-byte sectors instead of four:
    measured on the layer-1 forward GEMM (K = 1050), 79.9 us at a pitch of 1052 floats against 68.2 us at 1056."""
    return (n + 31) // 32 * 32


def pad4(n: int) -> int:
    return (n + 3) // 4 * 4


class _Layer:
    __slots__ = ("net", "index", "fan_in", "fan_out", "rows", "ld", "offset", "last")

    def __init__(self, net, index, fan_in, fan_out, last, offset):
        self.net, self.index, self.fan_in, self.fan_out, self.last, self.offset = net, index, fan_in, fan_out, last, offset
        self.rows = (fan_out + 15) // 16 * 16 if last else fan_out          # output layers are padded to a multiple of the smallest UMMA N
        self.ld = pitch(fan_in + 1)

    @property
    def numel(self):
        return self.rows * self.ld


def _desc(lib, scratch, kw):
    d = GemmDesc()
    for k, v in kw.items():
        setattr(d, k, v)
    if scratch is not None:
        d.precision = HB_GEMM_3XTF32
        need = int(lib.hb_gemm_workspace_floats(C.byref(d)))
        if scratch[0] is None or scratch[0].numel() < need:
            scratch[0] = torch.empty(need, device=scratch[1])
        d.workspace, d.workspace_floats = scratch[0].data_ptr(), scratch[0].numel()
    return d


def gemm(lib, st, scratch=None, **kw):
    """One hb_gemm_tf32 call.  `scratch` (a one-element list holding a float tensor or None) switches the call to
    HB_GEMM_3XTF32 and supplies / grows the workspace of the split operands."""
    _lib.check(lib.hb_gemm_tf32(C.byref(_desc(lib, scratch, kw)), st), "hb_gemm_tf32")


def gemm2(lib, st, scratch0, kw0, scratch1, kw1):
    """The same GEMM of both networks in one launch (hb_gemm_tf32_grouped)."""
    _lib.check(lib.hb_gemm_tf32_grouped(C.byref(_desc(lib, scratch0, kw0)), C.byref(_desc(lib, scratch1, kw1)), st),
               "hb_gemm_tf32_grouped")


class ActorCritic:
    is_recurrent = False

    def __init__(self, num_actor_obs, num_critic_obs, num_actions, actor_hidden_dims=(256, 256, 256),
                 critic_hidden_dims=(256, 256, 256), init_noise_std=1.0, activation=None, device="cuda:0", **kwargs):
        precision = kwargs.pop("precision", "tf32")
        if precision not in ("tf32", "3xtf32"):
            raise ValueError("precision must be 'tf32' or '3xtf32'")
        if kwargs:
            print("ActorCritic.__init__ got unexpected arguments, which will be ignored: " + str(list(kwargs)))
        if num_actions not in PPO_NUM_ACTIONS:
            raise ValueError(f"the loss-head kernels are built for {PPO_NUM_ACTIONS} actions (hector, XBot-L, hector_full)")
        if len(actor_hidden_dims) != 3 or len(critic_hidden_dims) != 3:
            raise ValueError("three hidden layers per network (hector_config.py:207-210)")
        self._lib = _lib.load()
        self.device = torch.device(device)
        self.num_actor_obs, self.num_critic_obs, self.num_actions = num_actor_obs, num_critic_obs, num_actions
        self.layers: List[_Layer] = []
        off = 0
        for net, dims in (("actor", (num_actor_obs, *actor_hidden_dims, num_actions)),
                          ("critic", (num_critic_obs, *critic_hidden_dims, 1))):
            for i in range(4):
                L = _Layer(net, 2 * i, dims[i], dims[i + 1], i == 3, off)
                self.layers.append(L)
                off += L.numel
        self._std_offset = off
        off += (num_actions + 15) // 16 * 16
        self.flat = torch.zeros(off, device=self.device)
        # the gradient buffer carries 8 spare floats behind the parameters: a data-parallel replica parks its loss
        # statistics there so that ONE all-reduce per optimizer step covers both (isaac_b200/parallel.py)
        self._grad_wire = torch.zeros(off + 8, device=self.device)
        self.grad = self._grad_wire[:off]
        # default nn.Linear initialisation, drawn in the reference's module construction order
        # (actor layers, critic layers, then std: actor_critic.py:54-83) so the same torch seed gives the same net
        for L in self.layers:
            lin = torch.nn.Linear(L.fan_in, L.fan_out)
            self._matrix(self.flat, L)[:L.fan_out, :L.fan_in] = lin.weight.detach().to(self.device)
            self._matrix(self.flat, L)[:L.fan_out, L.fan_in] = lin.bias.detach().to(self.device)
        self.flat[self._std_offset:self._std_offset + num_actions] = init_noise_std
        self.distribution = None
        self._ws: Dict[int, dict] = {}
        self._last: Optional[dict] = None
        # "tf32": operands truncated to TF32 by the tensor core (fast path).  "3xtf32": hi/lo split operands, three
        # partial products in one fp32 accumulator - fp32-grade results like the reference's nn.Linear (parity mode)
        self.precision = precision
        self.grouped = True          # the actor's and the critic's GEMM of a layer share one launch (hb_gemm_tf32_grouped)
        self._scratch = {"actor": [None, self.device], "critic": [None, self.device]}     # one per stream / network

    def rebind(self, flat: torch.Tensor, grad_wire: torch.Tensor) -> None:
        """Move the flat parameter / gradient buffers to caller-provided storage of the same size (symmetric memory for
        data-parallel replicas); contents are the caller's responsibility."""
        if flat.numel() != self.flat.numel() or grad_wire.numel() != self._grad_wire.numel():
            raise ValueError("rebind needs buffers of the same size")
        self.flat, self._grad_wire = flat, grad_wire
        self.grad = grad_wire[:flat.numel()]

    # ------------------------------------------------------------------ parameter views
    def _matrix(self, flat, L):
        return flat[L.offset:L.offset + L.numel].view(L.rows, L.ld)

    @property
    def std(self):
        return self.flat[self._std_offset:self._std_offset + self.num_actions]

    def named_parameters(self):
        """Views in the order of the reference module's `parameters()` (std first, then actor, critic)."""
        yield "std", self.std
        for L in self.layers:
            P = self._matrix(self.flat, L)
            yield f"{L.net}.{L.index}.weight", P[:L.fan_out, :L.fan_in]
            yield f"{L.net}.{L.index}.bias", P[:L.fan_out, L.fan_in]

    def named_gradients(self):
        yield "std", self.grad[self._std_offset:self._std_offset + self.num_actions]
        for L in self.layers:
            G = self._matrix(self.grad, L)
            yield f"{L.net}.{L.index}.weight", G[:L.fan_out, :L.fan_in]
            yield f"{L.net}.{L.index}.bias", G[:L.fan_out, L.fan_in]

    def parameters(self):
        return [p for _, p in self.named_parameters()]

    def state_dict(self):
        return OrderedDict((k, v.detach().clone().contiguous()) for k, v in self.named_parameters())

    def load_state_dict(self, sd, strict=True):
        views = dict(self.named_parameters())
        if strict and set(sd) != set(views):
            raise KeyError(f"state_dict keys differ: {sorted(set(sd) ^ set(views))}")
        for k, v in sd.items():
            if k in views:
                views[k].copy_(v.to(self.device))

    def to(self, device):
        if torch.device(device) != self.device:
            raise ValueError("ActorCritic lives on the device it was created on")
        return self

    def train(self, mode=True):
        return self

    def eval(self):
        return self

    def reset(self, dones=None):
        pass

    def forward(self):
        raise NotImplementedError

    # ------------------------------------------------------------------ workspaces
    def workspace(self, m: int) -> dict:
        """Activation buffers for a batch of m rows; hidden activations carry the constant ones column at
        index `width` that feeds the bias gradients."""
        ws = self._ws.get(m)
        if ws is None:
            z = lambda r, c: torch.zeros(r, c, device=self.device)
            ws = {"m": m}
            for net in ("actor", "critic"):
                Ls = [L for L in self.layers if L.net == net]
                hs = []
                for L in Ls[:-1]:
                    h = z(m, pitch(L.fan_out + 1))
                    h[:, L.fan_out] = 1.0
                    hs.append(h)
                ws[net] = {"h": hs, "out": z(m, Ls[-1].rows), "d_out": z(m, Ls[-1].rows),
                           "dz": [z(m, pitch(L.fan_out)) for L in Ls[:-1]]}
            self._ws[m] = ws
        return ws

    def _scratch_for(self, net: str):
        return self._scratch[net] if self.precision == "3xtf32" else None

    # ------------------------------------------------------------------ forward
    @property
    def fused_head(self) -> bool:
        """Both last hidden layers are 128 wide (and at most 15 actions): the update path evaluates the output layers inside
        the loss kernel (hb_ppo_head_fused) instead of as 128-row tensor-core tiles with 10 / 1 useful columns."""
        return all(L.fan_in == 128 for L in self.layers if L.last) and self.num_actions <= 15

    def _mlp_forward(self, net: str, x: torch.Tensor, ws: dict, hidden_only: bool = False):
        """x: [m, ld] with ld % 4 == 0 (only the first fan_in columns are read).  Returns out [m,16]
        (or, with hidden_only, the last hidden activations [m, pitch(width + 1)])."""
        lib, st = self._lib, torch.cuda.current_stream(self.device).cuda_stream
        Ls = [L for L in self.layers if L.net == net]
        a, lda = x, x.stride(0)
        m = x.shape[0]
        for i, L in enumerate(Ls):
            if L.last and hidden_only:
                return ws[net]["h"][-1]
            P = self._matrix(self.flat, L)
            d = ws[net]["out"] if L.last else ws[net]["h"][i]
            gemm(lib, st, self._scratch_for(net), A=a.data_ptr(), B=P.data_ptr(), D=d.data_ptr(), M=m, N=L.rows if L.last else L.fan_out,
                 K=L.fan_in, lda=lda, ldb=L.ld, ldd=d.stride(0), epilogue=HB_EPI_BIAS if L.last else HB_EPI_BIAS_ELU,
                 bias=P.data_ptr() + L.fan_in * 4, bias_stride=L.ld)
            a, lda = d, d.stride(0)
        return ws[net]["out"]

    # ------------------------------------------------------------------ both networks, layer by layer, grouped launches
    def _fwd_kw(self, net, i, a, lda, m, ws):
        L = [L for L in self.layers if L.net == net][i]
        P = self._matrix(self.flat, L)
        d = ws[net]["out"] if L.last else ws[net]["h"][i]
        return dict(A=a.data_ptr(), B=P.data_ptr(), D=d.data_ptr(), M=m, N=L.rows if L.last else L.fan_out, K=L.fan_in,
                    lda=lda, ldb=L.ld, ldd=d.stride(0), epilogue=HB_EPI_BIAS if L.last else HB_EPI_BIAS_ELU,
                    bias=P.data_ptr() + L.fan_in * 4, bias_stride=L.ld), d

    def forward_both(self, xa: torch.Tensor, xc: torch.Tensor, ws: dict):
        """The three hidden layers of the actor (on xa) and the critic (on xc): three grouped launches instead of six.
        Returns the last hidden activations (h3_actor, h3_critic)."""
        lib, st = self._lib, torch.cuda.current_stream(self.device).cuda_stream
        m = xa.shape[0]
        a, c = (xa, xa.stride(0)), (xc, xc.stride(0))
        for i in range(3):
            kwa, da = self._fwd_kw("actor", i, a[0], a[1], m, ws)
            kwc, dc = self._fwd_kw("critic", i, c[0], c[1], m, ws)
            gemm2(lib, st, self._scratch_for("critic"), kwc, self._scratch_for("actor"), kwa)      # the larger problem first
            a, c = (da, da.stride(0)), (dc, dc.stride(0))
        return ws["actor"]["h"][-1], ws["critic"]["h"][-1]

    def backward_both(self, xa: torch.Tensor, xc: torch.Tensor, ws: dict):
        """Gradients of the three hidden layers of both networks from dz3 (left in ws by hb_ppo_head_fused): per layer one
        grouped data-gradient launch (critical path first) and one grouped weight-gradient launch - five launches for ten
        GEMMs."""
        lib, st = self._lib, torch.cuda.current_stream(self.device).cuda_stream
        m = xa.shape[0]
        nets = {"actor": xa, "critic": xc}
        cur = {net: (ws[net]["dz"][-1], ws[net]["dz"][-1].stride(0)) for net in nets}
        for i in (2, 1, 0):
            kw_w, kw_d, nxt = {}, {}, {}
            for net, x in nets.items():
                L = [L for L in self.layers if L.net == net][i]
                G, P = self._matrix(self.grad, L), self._matrix(self.flat, L)
                d_cur, ld_cur = cur[net]
                act_in = x if i == 0 else ws[net]["h"][i - 1]
                kw_w[net] = dict(A=d_cur.data_ptr(), B=act_in.data_ptr(), D=G.data_ptr(), M=L.rows, N=L.fan_in + 1, K=m, lda=ld_cur,
                                 ldb=act_in.stride(0), ldd=L.ld, a_mn_major=1, b_mn_major=1, epilogue=HB_EPI_ATOMIC_ADD, split_k=0)
                if i > 0:
                    h_prev, dz = ws[net]["h"][i - 1], ws[net]["dz"][i - 1]
                    kw_d[net] = dict(A=d_cur.data_ptr(), B=P.data_ptr(), D=dz.data_ptr(), M=m, N=L.fan_in, K=L.rows, lda=ld_cur,
                                     ldb=L.ld, ldd=dz.stride(0), b_mn_major=1, epilogue=HB_EPI_ELU_BWD, H=h_prev.data_ptr(),
                                     ldh=h_prev.stride(0))
                    nxt[net] = (dz, dz.stride(0))
            if i > 0:
                gemm2(lib, st, self._scratch_for("critic"), kw_d["critic"], self._scratch_for("actor"), kw_d["actor"])
            gemm2(lib, st, self._scratch_for("critic"), kw_w["critic"], self._scratch_for("actor"), kw_w["actor"])
            cur = nxt

    def _mlp_backward(self, net: str, x: torch.Tensor, ws: dict, from_hidden: bool = False):
        """Gradients of every packed matrix of `net` from d_out (filled by the loss head); accumulates into
        self.grad with split-K atomics (the buffer is zero on entry: Adam zeroes it).  With from_hidden the
        output layer has been handled by hb_ppo_head_fused, which left dz of the last hidden layer in ws."""
        lib, st = self._lib, torch.cuda.current_stream(self.device).cuda_stream
        Ls = [L for L in self.layers if L.net == net]
        m = x.shape[0]
        d_cur, ld_cur = ws[net]["d_out"], ws[net]["d_out"].stride(0)               # gradient w.r.t. the layer's pre-activation output
        if from_hidden:
            d_cur, ld_cur = ws[net]["dz"][-1], ws[net]["dz"][-1].stride(0)
        for i in reversed(range(3 if from_hidden else 4)):
            L = Ls[i]
            G = self._matrix(self.grad, L)
            act_in = x if i == 0 else ws[net]["h"][i - 1]      # [m, fan_in (+ ones column)]
            n_w = L.fan_in + 1
            # weight gradient: G[rows, fan_in + 1] += d_cur^T [rows, m] * [act_in | 1] [m, fan_in + 1]
            gemm(lib, st, self._scratch_for(net), A=d_cur.data_ptr(), B=act_in.data_ptr(), D=G.data_ptr(), M=L.rows, N=n_w, K=m, lda=ld_cur,
                 ldb=act_in.stride(0), ldd=L.ld, a_mn_major=1, b_mn_major=1, epilogue=HB_EPI_ATOMIC_ADD, split_k=0)
            if i == 0:
                break
            # data gradient through the previous ELU: dz_prev = (d_cur * W) . elu'(h_prev)
            P = self._matrix(self.flat, L)
            h_prev, dz = ws[net]["h"][i - 1], ws[net]["dz"][i - 1]
            gemm(lib, st, self._scratch_for(net), A=d_cur.data_ptr(), B=P.data_ptr(), D=dz.data_ptr(), M=m, N=L.fan_in, K=L.rows, lda=ld_cur,
                 ldb=L.ld, ldd=dz.stride(0), b_mn_major=1, epilogue=HB_EPI_ELU_BWD, H=h_prev.data_ptr(),
                 ldh=h_prev.stride(0))
            d_cur, ld_cur = dz, dz.stride(0)

    def _as_operand(self, x: torch.Tensor, width: int) -> torch.Tensor:
        """TMA needs 16-byte aligned rows: use x in place when its row stride allows, else stage a padded copy."""
        if x.is_cuda and x.dtype == torch.float32 and x.stride(-1) == 1 and x.stride(0) % 4 == 0 and x.data_ptr() % 16 == 0:
            return x
        buf = torch.zeros(x.shape[0], pitch(width + 1), device=self.device)
        buf[:, :width] = x
        return buf

    # ------------------------------------------------------------------ reference-named API
    def act_inference(self, observations):
        ws = self.workspace(observations.shape[0])
        out = self._mlp_forward("actor", self._as_operand(observations, self.num_actor_obs), ws)
        return out[:, :self.num_actions].clone()

    def evaluate(self, critic_observations, **kwargs):
        ws = self.workspace(critic_observations.shape[0])
        out = self._mlp_forward("critic", self._as_operand(critic_observations, self.num_critic_obs), ws)
        return out[:, :1].clone()

    def update_distribution(self, observations):
        ws = self.workspace(observations.shape[0])
        out = self._mlp_forward("actor", self._as_operand(observations, self.num_actor_obs), ws)
        mean = out[:, :self.num_actions]
        self.distribution = torch.distributions.Normal(mean, mean * 0.0 + self.std, validate_args=False)

    def act(self, observations, **kwargs):
        self.update_distribution(observations)
        return self.distribution.sample()

    def get_actions_log_prob(self, actions):
        return self.distribution.log_prob(actions).sum(dim=-1)

    @property
    def action_mean(self):
        return self.distribution.mean

    @property
    def action_std(self):
        return self.distribution.stddev

    @property
    def entropy(self):
        return self.distribution.entropy().sum(dim=-1)
