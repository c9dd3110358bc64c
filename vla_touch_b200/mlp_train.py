"""Training plans of the 3-layer GELU encoders: DiffusionController.state_encoder (bridge_controller.py:42-48) and
TactileLSTMController.obs_encoder (lstm_step_controller.py:40-46), Sequential(Linear, GELU, Linear, GELU, Linear).

Forward keeps the GELU inputs; backward (from d loss / d obs_cond, which the diffusion / LSTM training program produces) is
dgrad / wgrad GEMMs on gemm_tc_kernel (the batch is the K dimension of the weight gradients), `a * gelu'(b)` and column sums.
`MlpTrainFn` exposes it to torch autograd: forward() runs the forward range of the program, backward() the backward range.
GPU-checked by tests/test_zz_backward_gpu.py::test_native_encoder_training.
"""
from __future__ import annotations

from typing import Dict, Sequence

import torch

from . import native as nv
from . import unet_bwd as ub
from .lstm_train import _ew, _pack, _Packer
from .plan import Plan, linear_desc, ptr, round_up
from .unet import _View


class MlpTrainProgram:
    """x [B][in_dim] fp32 -> out [B][H] fp32, and from d_out [B][H] the gradients of '0.weight', '0.bias', '2.weight', '2.bias',
    '4.weight', '4.bias' (nn.Sequential keys)."""

    def __init__(self, sd: Dict[str, torch.Tensor], B: int, device):
        self.plan = p = Plan(device)
        f32, bf = torch.float32, torch.bfloat16
        dims = [sd["0.weight"].shape[1], sd["0.weight"].shape[0], sd["2.weight"].shape[0], sd["4.weight"].shape[0]]
        assert all(d % 64 == 0 for d in dims[1:]), dims
        self.in_dim, self.B = dims[0], B
        kpad = round_up(dims[0], 64)
        self.pk = pk = _Packer(p, {"mlp": sd})
        G = lambda k: (lambda m: m["mlp"][k])
        self.x = p.buf("in.x", (B, dims[0]), f32)
        self.d_out = p.buf("in.d_out", (B, dims[3]), f32)
        x_op = p.buf("x_op", (B, kpad), bf)
        _pack(p, self.x, dims[0], B, dims[0], x_op, kpad, 0, nv.ACT_NONE, "mlp.x -> operand", zero_to=kpad)
        acts, ops = [], [x_op]                                     # pre-activations (fp32), GEMM operands (bf16)
        for i, key in enumerate(("0", "2", "4")):
            k_in = kpad if i == 0 else dims[i]
            a = p.buf(f"a{i}", (B, dims[i + 1]), f32)
            p.add(linear_desc(a=ops[-1], rows=B, k=k_in, a_ld=k_in, w=pk.lin(G(key + ".weight"), k_in), n=dims[i + 1],
                              n_pad=dims[i + 1], w_ld=k_in, out=a, ldc=dims[i + 1], bias=pk.vec(G(key + ".bias"))), f"mlp.{key}")
            acts.append(a)
            if i < 2:
                g = p.buf(f"g{i}", (B, dims[i + 1]), bf)
                _pack(p, a, dims[i + 1], B, dims[i + 1], g, dims[i + 1], 0, nv.ACT_GELU, f"mlp.gelu{i}")
                ops.append(g)
        self.out = acts[-1]
        self.n_forward_ops = len(p)
        ctx = ub.DgradCtx(1, precise=False)
        V = lambda t: _View(t.view(1, B, 1, t.shape[-1]), 1, t.shape[-1])
        self.grads: Dict[str, torch.Tensor] = {}
        d = self.d_out
        for i, key in reversed(list(enumerate(("0", "2", "4")))):
            dw = ub.conv_wgrad(p, ctx, B, V(d), V(ops[i]), tap_off=[0], t_out=1, tag=f"mlp.{key}.wgrad")[0]
            self.grads[key + ".weight"] = dw[:, : self.in_dim] if i == 0 else dw
            self.grads[key + ".bias"] = ub.colsum(p, 1, B, V(d), 1, f"mlp.{key}.dbias")[0]
            if i > 0:
                db_ = ub.cast_bf16(p, 1, B, V(d), 1, f"mlp.{key}.dout.bf16")
                dg = p.buf(f"dg{i}", (B, dims[i]), f32)
                p.add(linear_desc(a=db_.t, rows=B, k=dims[i + 1], a_ld=dims[i + 1], w=pk.lin(G(key + ".weight"), dims[i + 1], transpose=True),
                                  n=dims[i], n_pad=dims[i], w_ld=dims[i + 1], out=dg, ldc=dims[i]), f"mlp.{key}.dgrad")
                da = p.buf(f"da{i}", (B, dims[i]), f32)
                _ew(p, ptr(dg), dims[i], ptr(acts[i - 1]), dims[i], da, B, dims[i], nv.EW_GELU_BWD, f"mlp.gelu{i - 1}.bwd")
                d = da

    def refresh(self, sd: Dict[str, torch.Tensor]) -> None:
        self.pk.refresh({"mlp": sd})

    def forward(self, x: torch.Tensor) -> torch.Tensor:
        self.x.copy_(x)
        self.runs = getattr(self, "runs", 0) + 1
        self.plan.compile().run(0, self.n_forward_ops)
        return self.out

    def backward(self, d_out: torch.Tensor) -> Dict[str, torch.Tensor]:
        self.d_out.copy_(d_out)
        self.plan.compile().run(self.n_forward_ops, len(self.plan) - self.n_forward_ops)
        return self.grads


class MlpTrainFn(torch.autograd.Function):
    """out = encoder(x) with the parameter gradients computed by the native backward range."""

    @staticmethod
    def forward(ctx, prog: MlpTrainProgram, names: Sequence[str], x, *params):
        ctx.prog, ctx.names = prog, names
        out = prog.forward(x).clone()
        ctx.run_id = prog.runs
        return out

    @staticmethod
    def backward(ctx, gout):
        if ctx.prog.runs != ctx.run_id:
            raise RuntimeError("the encoder was evaluated again (same batch size) before this backward(): its saved activations "
                               "were overwritten -- call backward() first, or evaluate the second batch under torch.no_grad()")
        g = ctx.prog.backward(gout.contiguous())
        return (None, None, None) + tuple(g[n].clone() for n in ctx.names)


def encoder_forward(module: torch.nn.Sequential, cache: dict, x: torch.Tensor) -> torch.Tensor:
    """Differentiable (w.r.t. the module's parameters) native forward of a Sequential(Linear, GELU, Linear, GELU, Linear)."""
    B = x.shape[0]
    ver = tuple(p._version for p in module.parameters())
    ent = cache.get(B)
    sd = {k: v.detach() for k, v in module.state_dict().items()}
    if ent is None:
        ent = cache[B] = [MlpTrainProgram(sd, B, x.device), ver]
    elif ent[1] != ver:
        ent[0].refresh(sd)
        ent[1] = ver
    names = [n for n, _ in module.named_parameters()]
    return MlpTrainFn.apply(ent[0], names, x.detach().float(), *module.parameters())
