"""StochasticInterpolants (reference: bridge/bridge_model.py:28-447) on the B200 engine.

`sample()` runs sde_vs (:334-387) with the EMA weights (:267) as one native program: per-sample FiLM table, then
n x [36 grouped tcgen05 implicit-GEMM launches for v_net+s_net, Euler-Maruyama update], no host work per step."""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch

from ..ema import ExponentialMovingAverage
from ..engine import BridgeEngine
from ..params import sub_state_dict
from ..schedule import check_model_args
from ..unet import LossProgram
from ..unet_train import LossBackwardProgram
from .networks.conditional_unet_1D_si import InterpolantsConditionalUnet1D


class StochasticInterpolants:
    def __init__(self, model_args=None, precise: bool = False):
        self.precise = precise
        self.net = None
        self.ema = None
        self._engines: Dict[tuple, BridgeEngine] = {}
        self._engine_version: Dict[tuple, int] = {}
        self.noise_override: Optional[torch.Tensor] = None    # [n_steps,B,T,A]: injected N(0,1) draws (parity tests)
        self.step_override: Optional[torch.Tensor] = None     # get_loss: injected U(0,1) draws [B]
        self.z_override: Optional[torch.Tensor] = None        # get_loss: injected N(0,1) draws [B,T,A]
        self._loss_programs: Dict[tuple, list] = {}
        self._seed = int(torch.initial_seed()) & 0x7FFFFFFF       # Philox base seed follows torch.manual_seed
        if model_args:
            self.load_model_args(model_args)

    def load_model_args(self, model_args):
        check_model_args(model_args)                    # NotImplementedError for schedules outside App. B
        self.interpolant_type = model_args['interpolant_type']
        self.gamma_type = model_args['gamma_type']
        self.epsilon_type = model_args['epsilon_type']
        self.prior_policy = model_args['prior_policy']
        self.d = model_args['beta_max']
        self.t_min = 0.001
        self.gamma_inv_max = 200.0
        self.net = None
        self.ema = None
        self.prior_model = None
        self.sde_type = model_args.get('sde_type', 'vs')

    def load_model(self, model_args, device):
        self.load_model_args(model_args)
        if model_args['net_type'] != 'unet1D_si':
            raise NotImplementedError
        self.net = InterpolantsConditionalUnet1D(input_dim=model_args['action_dim'],
                                                 global_cond_dim=model_args['obs_dim'] * model_args['obs_horizon'],
                                                 precise=self.precise)
        self.ema = ExponentialMovingAverage(self.net.parameters(), decay=0.75)
        if model_args['pretrain']:
            checkpoint = torch.load(os.path.join(model_args['ckpt_path'], "bridge_model.pt"), map_location="cpu",
                                    weights_only=False)
            self.net.load_state_dict(checkpoint['net'])
            self.ema.load_state_dict(checkpoint["ema"])
        self.net.to(device)
        self.ema.to(device)
        self.device = device
        self._engines.clear()
        self._loss_programs.clear()      # packed copies of the previous net's weights must not survive a (re)load

    def save_model(self, ckpt_path):
        torch.save({"net": self.net.state_dict(), "ema": self.ema.state_dict()}, os.path.join(ckpt_path, "bridge_model.pt"))

    def train(self):
        return self

    def eval(self):
        return self

    # ---- EMA weights as flat state dicts for the engine ----
    def ema_state_dicts(self):
        names = [n for n, _ in self.net.named_parameters()]
        full = {n: s for n, s in zip(names, self.ema.shadow_params)}
        drift = "b_net." if self.sde_type == 'bs' else "v_net."          # sde_bs integrates b_net directly (:271-273)
        return sub_state_dict(full, drift), sub_state_dict(full, "s_net.")

    def _engine(self, B: int, T: int, diffuse_step: int, inject: bool) -> BridgeEngine:
        key = (B, T, diffuse_step, inject)
        eng = self._engines.get(key)
        if eng is None:
            v_sd, s_sd = self.ema_state_dicts()
            eng = BridgeEngine(dino=None, enc_sd=None, v_sd=v_sd, s_sd=s_sd, action_dim=self.net.input_dim,
                               state_dim=self.net.input_dim, force_dim=0, use_force=False, B=B, T=T, diffuse_step=diffuse_step,
                               beta_max=self.d, device=self.device, precise=self.precise, hidden_dim=self.net.global_cond_dim,
                               inject_noise=inject, sde_type=self.sde_type)
            self._engines[key] = eng
            self._engine_version[key] = self.ema.version
        elif self._engine_version[key] != self.ema.version:
            eng.refresh_unet(*self.ema_state_dicts())
            self._engine_version[key] = self.ema.version
        return eng

    @torch.no_grad()
    def sample(self, x_prior, cond, diffuse_step=10, recod_traj=False):
        """x_prior [B,T,A] (normalised), cond [B,obs_dim] -> x_target [B,T,A] (and the trajectory list)."""
        if self.sde_type not in ('vs', 'bs'):
            raise NotImplementedError
        B, T, A = x_prior.shape
        inject = self.noise_override is not None
        eng = self._engine(B, T, diffuse_step, inject)
        eng.x.copy_(x_prior)
        eng.cond.copy_(cond)
        if inject:
            eng.noise.copy_(self.noise_override)
        else:
            self._seed += 1
            eng.seed.fill_(self._seed)
        eng.run_ranges(["film_c", "xprior"])
        if not recod_traj:
            eng.run_steps()
            return eng.x.clone()
        traj = [eng.x.clone()]
        for k in range(eng.n_steps):
            eng.run_steps(k, 1)
            traj.append(eng.x.clone())
        return traj[-1], traj

    def get_loss(self, batch_dict, device=None):
        """bridge_model.py:220-246: (loss, {'v_loss', 's_loss', 'b_loss'}) on the live b_net / v_net / s_net weights.

        With autograd enabled and trainable nets (the training loop, bridge_train.py:315-334) the returned loss is
        differentiable: one native program computes the three losses AND their gradients (unet_train.LossBackwardProgram:
        training forward of the three U-Nets, explicit backward on the same tcgen05 GEMM kernel), and `loss.backward()`
        hands them to the nets' nn.Parameters and, through d loss / d obs_cond, to whatever produced `obs_cond` (the state
        encoder).  Under torch.no_grad() (`_validate`, bridge_train.py:380-438) only the forward program runs.  bf16 operands
        in training; `precise=True` controllers train in bf16 as well (the fp32 split-tf32 mode is inference only)."""
        device = device or self.device
        obs = batch_dict['obs_cond'].to(device)
        naction = batch_dict['expert_act'].to(device).float()
        if 'vla_act' in batch_dict:
            prior_action = batch_dict['vla_act'].to(device).float()
        else:
            prior_action = torch.randn(naction.shape).float().to(naction.device)
        B, T, A = naction.shape
        step = self.step_override if self.step_override is not None else torch.rand(B, device=device)
        z = self.z_override if self.z_override is not None else torch.randn_like(naction)
        params = list(self.net.parameters())
        if torch.is_grad_enabled() and (obs.requires_grad or any(p.requires_grad for p in params)):
            return self._get_loss_with_grad(obs, prior_action, naction, step, z, params)
        with torch.no_grad():
            return self._get_loss_value(obs.float().flatten(1), prior_action, naction, step, z)

    def _net_state_dicts(self):
        sd = {k: v.detach() for k, v in self.net.state_dict().items()}
        return [sub_state_dict(sd, "b_net."), sub_state_dict(sd, "v_net."), sub_state_dict(sd, "s_net.")]

    def _net_param_dicts(self):
        """{key: Parameter} of b_net / v_net / s_net in the order of `_net_state_dicts` (only parameters: the nets have no buffers)"""
        out = []
        for name in ("b_net", "v_net", "s_net"):
            out.append({k: p for k, p in getattr(self.net, name).named_parameters()})
        return out

    def _weights_token(self):
        """Identity + version of the live parameters: a program's packed operand copies are current iff its token matches.
        (Identity matters: a freshly loaded net has the same version sum as the one it replaces.)"""
        params = list(self.net.parameters())
        return (id(self.net), sum(p._version for p in params), getattr(self.ema, "param_writes", 0))

    def train_program(self, B: int, T: int) -> LossBackwardProgram:
        """The forward + backward program of get_loss for a batch shape (built once, weights re-packed in place)."""
        key = (B, T, "bwd")
        ent = self._loss_programs.get(key)
        if ent is None:
            ent = [LossBackwardProgram(self._net_state_dicts(), self.net.input_dim, B, T, float(self.d), self.device), self._weights_token()]
            self._loss_programs[key] = ent
        return ent[0]

    def sync_train_program(self, prog: LossBackwardProgram) -> None:
        """Re-pack the operand copies if the parameters changed since the program last saw them."""
        for ent in self._loss_programs.values():
            if ent[0] is prog:
                tok = self._weights_token()
                if ent[1] != tok:
                    if getattr(prog, "_gather", None) is None and not getattr(prog, "_gather_tried", False):
                        # first re-pack: build and verify the gather maps; the parameters move into one contiguous arena
                        # (p.data becomes a view of it)
                        prog._gather_tried = True
                        if os.environ.get("VT_GATHER_REPACK", "1") != "0":
                            prog.setup_gather(self._net_param_dicts(), self._net_state_dicts())
                    if getattr(prog, "_gather", None) is not None and not prog.gather_valid():
                        prog._gather = None                 # parameters were moved out of the arena: tensor-op re-pack from now on
                    if getattr(prog, "_gather", None) is not None:
                        prog.refresh_gather()
                    else:
                        prog.refresh_graphed(self._net_state_dicts())
                    ent[1] = self._weights_token()
                return
        raise KeyError("not a program of this model")

    def _get_loss_with_grad(self, obs, x0, x1, step, z, params):
        B, T, A = x1.shape
        prog = self.train_program(B, T)
        self.sync_train_program(prog)
        names = [n for n, _ in self.net.named_parameters()]
        out = _BridgeLossFn.apply(prog, names, x0, x1, step, z, obs.float().flatten(1), *params)
        return out[0], {'v_loss': out[1].detach(), 's_loss': out[2].detach(), 'b_loss': out[3].detach()}

    def _get_loss_value(self, nobs, prior_action, naction, step, z):
        device = self.device
        B, T, A = naction.shape
        key = (B, T)
        version = self._weights_token()
        ent = self._loss_programs.get(key)
        sds = self._net_state_dicts()
        if ent is None:
            ent = [LossProgram(sds, A, B, T, float(self.d), device, self.precise), version]
            self._loss_programs[key] = ent
        elif ent[1] != version:
            ent[0].refresh(sds)
            ent[1] = version
        out = ent[0](prior_action, naction, nobs, step, z).clone()
        return out[0], {'v_loss': out[1], 's_loss': out[2], 'b_loss': out[3]}


class _BridgeLossFn(torch.autograd.Function):
    """loss[4] = (total, v, s, b) with the gradients of the TOTAL computed eagerly by the native program in forward();
    backward() scales them by the incoming gradient of loss[0] (the three partial losses are reported for logging only,
    bridge_train.py:318-327, and are returned detached by get_loss)."""

    @staticmethod
    def forward(ctx, prog, names, x0, x1, step, z, obs, *params):
        prog.set_inputs(x0, x1, obs, step, z)
        out = prog.run().clone()
        g = prog.grads
        # the gradients stay in the program's own buffers (no 412 MB copy per step); backward() checks they are still current
        ctx.prog, ctx.run_id = prog, prog.runs
        ctx.save_for_backward(prog.d_cond, *[g[n] for n in names])
        return out

    @staticmethod
    def backward(ctx, gout):
        if ctx.prog.runs != ctx.run_id:
            raise RuntimeError("get_loss() was called again before this loss.backward(): the program's gradient buffers now hold "
                               "the later call's gradients (call backward() before the next get_loss(), as bridge_train.py does)")
        d_cond, *grads = ctx.saved_tensors
        s = gout[0]
        return (None, None, None, None, None, None, s * d_cond) + tuple(torch._foreach_mul(list(grads), s))
