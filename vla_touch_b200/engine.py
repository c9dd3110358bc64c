"""BridgeEngine: the whole DiffusionController.predict() hot path as ONE native program per input configuration
(reference call stack: bridge_controller.py:149-182 -> :112-134 -> visual_encoder.py:56-106 -> HF Dinov2Model;
controller_dataset.py:303-384; bridge/bridge_model.py:259-279,334-387; conditional_unet_1D.py:194-247).

    images x2 -> [imgstats, patchify, patch-embed, 12 x ViT block, CLS LayerNorm]  -> features
    features, state, force -> state_encoder (3 GEMMs)                              -> cond
    cond -> per-sample FiLM table;  normalise(vla chunk)                           -> x_0
    n x [36 grouped implicit-GEMM launches (v_net + s_net), Euler-Maruyama update] -> x_n
    de-normalise                                                                   -> refined chunk

No host synchronisation happens inside: the reference's data-dependent branches (`images.max() > 1`,
`images.mean() < 0.5`) are evaluated on the device, the per-step scalar schedule is a host-side table, and the
whole sequence is replayed as a single CUDA graph launch.
"""
from __future__ import annotations

import os
from typing import Dict, List, Optional, Sequence

import torch

from . import native as nv
from .dino import DinoProgram, DinoWeights, native_pos_resize
from .persist import Collector, make_desc
from .plan import Plan, linear_desc, ptr, round_up
from .schedule import sde_coefficients, sde_schedule
from .unet import (FILM_ROWS, Mode, UnetBuffers, UnetWeights, build_cond_film, build_time_film, build_unet_eval)

SD = Dict[str, torch.Tensor]


def _pack_desc(src, src_ld, rows, cols, out, out_off, mode: Mode, out_cols_total, dst_c0, zero_to=0, act=nv.ACT_NONE):
    d = nv.PackDesc()
    d.src, d.src_ld, d.rows, d.cols, d.act = ptr(src), src_ld, rows, cols, act
    d.out, d.out_dtype, d.out_ld, d.dst_c0 = ptr(out, out_off), mode.dt, mode.ld(out_cols_total), dst_c0
    d.out_plane, d.zero_to = mode.plane(out_cols_total), zero_to
    return d


def _affine_desc(x, out, mins, maxs, rows, A, denorm, xpad=None, mode: Optional[Mode] = None, xpad_c=0, add=None, pad=1.4):
    d = nv.AffineDesc()
    d.x, d.out, d.mins, d.maxs, d.rows, d.A, d.denorm, d.pad = ptr(x), ptr(out), ptr(mins), ptr(maxs), rows, A, denorm, pad
    if xpad is not None:
        d.xpad, d.xpad_dtype, d.xpad_ld, d.xpad_plane = ptr(xpad), mode.dt, mode.ld(xpad_c), mode.plane(xpad_c)
    d.add = ptr(add)
    return d


class MlpWeights:
    """nn.Sequential(Linear, GELU, Linear, GELU, Linear) (bridge_controller.py:42-48) packed for the GEMM kernel."""

    def __init__(self, sd: SD, device, mode: Mode, idx=(0, 2, 4)):
        self.mode = mode
        self.w, self.b, self.dims = [], [], []
        for i in idx:
            w = sd[f"{i}.weight"].detach().to(device, torch.float32)
            n, k = w.shape
            n_pad, k_pad = round_up(n, 128 if n > 32 else 32), round_up(k, 64)
            wp = torch.zeros(n_pad, k_pad, device=device)
            wp[:n, :k] = w
            bp = torch.zeros(n_pad, device=device)
            bp[:n] = sd[f"{i}.bias"].detach().to(device, torch.float32)
            self.w.append(mode.pack_w(wp))
            self.b.append(bp)
            self.dims.append((n, k, n_pad, k_pad))

    def register(self, plan: Plan):
        for t in self.w + self.b:
            plan.reg(t)


def build_mlp(plan: Plan, W: MlpWeights, x_op: torch.Tensor, rows: int, out: torch.Tensor, tag: str, acts=None):
    """x_op: operand-dtype [rows][ld(k_pad0)] -> out fp32 [rows][n_last] (plain).  Hidden activations in operand dtype."""
    m = W.mode
    cur, cur_k = x_op, W.dims[0][3]
    nl = len(W.w)
    acts = acts or [nv.ACT_GELU] * (nl - 1) + [nv.ACT_NONE]
    for i in range(nl):
        n, k, n_pad, k_pad = W.dims[i]
        assert k_pad == cur_k, (k_pad, cur_k)
        last = i == nl - 1
        if last:
            dst, ldc, oplane = out, out.shape[-1], 0
        else:
            dst = plan.buf(f"{tag}.h{i}", (rows, m.ld(n_pad)), m.tdt)
            ldc, oplane = m.ld(n_pad), m.plane(n_pad)
        plan.add(linear_desc(a=cur, rows=rows, k=k_pad, a_ld=m.ld(k_pad), w=W.w[i], n=n if last else n_pad, n_pad=n_pad,
                             w_ld=W.w[i].shape[-1], out=dst, ldc=ldc, bias=W.b[i], act=acts[i], out_plane=oplane,
                             passes=m.passes, a_plane=m.plane(k_pad), w_plane=k_pad if m.precise else 0),
                 f"{tag}.linear{i}")
        cur, cur_k = dst, n_pad


class BridgeEngine:
    def __init__(self, *, dino: Optional[DinoWeights], enc_sd: SD, v_sd: SD, s_sd: SD, action_dim: int, state_dim: int,
                 force_dim: int, use_force: bool, B: int, T: int, H: int = 0, W: int = 0,
                 img_dtype: torch.dtype = torch.uint8, layout: int = nv.LAYOUT_BHWC, diffuse_step: int = 10,
                 beta_max: float = 0.03, device="cuda", precise: bool = False, resize=native_pos_resize,
                 hidden_dim: int = 256, inject_noise: bool = False, sde_type: str = "vs"):
        """v_sd: the drift net -- v_net for sde_type 'vs' (sde_vs, bridge_model.py:334-387), b_net for 'bs' (sde_bs, :281-332)."""
        if sde_type not in ("vs", "bs"):
            raise NotImplementedError(f"sde_type={sde_type!r}")
        self.sde_type = sde_type
        self.device = torch.device(device)
        self.mode = m = Mode(precise)
        self.B, self.T, self.A = B, T, action_dim
        self.n_steps, self.delta_t, ts = sde_schedule(diffuse_step)
        self.beta_max = float(beta_max)
        self.inject_noise = inject_noise
        self.use_force = use_force
        self.plan = p = Plan(self.device)
        dev = self.device
        A = action_dim

        # ---- inputs / outputs (fixed addresses: the program is replayed as a CUDA graph) ----
        self.state = p.buf("in.state", (B, state_dim), torch.float32)
        self.forces = p.buf("in.forces", (B, max(force_dim, 1)), torch.float32)
        self.vla = p.buf("in.vla", (B, T, A), torch.float32)
        self.stats = {k: p.buf(f"in.stats.{k}", (A,), torch.float32) for k in ("vla_mins", "vla_maxs", "action_mins", "action_maxs")}
        self.noise = p.buf("in.noise", (self.n_steps, B, T, A), torch.float32)
        self.seed = p.buf("in.seed", (1,), torch.int64)
        self.cond = p.buf("cond", (B, hidden_dim), torch.float32)
        self.x = p.buf("x", (B, T, A), torch.float32)
        self.out = p.buf("out", (B, T, A), torch.float32)
        self.ranges: Dict[str, tuple] = {}

        # ---- observation encoder ----
        start = len(p)
        self.dino_prog = None
        if dino is not None:
            self.dino_prog = DinoProgram(p, dino, 2, B, H, W, img_dtype, layout, resize)
            D = dino.D
        else:
            D = 0
        self.ranges["dino"] = (start, len(p))
        start = len(p)
        self.enc = None
        if enc_sd is not None:          # None: `cond` is an input (StochasticInterpolants.sample called directly)
            self.enc = MlpWeights(enc_sd, dev, m)
            self.enc.register(p)
            obs_dim = 2 * D + state_dim + (force_dim if use_force else 0)
            if self.enc.dims[0][1] != obs_dim:
                raise ValueError(f"state_encoder expects {self.enc.dims[0][1]} inputs, controller provides {obs_dim}")
            kpad = self.enc.dims[0][3]
            obs = p.buf("enc.obs", (B, m.ld(kpad)), m.tdt)
            c0 = 0
            if dino is not None:
                for c in range(2):
                    p.add(_pack_desc(self.dino_prog.feat[c], D, B, D, obs, 0, m, kpad, c * D), f"enc.cat.cam{c}")
                c0 = 2 * D
            p.add(_pack_desc(self.state, state_dim, B, state_dim, obs, 0, m, kpad, c0), "enc.cat.state")
            if use_force:
                p.add(_pack_desc(self.forces, force_dim, B, force_dim, obs, 0, m, kpad, c0 + state_dim), "enc.cat.force")
            build_mlp(p, self.enc, obs, B, self.cond, "enc")
        self.ranges["enc"] = (start, len(p))

        # ---- sampler ----
        self.unet = UnetWeights([v_sd, s_sd], A, dev, precise)
        self.unet.register(p)
        self.film_c = p.buf("film_c", (2, B, FILM_ROWS), torch.float32)
        self.film_t = p.buf("film_t", (2, self.n_steps, FILM_ROWS), torch.float32)
        self.bufs = UnetBuffers(p, self.unet, B, T)
        start = len(p)
        build_cond_film(p, self.unet, self.cond, B, self.film_c)
        self.ranges["film_c"] = (start, len(p))
        start = len(p)
        p.add(_affine_desc(self.vla, self.x, self.stats["vla_mins"], self.stats["vla_maxs"], B * T, A, 0, self.bufs.xpad, m,
                           self.unet.cin0), "normalize_actions(vla)")
        self.ranges["normalize"] = (start, len(p))
        # prior already normalised (StochasticInterpolants.sample called directly): x -> xpad only
        start = len(p)
        from .unet import xpad_desc
        p.add(xpad_desc(self.unet, self.x, B * T, self.bufs), "x_prior->xpad")
        self.ranges["xprior"] = (start, len(p))
        self._ts = ts
        # The sampling loop.  bf16 mode: ONE persistent kernel launch for all steps (csrc/vt_persist.cuh; the op list of one
        # evaluation + the Euler-Maruyama update, repeated n_steps times inside the kernel, ordered by per-sample-block
        # counters instead of kernel boundaries).  fp32 (split-tf32) mode, VT_PERSIST=0, or stepping through the trajectory
        # (`run_steps(k, 1)`): the multi-launch form, 36 implicit-GEMM launches + one update kernel per step.
        self.persistent = (not precise) and os.environ.get("VT_PERSIST", "1") != "0"
        self.step_ranges: List[tuple] = []
        self._step_plan: Optional[Plan] = None
        if self.persistent:
            start = len(p)
            col = Collector()
            build_unet_eval(col, self.unet, self.bufs, self.film_c, ptr(self.film_t, 0), self.n_steps * FILM_ROWS, tag="unet")
            coef = []
            for k in range(self.n_steps):
                ginv, dgg, eps, nscale = sde_coefficients(ts[k], self.delta_t, self.sde_type)
                coef.append((ginv, dgg, eps, self.delta_t, nscale))
            d = self._sde_desc(0)
            p.add(make_desc(p._reg, col.descs, sde=d, n_steps=self.n_steps, film_t_step=FILM_ROWS, coef=coef, noise_step=B * T * A,
                            sde_T=T, tags=col.tags), f"sde_vs: {self.n_steps} x [v_net + s_net evaluation, Euler-Maruyama] (persistent)")
            self.sampler_range = (start, len(p))
        else:
            self._build_steps(p)
            self.sampler_range = (self.step_ranges[0][0], self.step_ranges[-1][1])
        start = len(p)
        p.add(_affine_desc(self.x, self.out, self.stats["action_mins"], self.stats["action_maxs"], B * T, A, 1),
              "denormalize_actions(expert)")
        self.ranges["denormalize"] = (start, len(p))

        # ---- one-off program: time half of the FiLM tables for all steps ----
        self.setup = Plan(self.device)
        self.unet.register(self.setup)
        self.setup.reg(self.film_t)
        self._t_steps = self.setup.buf("t_steps", (self.n_steps,), torch.float32)
        self._t_steps.copy_(torch.cat(ts).to(self.device))
        build_time_film(self.setup, self.unet, self._t_steps, self.n_steps, None, self.film_t, tag="film_t")
        self._setup_done = False
        self._graphs: Dict[tuple, object] = {}
        self._noise_mode: Optional[bool] = None

    def _sde_desc(self, k: int) -> nv.SdeDesc:
        B, T, A, m = self.B, self.T, self.A, self.mode
        ginv, dgg, eps, nscale = sde_coefficients(self._ts[k], self.delta_t, self.sde_type)
        d = nv.SdeDesc()
        d.x, d.v, d.s = ptr(self.x), ptr(self.bufs.out), ptr(self.bufs.out, B * T * A)
        d.noise = ptr(self.noise, k * B * T * A) if self.inject_noise else None
        d.rows, d.A = B * T, A
        d.ginv, d.dgg, d.eps, d.dt, d.nscale, d.d = ginv, dgg, eps, self.delta_t, nscale, self.beta_max
        d.seed, d.seed_dev, d.step = 0, ptr(self.seed), k
        d.xpad, d.xpad_dtype, d.xpad_ld, d.xpad_plane = ptr(self.bufs.xpad), m.dt, m.ld(self.unet.cin0), m.plane(self.unet.cin0)
        return d

    def _build_steps(self, p: Plan) -> None:
        """The multi-launch form of the sampling loop appended to plan `p`: per step 36 launches + the update kernel."""
        self.step_ranges = []
        for k in range(self.n_steps):
            start = len(p)
            build_unet_eval(p, self.unet, self.bufs, self.film_c, ptr(self.film_t, k * FILM_ROWS), self.n_steps * FILM_ROWS,
                            tag=f"step{k}")
            p.add(self._sde_desc(k), f"step{k}.euler_maruyama")
            self.step_ranges.append((start, len(p)))

    def _steps_plan(self) -> Plan:
        """Multi-launch steps for callers that need the state after every step (sample(recod_traj=True)); built on first use,
        shares every buffer with the main plan."""
        if not self.persistent:
            return self.plan
        if self._step_plan is None:
            sp = Plan(self.device)
            sp._reg = self.plan._reg            # same tensors: addresses resolve identically
            self._build_steps(sp)
            self._step_plan = sp
        return self._step_plan

    # ---- weight refresh (same device addresses: programs and graphs stay valid) ----
    def refresh_unet(self, v_sd: SD, s_sd: SD) -> None:
        self.unet.refresh([v_sd, s_sd])
        self._setup_done = False        # the time half of the FiLM tables depends on the weights

    def refresh_enc(self, enc_sd: SD) -> None:
        new = MlpWeights(enc_sd, self.device, self.mode)
        for dst, src in zip(self.enc.w + self.enc.b, new.w + new.b):
            dst.copy_(src)

    # ---- execution ----
    def set_stats(self, stats: Dict[str, torch.Tensor]) -> None:
        for k, buf in self.stats.items():
            buf.copy_(torch.as_tensor(stats[k], dtype=torch.float32).to(self.device))

    def _ensure_setup(self):
        if not self._setup_done:
            self.setup.compile().run()
            self._setup_done = True

    def _run(self, a: int, b: int):
        self.plan.compile().run(a, b - a)

    def run_ranges(self, names: Sequence[str]):
        self._ensure_setup()
        for n in names:
            a, b = self.ranges[n]
            if b > a:
                self._run(a, b)

    def run_steps(self, first: int = 0, count: Optional[int] = None):
        self._ensure_setup()
        last = self.n_steps if count is None else first + count
        if self.persistent and first == 0 and last == self.n_steps:
            self._run(*self.sampler_range)
            return
        sp = self._steps_plan()
        sp.compile().run(self.step_ranges[first][0], self.step_ranges[last - 1][1] - self.step_ranges[first][0])

    def predict_range(self) -> tuple:
        a = self.ranges["dino"][0]
        return a, self.ranges["denormalize"][1]

    def run_predict(self, graph: bool = True):
        """dino -> enc -> film_c -> normalize -> steps -> denormalize (the `xprior` op is skipped)."""
        self._ensure_setup()
        prog = self.plan.compile()
        a0, a1 = self.ranges["dino"][0], self.ranges["normalize"][1]
        b0, b1 = self.sampler_range[0], self.ranges["denormalize"][1]
        if not graph:
            prog.run(a0, a1 - a0)
            prog.run(b0, b1 - b0)
            return
        # two graph segments would need two programs; the single skipped op (`xprior`) is a tiny pack that rewrites
        # xpad from x, which `normalize` has just written with identical contents, so replaying it is harmless.
        if "predict" not in self._graphs:
            prog.graph_build(a0, b1 - a0)
            self._graphs["predict"] = True
        prog.graph_launch()

    def num_launches(self) -> int:
        prog = self.plan.compile()
        a0, b1 = self.ranges["dino"][0], self.ranges["denormalize"][1]
        return prog.num_launches(a0, b1 - a0)
