"""TactileLSTMController with the reference's interface (lstm_step_controller.py:13-391): a residual 2-layer LSTM over
the VLA chunk and the per-step tactile force, conditioned on a DinoV2 + state observation code.

All linear maps (force encoder, LSTM input projections of ALL time steps at once, output head) are tcgen05 GEMMs; the
recurrence itself is one launch per layer of the LSTM_SEQ kernel (csrc/vt_lstm.cuh) for the whole sequence, or one
T=1 launch per control tick with the (h, c) state kept on the device (`predict`, :232-286)."""
from __future__ import annotations

import os
from typing import Dict, Optional

import torch
import torch.nn as nn
import torch.nn.functional as F

from . import native as nv
from .controller_dataset import denormalize_actions, normalize_actions
from .engine import MlpWeights, _affine_desc, _pack_desc, build_mlp
from .plan import Plan, linear_desc, ptr, round_up
from .unet import Mode
from .visual_encoder import DINOv2Encoder, forward_two_cameras

SD = Dict[str, torch.Tensor]


class LstmEngine:
    """vla_n [B,T,A], forces [B,T,F], obs_cond [B,H] -> out [B,T,A] (= vla_n + delta, optionally de-normalised)."""

    def __init__(self, mods: Dict[str, SD], A: int, Fd: int, H: int, L: int, B: int, T: int, device, precise: bool, denorm: bool):
        self.mode = m = Mode(precise)
        self.plan = p = Plan(device)
        self.B, self.T, self.A, self.H, self.L = B, T, A, H, L
        dev = torch.device(device)
        R = B * T
        f32 = torch.float32
        self.vla = p.buf("in.vla", (B, T, A), f32)
        self.forces = p.buf("in.forces", (B, T, Fd), f32)
        self.cond = p.buf("in.cond", (B, H), f32)
        self.h = p.buf("state.h", (L, B, H), f32)
        self.c = p.buf("state.c", (L, B, H), f32)
        self.out = p.buf("out", (B, T, A), f32)
        self.stats = {k: p.buf(f"in.stats.{k}", (A,), f32) for k in ("action_mins", "action_maxs")}
        # ---- weights ----
        self.fe = MlpWeights(mods["force_encoder"], dev, m, idx=(0, 2))
        self.fe.register(p)
        lstm = mods["lstm"]
        kin = H // 2 + A
        self.kin_pad = round_up(kin, 64)
        self.w_ih, self.b_sum, self.w_hh_t, self.w_hh_tc = [], [], [], []
        for l in range(L):
            w = lstm[f"weight_ih_l{l}"].detach().to(dev, f32)
            kp = self.kin_pad if l == 0 else H
            wp = torch.zeros(4 * H, kp, device=dev)
            wp[:, : w.shape[1]] = w
            self.w_ih.append(p.reg(m.pack_w(wp)))
            self.b_sum.append(p.reg((lstm[f"bias_ih_l{l}"] + lstm[f"bias_hh_l{l}"]).detach().to(dev, f32).contiguous()))
            self.w_hh_t.append(p.reg(lstm[f"weight_hh_l{l}"].detach().to(dev, f32).t().contiguous()))
            # tensor-core recurrence (csrc/vt_lstm_tc.cuh): W_hh rows regrouped per hidden unit (row = unit * 4 + gate), bf16
            self.w_hh_tc.append(None if precise else p.reg(
                lstm[f"weight_hh_l{l}"].detach().to(dev, f32).view(4, H, H).permute(1, 0, 2).reshape(4 * H, H).to(torch.bfloat16).contiguous()))
        head = mods["output_head"]
        self.head0 = MlpWeights({"0.weight": head["0.weight"], "0.bias": head["0.bias"]}, dev, m, idx=(0,))
        self.head0.register(p)
        self.ln_w = p.reg(head["1.weight"].detach().to(dev, f32).contiguous())
        self.ln_b = p.reg(head["1.bias"].detach().to(dev, f32).contiguous())
        self.head4 = MlpWeights({"0.weight": head["4.weight"], "0.bias": head["4.bias"]}, dev, m, idx=(0,))
        self.head4.register(p)
        # ---- program ----
        fpad = self.fe.dims[0][3]
        f_op = p.buf("f_op", (R, m.ld(fpad)), m.tdt)
        p.add(_pack_desc(self.forces, Fd, R, Fd, f_op, 0, m, fpad, 0), "lstm.force->operand")
        lin = p.buf("lstm_in", (R, m.ld(self.kin_pad)), m.tdt)
        fe_out = p.buf("fe_out", (R, H // 2), f32)
        build_mlp(p, self.fe, f_op, R, fe_out, "force_encoder", acts=[nv.ACT_GELU, nv.ACT_NONE])
        p.add(_pack_desc(fe_out, H // 2, R, H // 2, lin, 0, m, self.kin_pad, 0), "lstm.cat.force_code")
        p.add(_pack_desc(self.vla, A, R, A, lin, 0, m, self.kin_pad, H // 2), "lstm.cat.vla")
        head_in = p.buf("head_in", (R, m.ld(2 * H)), m.tdt)
        y_prev, y_ld, y_k = lin, m.ld(self.kin_pad), self.kin_pad
        for l in range(L):
            xw = p.buf(f"xw{l}", (R, 4 * H), f32)
            p.add(linear_desc(a=y_prev, rows=R, k=y_k, a_ld=y_ld, w=self.w_ih[l], n=4 * H, n_pad=4 * H,
                              w_ld=self.w_ih[l].shape[-1], out=xw, ldc=4 * H, bias=self.b_sum[l], passes=m.passes,
                              a_plane=m.plane(y_k), w_plane=y_k if m.precise else 0), f"lstm.l{l}.input_proj")
            last = l == L - 1
            y = head_in if last else p.buf(f"y{l}", (R, m.ld(H)), m.tdt)
            yk = 2 * H if last else H
            d = nv.LstmDesc()
            d.xw, d.w_hh, d.h, d.c = ptr(xw), ptr(self.w_hh_t[l]), ptr(self.h, l * B * H), ptr(self.c, l * B * H)
            d.y, d.y_dtype, d.y_ld, d.B, d.T, d.H = ptr(y), m.dt, m.ld(yk), B, T, H
            d.y_plane = m.plane(yk)
            if not precise and T > 1 and os.environ.get("VT_LSTM_TC", "1") != "0":        # whole-sequence passes start from a zero state (forward / predict_sequence, :196-204, 288-319)
                d.w_hh_tc, d.h_tc, d.zero_init = ptr(self.w_hh_tc[l]), ptr(p.buf(f"h_tc{l}", (B, T, H), torch.bfloat16)), 1
            p.add(d, f"lstm.l{l}.recurrence")
            y_prev, y_ld, y_k = y, m.ld(yk), yk
        d = _pack_desc(self.cond, H, R, H, head_in, 0, m, 2 * H, H)
        d.src_row_div = T
        p.add(d, "lstm.cat.obs_cond")
        z = p.buf("head.z", (R, H), f32)
        build_mlp(p, self.head0, head_in, R, z, "head0", acts=[nv.ACT_NONE])
        zn = p.buf("head.zn", (R, m.ld(H)), m.tdt)
        d = nv.LnDesc()
        d.x, d.in_ld, d.in_row_stride, d.rows, d.D = ptr(z), H, 1, R, H
        d.gamma, d.beta, d.eps = ptr(self.ln_w), ptr(self.ln_b), 1e-5
        d.out, d.out_dtype, d.out_ld, d.out_plane, d.act = ptr(zn), m.dt, m.ld(H), m.plane(H), nv.ACT_GELU
        p.add(d, "head.layernorm+gelu")
        delta = p.buf("head.delta", (R, A), f32)
        build_mlp(p, self.head4, zn, R, delta, "head4", acts=[nv.ACT_NONE])
        p.add(_affine_desc(self.vla, self.out, self.stats["action_mins"], self.stats["action_maxs"], R, A, 1 if denorm else 2,
                           add=delta), "vla+delta" + ("->denormalize" if denorm else ""))

    def run(self):
        self.plan.compile().run()


class TactileLSTMController:
    def __init__(self, state_dim=10, hidden_dim=256, num_layers=2, dropout=0.1, image_model_path="facebook/dinov2-small",
                 device="cuda", force_dim=3, use_force=True, *, image_state_dict=None, precise: bool = False,
                 allow_synthetic_dino: bool = False, image_num_layers: Optional[int] = None):
        if hidden_dim != 256:
            raise NotImplementedError("the LSTM_SEQ kernel is instantiated for hidden_dim=256 (the reference's only setting)")
        self.state_dim, self.hidden_dim, self.device = state_dim, hidden_dim, device
        self.force_dim, self.use_force, self.precise = force_dim, use_force, precise
        self.image_encoder = DINOv2Encoder(model_name=image_model_path, device=device, state_dict=image_state_dict,
                                           precise=precise, allow_synthetic_weights=allow_synthetic_dino,
                                           num_layers=image_num_layers)
        self.latent_obs_dim = self.image_encoder.hidden_size
        self.force_encoder = nn.Sequential(nn.Linear(force_dim, hidden_dim // 2), nn.GELU(),
                                           nn.Linear(hidden_dim // 2, hidden_dim // 2)).to(device)
        self.obs_dim = self.latent_obs_dim * 2 + self.state_dim
        self.obs_encoder = nn.Sequential(nn.Linear(self.obs_dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, hidden_dim),
                                         nn.GELU(), nn.Linear(hidden_dim, hidden_dim)).to(device)
        self.lstm_input_dim = hidden_dim // 2 + state_dim
        self.lstm = nn.LSTM(input_size=self.lstm_input_dim, hidden_size=hidden_dim, num_layers=num_layers, bidirectional=False,
                            batch_first=True, dropout=0.1 if num_layers > 1 else 0).to(device)
        self.output_head = nn.Sequential(nn.Linear(hidden_dim + hidden_dim, hidden_dim), nn.LayerNorm(hidden_dim), nn.GELU(),
                                         nn.Dropout(dropout), nn.Linear(hidden_dim, state_dim)).to(device)
        self.use_residual = True
        self.hidden_state = None
        self.cell_state = None
        self.stats = None
        self.trainable_modules = [self.obs_encoder, self.force_encoder, self.lstm, self.output_head]
        self._engines: Dict[tuple, tuple] = {}
        self._obs_engines: Dict[tuple, tuple] = {}

    def to(self, device):
        self.device = device
        for module in self.trainable_modules:
            module.to(device)
        return self

    def train(self, mode=True):
        for module in self.trainable_modules:
            module.train(mode)
        return self

    def eval(self):
        for module in self.trainable_modules:
            module.eval()
        return self

    def _version(self):
        return tuple(p._version for mod in self.trainable_modules for p in mod.parameters())

    def _engine(self, B, T, denorm) -> LstmEngine:
        key = (B, T, denorm)
        ent = self._engines.get(key)
        ver = self._version()
        if ent is None or ent[1] != ver:
            mods = {"force_encoder": self.force_encoder.state_dict(), "lstm": self.lstm.state_dict(),
                    "output_head": self.output_head.state_dict()}
            ent = (LstmEngine(mods, self.state_dim, self.force_dim, self.hidden_dim, self.lstm.num_layers, B, T, self.device,
                              self.precise, denorm), ver)
            self._engines[key] = ent
        return ent[0]

    @torch.no_grad()
    def encode_images(self, images_cam1, images_cam2):
        return forward_two_cameras(self.image_encoder, images_cam1, images_cam2)

    def _grad_wanted(self, modules) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for m in modules for p in m.parameters())

    def encode_observation(self, state, images_cam1=None, images_cam2=None, *, image_features=None):
        """obs_encoder(cat(cam1, cam2, state)) (:126-146).  image_features=(f1, f2): the frozen encoder's outputs, already computed
        (the DinoV2 feature cache of episode_store.DeviceEpisodeStore) -- the images are then not needed.  Under torch.no_grad() / inference_mode() (deployment, validation): the
        DinoV2 program + three tcgen05 GEMMs.  With autograd enabled and a trainable obs_encoder (lstm_train.py:70): the frozen
        DinoV2 features from the native kernels, then the 3-layer encoder as the torch module it is, so that
        `get_loss(...).backward()` trains it through d loss / d obs_cond."""
        if image_features is not None or self._grad_wanted([self.obs_encoder]):
            if image_features is not None:
                f1, f2 = (f.to(self.device).float() for f in image_features)
            else:
                with torch.no_grad():
                    f1, f2 = self.encode_images(images_cam1, images_cam2)
            st = state.to(self.device).float().reshape(f1.shape[0], -1)
            return self.obs_encoder(torch.cat((f1, f2, st), dim=-1))
        with torch.no_grad():
            return self._encode_observation_native(state, images_cam1, images_cam2)

    def _encode_observation_native(self, state, images_cam1, images_cam2):
        f1, f2 = self.encode_images(images_cam1, images_cam2)
        B = f1.shape[0]
        key = (B,)
        ent = self._obs_engines.get(key)
        ver = tuple(p._version for p in self.obs_encoder.parameters())
        if ent is None or ent[1] != ver:
            m = Mode(self.precise)
            plan = Plan(self.device)
            W = MlpWeights(self.obs_encoder.state_dict(), torch.device(self.device), m)
            W.register(plan)
            kpad = W.dims[0][3]
            D = self.latent_obs_dim
            feats = [plan.buf(f"f{c}", (B, D), torch.float32) for c in range(2)]
            st = plan.buf("state", (B, self.state_dim), torch.float32)
            obs = plan.buf("obs", (B, m.ld(kpad)), m.tdt)
            cond = plan.buf("cond", (B, self.hidden_dim), torch.float32)
            for c in range(2):
                plan.add(_pack_desc(feats[c], D, B, D, obs, 0, m, kpad, c * D), f"cat.cam{c}")
            plan.add(_pack_desc(st, self.state_dim, B, self.state_dim, obs, 0, m, kpad, 2 * D), "cat.state")
            build_mlp(plan, W, obs, B, cond, "obs_encoder")
            ent = ((plan, feats, st, cond), ver)
            self._obs_engines[key] = ent
        plan, feats, st, cond = ent[0]
        feats[0].copy_(f1)
        feats[1].copy_(f2)
        st.copy_(state.to(self.device))
        plan.compile().run()
        return cond.clone()

    @torch.no_grad()
    def encode_force(self, force):
        """force_encoder over [B,T,F] or [B,F] (:148-168): Linear(F,128) GELU Linear(128,128) as two tcgen05 GEMMs."""
        orig = force.shape
        f = force.to(self.device).float().reshape(-1, orig[-1])
        R = f.shape[0]
        ver = tuple(p._version for p in self.force_encoder.parameters())
        ent = self._force_engines.get(R) if hasattr(self, "_force_engines") else None
        if not hasattr(self, "_force_engines"):
            self._force_engines = {}
        if ent is None or ent[1] != ver:
            m = Mode(self.precise)
            plan = Plan(self.device)
            W = MlpWeights(self.force_encoder.state_dict(), torch.device(self.device), m, idx=(0, 2))
            W.register(plan)
            fpad = W.dims[0][3]
            fin = plan.buf("in.force", (R, orig[-1]), torch.float32)
            f_op = plan.buf("f_op", (R, m.ld(fpad)), m.tdt)
            out = plan.buf("out", (R, self.hidden_dim // 2), torch.float32)
            plan.add(_pack_desc(fin, orig[-1], R, orig[-1], f_op, 0, m, fpad, 0), "force->operand")
            build_mlp(plan, W, f_op, R, out, "force_encoder", acts=[nv.ACT_GELU, nv.ACT_NONE])
            ent = ((plan, fin, out), ver)
            self._force_engines[R] = ent
        plan, fin, out = ent[0]
        fin.copy_(f)
        plan.compile().run()
        return out.clone().reshape(*orig[:-1], -1)

    def _run(self, eng: LstmEngine, vla_n, obs_cond, forces):
        eng.vla.copy_(vla_n.reshape(eng.B, eng.T, -1))
        eng.forces.copy_(forces.reshape(eng.B, eng.T, -1))
        eng.cond.copy_(obs_cond)
        if self.stats is not None:
            for k, buf in eng.stats.items():
                buf.copy_(torch.as_tensor(self.stats[k], dtype=torch.float32).to(self.device))
        eng.run()

    @torch.no_grad()
    def forward(self, batch_dict):
        """Whole-sequence pass from a zero state (:170-213, eval semantics: dropout off) -> vla_act + delta [B,T,A]."""
        vla, cond, forces = batch_dict['vla_act'], batch_dict['obs_cond'], batch_dict['forces']
        B, T, _ = vla.shape
        eng = self._engine(B, T, False)
        eng.h.zero_()
        eng.c.zero_()
        self._run(eng, vla, cond, forces)
        return eng.out.clone()

    def reset_state(self, batch_size=1):
        eng = self._engine(batch_size, 1, True)
        eng.h.zero_()
        eng.c.zero_()
        self.hidden_state, self.cell_state = eng.h, eng.c

    @torch.no_grad()
    def predict(self, obs_cond, vla_action, force, initialize=False):
        """One control tick (:232-286): T=1 LSTM step carrying (hidden_state, cell_state) on the device."""
        self.eval()
        B = vla_action.shape[0]
        eng = self._engine(B, 1, True)
        if initialize or self.hidden_state is None or self.hidden_state is not eng.h:
            self.reset_state(B)
        self._run(eng, vla_action.to(self.device), obs_cond, force.to(self.device))
        return eng.out[:, 0].clone()

    @torch.no_grad()
    def predict_sequence(self, obs_cond, vla_actions, force_seq):
        """:288-319 -- step-by-step prediction over a chunk == one whole-sequence pass from a zero state, de-normalised."""
        B, T, _ = vla_actions.shape
        vla_n = normalize_actions(vla_actions.to(self.device), self.stats, 'vla')
        eng = self._engine(B, T, True)
        eng.h.zero_()
        eng.c.zero_()
        self._run(eng, vla_n, obs_cond, force_seq.to(self.device))
        return eng.out.clone()

    def get_loss(self, batch_dict, differentiable=None):
        """MSE(forward, expert_act) (:321-337).

        differentiable=None (default) follows torch's grad mode: under torch.no_grad() (validation, lstm_train.py:190-215) it is
        the loss VALUE from the inference program; with autograd enabled and trainable modules (the training step,
        lstm_train.py:120-130) the returned loss is differentiable: one native program computes the loss and, by an explicit
        backward (lstm_train.LstmLossBackwardProgram: BPTT recurrence kernel + dgrad / wgrad GEMMs), the gradients of the 18
        parameters of force_encoder / lstm / output_head and of obs_cond; `loss.backward()` hands them to the nn.Parameters and
        to the producer of `batch_dict['obs_cond']`.  bf16 operands.  In training mode (`controller.train()`) the reference's
        dropout (0.1 between the LSTM layers and in the head) is applied with in-kernel Philox masks, a new pair every call; in
        eval mode the gradients are those of the deterministic network."""
        if differentiable is None:
            differentiable = self._grad_wanted([self.force_encoder, self.lstm, self.output_head]) or \
                (torch.is_grad_enabled() and batch_dict['obs_cond'].requires_grad)
        if not differentiable:
            with torch.no_grad():
                return F.mse_loss(self.forward(batch_dict), batch_dict['expert_act'].to(self.device))
        from .lstm_train import LstmLossBackwardProgram
        vla = batch_dict['vla_act'].to(self.device).float()
        B, T, A = vla.shape
        mods_of = lambda: {"force_encoder": self.force_encoder.state_dict(), "lstm": self.lstm.state_dict(),
                           "output_head": self.output_head.state_dict()}
        ver = self._version()
        # training mode (controller.train(), lstm_train.py:118): the reference's nn.LSTM(dropout=0.1) / nn.Dropout(0.1) are active
        p_drop = (float(self.lstm.dropout), float(self.output_head[3].p)) if self.lstm.training else (0.0, 0.0)
        if not hasattr(self, "_train_programs"):
            self._train_programs, self._drop_seed = {}, 0
        ent = self._train_programs.get((B, T, p_drop))
        if ent is None:
            ent = [LstmLossBackwardProgram(mods_of(), A, batch_dict['forces'].shape[-1], B, T, self.device, dropout=p_drop), ver]
            self._train_programs[(B, T, p_drop)] = ent
        elif ent[1] != ver:
            ent[0].refresh(mods_of())
            ent[1] = ver
        if max(p_drop) > 0:
            self._drop_seed += 1
            ent[0].seed.fill_(self._drop_seed)                   # fresh Philox masks every step, no re-encoding of the program
        params, names = [], []
        for mname, mod in (("force_encoder", self.force_encoder), ("lstm", self.lstm), ("output_head", self.output_head)):
            for n, p_ in mod.named_parameters():
                params.append(p_)
                names.append(f"{mname}.{n}")
        return _LstmLossFn.apply(ent[0], names, vla, batch_dict['forces'].to(self.device).float(),
                                 batch_dict['expert_act'].to(self.device).float(), batch_dict['obs_cond'].to(self.device).float(), *params)

    def save(self, path):
        state_dict = {'stats': self.stats, 'model_args': getattr(self, 'model_args', None),
                      'modules': {'obs_encoder': self.obs_encoder.state_dict(), 'force_encoder': self.force_encoder.state_dict(),
                                  'lstm': self.lstm.state_dict(), 'output_head': self.output_head.state_dict()}}
        torch.save(state_dict, f"{path}/tactile_controller.pt")

    def load(self, path):
        checkpoint = torch.load(f"{path}/tactile_controller.pt", map_location=self.device, weights_only=False)
        modules = checkpoint['modules']
        self.obs_encoder.load_state_dict(modules['obs_encoder'])
        self.force_encoder.load_state_dict(modules['force_encoder'])
        self.lstm.load_state_dict(modules['lstm'])
        self.output_head.load_state_dict(modules['output_head'])
        self.stats = {key: torch.as_tensor(value, dtype=torch.float32).to(self.device) for key, value in checkpoint['stats'].items()}
        self.model_args = checkpoint.get('model_args', None)


def load_lstm_controller(path=None, state_dim=10, force_dim=3, device="cuda", **kw):
    """The reference's loader references undefined globals (lstm_step_controller.py:382-391); this one takes them as args."""
    controller = TactileLSTMController(state_dim=state_dim, hidden_dim=256, num_layers=2, dropout=0.1, device=device,
                                       force_dim=force_dim, **kw)
    if path:
        controller.load(path)
    return controller


class _LstmLossFn(torch.autograd.Function):
    """Scalar MSE loss whose gradients were computed eagerly by the native program in forward()."""

    @staticmethod
    def forward(ctx, prog, names, vla, forces, expert, cond, *params):
        prog.set_inputs(vla, forces, cond, expert)
        prog.run()
        ctx.prog, ctx.run_id = prog, prog.runs
        ctx.save_for_backward(prog.d_cond, *[prog.grads[n].reshape(p.shape) for n, p in zip(names, params)])
        return prog.loss_tensor().clone()

    @staticmethod
    def backward(ctx, gout):
        if ctx.prog.runs != ctx.run_id:
            raise RuntimeError("get_loss() was called again before this loss.backward(): the gradient buffers were overwritten")
        d_cond, *grads = ctx.saved_tensors
        return (None, None, None, None, None, gout * d_cond) + tuple(torch._foreach_mul(list(grads), gout))
