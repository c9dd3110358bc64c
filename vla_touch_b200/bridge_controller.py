"""DiffusionController with the reference's interface (bridge_controller.py:10-273): refines an RDT action chunk with
a stochastic-interpolant sampler conditioned on two camera images, the robot state and the tactile force.

`predict()` = DinoV2 x2 -> state_encoder -> normalise -> n-step velocity/score SDE over two conditional 1-D U-Nets ->
de-normalise, executed as ONE CUDA-graph launch of hand-written sm_100a kernels (vla_touch_b200.engine.BridgeEngine).
`state_encoder` / `force_decoder` stay `nn.Sequential` objects and the U-Nets stay parameter trees with the reference's
state-dict keys, so optimizers and the `controller.pt` / `bridge_model.pt` checkpoint files interchange with the reference.
"""
from __future__ import annotations

from typing import Dict, Optional

import numpy as np
import torch
import torch.nn as nn

from . import native as nv
from .bridge.bridge_model import StochasticInterpolants
from .controller_dataset import denormalize_actions, normalize_actions  # noqa: F401  (re-exported like the reference)
from .engine import BridgeEngine
from .visual_encoder import DINOv2Encoder, forward_two_cameras, prepare_images


class DiffusionController:
    def __init__(self, state_dim=10, hidden_dim=256, image_model_path="facebook/dinov2-small", diffusion_steps=10,
                 device="cuda", model_args=None, use_force=True, force_dim=3, *, image_state_dict=None,
                 precise: bool = False, allow_synthetic_dino: bool = False, image_num_layers: Optional[int] = None,
                 visual: bool = True):
        self.state_dim = state_dim
        self.hidden_dim = hidden_dim
        self.device = device
        self.diffusion_steps = diffusion_steps
        self.precise = precise
        if visual:
            self.image_encoder = DINOv2Encoder(model_name=image_model_path, device=device, state_dict=image_state_dict,
                                               precise=precise, allow_synthetic_weights=allow_synthetic_dino,
                                               num_layers=image_num_layers)
            self.latent_obs_dim = self.image_encoder.hidden_size
        else:                                   # bridge_controller_no_visual.py:32-33: obs = cat(state, force)
            self.image_encoder = None
            self.latent_obs_dim = 0
        self.use_force = use_force
        self.force_dim = force_dim
        self.model_args = model_args
        self.stats = None
        self.obs_dim = self.latent_obs_dim * 2 + self.state_dim + (self.force_dim if use_force else 0)
        self.state_encoder = nn.Sequential(
            nn.Linear(self.obs_dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, hidden_dim), nn.GELU(),
            nn.Linear(hidden_dim, hidden_dim)).to(device)
        if self.use_force:
            self.force_decoder = nn.Sequential(
                nn.Linear(hidden_dim, hidden_dim), nn.GELU(), nn.Linear(hidden_dim, int(hidden_dim / 2)), nn.GELU(),
                nn.Linear(int(hidden_dim / 2), force_dim)).to(device)
        self.diffusion_model = StochasticInterpolants(precise=precise)
        if self.model_args:
            self.diffusion_model.load_model(self.model_args, device)
        self._engines: Dict[tuple, BridgeEngine] = {}
        self._versions: Dict[tuple, tuple] = {}
        self.noise_override: Optional[torch.Tensor] = None   # [n_steps,B,T,A] injected N(0,1) draws (parity tests)
        self._seed = int(torch.initial_seed()) & 0x7FFFFFFF       # Philox base seed follows torch.manual_seed (new stream per call)
        self.to(device)

    # ---- module-ish plumbing ----
    def to(self, device):
        self.device = device
        self.state_encoder.to(device)
        if self.use_force:
            self.force_decoder.to(device)
        return self

    def train(self):
        self.state_encoder.train()
        self.diffusion_model.train()
        if self.use_force:
            self.force_decoder.train()
        return self

    def eval(self):
        self.state_encoder.eval()
        self.diffusion_model.eval()
        if self.use_force:
            self.force_decoder.eval()
        return self

    # ---- engine cache ----
    def _enc_version(self):
        return tuple(p._version for p in self.state_encoder.parameters())

    def _engine(self, B, T, H, W, img_dtype, layout, inject) -> BridgeEngine:
        dm = self.diffusion_model
        if dm.net is None:
            raise RuntimeError("DiffusionController needs model_args (diffusion model not initialised)")
        key = (B, T, H, W, img_dtype, layout, int(self.diffusion_steps), inject)
        eng = self._engines.get(key)
        ver = (dm.ema.version, self._enc_version())
        if eng is None:
            v_sd, s_sd = dm.ema_state_dicts()
            eng = BridgeEngine(dino=self.image_encoder.weights() if self.image_encoder is not None else None, enc_sd=self.state_encoder.state_dict(), v_sd=v_sd, s_sd=s_sd,
                               action_dim=dm.net.input_dim, state_dim=self.state_dim, force_dim=self.force_dim,
                               use_force=self.use_force, B=B, T=T, H=H, W=W, img_dtype=img_dtype, layout=layout,
                               diffuse_step=self.diffusion_steps, beta_max=dm.d, device=self.device, precise=self.precise,
                               hidden_dim=self.hidden_dim, inject_noise=inject, sde_type=dm.sde_type)
            self._engines[key] = eng
            self._versions[key] = ver
        elif self._versions[key] != ver:
            if self._versions[key][0] != ver[0]:
                eng.refresh_unet(*dm.ema_state_dicts())
            if self._versions[key][1] != ver[1]:
                eng.refresh_enc(self.state_encoder.state_dict())
            self._versions[key] = ver
        return eng

    def _upload_images(self, eng: BridgeEngine, img1, img2):
        """Pinned host images (the deployment call shape) are uploaded on a side stream into one of two landing buffers and
        copied device-to-device into the program's fixed input buffers, so the host->device transfer of call i+1 overlaps
        the kernels of call i (the caller's stream semantics are unchanged: everything it sees is ordered on its stream)."""
        dst = eng.dino_prog.img
        if not (img1.device.type == "cpu" and img1.is_pinned() and img2.is_pinned()):
            dst[0].copy_(img1, non_blocking=True)
            dst[1].copy_(img2, non_blocking=True)
            return
        st = getattr(eng, "_upload", None)
        if st is None:
            st = eng._upload = {"stream": torch.cuda.Stream(device=self.device), "k": 0,
                                "land": [[torch.empty_like(d) for d in dst] for _ in range(2)],
                                "free": [torch.cuda.Event(), torch.cuda.Event()]}
            main = torch.cuda.current_stream(self.device)
            for ev in st["free"]:
                ev.record(main)
        k = st["k"] = st["k"] ^ 1
        main, side = torch.cuda.current_stream(self.device), st["stream"]
        side.wait_event(st["free"][k])             # the device-to-device copy of two calls ago has drained this buffer
        with torch.cuda.stream(side):
            st["land"][k][0].copy_(img1, non_blocking=True)
            st["land"][k][1].copy_(img2, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(side)
        main.wait_event(ready)
        dst[0].copy_(st["land"][k][0], non_blocking=True)
        dst[1].copy_(st["land"][k][1], non_blocking=True)
        st["free"][k].record(main)

    def _load_inputs(self, eng: BridgeEngine, state, img1, img2, forces):
        if self.image_encoder is not None:
            self._upload_images(eng, img1, img2)
        eng.state.copy_(state.reshape(eng.B, -1), non_blocking=True)
        if self.use_force:
            if forces is None:
                raise ValueError("use_force=True but forces is None")
            eng.forces.copy_(forces.reshape(eng.B, -1), non_blocking=True)

    def _prep(self, state, images_cam1, images_cam2, T, inject=False):
        if self.image_encoder is None:          # no-visual ablation: the images, if any are passed, are ignored like the reference does
            return self._engine(state.shape[0], T, 0, 0, torch.uint8, nv.LAYOUT_BHWC, inject), None, None
        if images_cam1 is None or images_cam2 is None:
            raise ValueError("this controller was built with the DinoV2 encoder: images_cam1 / images_cam2 are required "
                             "(bridge_controller_no_visual.DiffusionController is the image-free variant)")
        img1, layout = prepare_images(images_cam1, self.device, keep_pinned_host=True)
        img2, layout2 = prepare_images(images_cam2, self.device, keep_pinned_host=True)
        if layout != layout2 or img1.shape != img2.shape or img1.dtype != img2.dtype:
            raise ValueError("both cameras must share shape, dtype and layout")
        B = img1.shape[0]
        H, W = (img1.shape[1], img1.shape[2]) if layout == nv.LAYOUT_BHWC else (img1.shape[2], img1.shape[3])
        if state.shape[0] != B:
            raise ValueError(f"state batch {state.shape[0]} != image batch {B}")
        return self._engine(B, T, H, W, img1.dtype, layout, inject), img1, img2

    # ---- reference API ----
    @torch.no_grad()
    def encode_images(self, images_cam1, images_cam2):
        if images_cam1 is None or images_cam2 is None or self.image_encoder is None:
            return None
        return forward_two_cameras(self.image_encoder, images_cam1, images_cam2)

    def _trains_encoder(self) -> bool:
        return torch.is_grad_enabled() and any(p.requires_grad for p in self.state_encoder.parameters())

    def encode_observation(self, state, images_cam1=None, images_cam2=None, forces=None, *, differentiable=None, image_features=None):
        """-> obs_cond [B, hidden_dim] (bridge_controller.py:112-134).

        image_features=(f1, f2): the frozen encoder's outputs for this batch, already computed (the DinoV2 feature cache of
        episode_store.DeviceEpisodeStore); the images are then not needed and the state encoder runs as the torch module.

        differentiable=None (default) follows torch's grad mode like a plain nn.Module call would: under torch.no_grad() /
        inference_mode() (predict, validation, deployment) it is the inference path, one native program (DinoV2 x 2 + the
        state-encoder GEMMs); with autograd enabled and a trainable state encoder (the training loop, bridge_train.py:151)
        the frozen DinoV2 features come from the native kernels and the 3-layer state encoder (0.5 MFLOP per sample,
        < 0.01 % of the step) is applied as the torch module it is, so that `get_loss(...).backward()` reaches its
        parameters through d loss / d obs_cond.  differentiable="native": the same with the encoder's forward and backward
        as native programs behind a torch.autograd.Function (mlp_train.py)."""
        if differentiable is None:
            differentiable = self._trains_encoder()
        if image_features is not None and self.image_encoder is None:
            raise ValueError("image_features passed to a controller without an image encoder")
        if differentiable or image_features is not None:
            B = state.shape[0]
            st = state.to(self.device).float().reshape(B, -1)
            if self.use_force:
                if forces is None:
                    raise ValueError("use_force=True but forces is None")
                st = torch.cat((st, forces.to(self.device).float().reshape(B, -1)), dim=-1)
            if image_features is not None:
                f1, f2 = image_features
                x = torch.cat((f1.to(self.device).float(), f2.to(self.device).float(), st), dim=-1)
            elif self.image_encoder is not None:
                with torch.no_grad():
                    f1, f2 = self.encode_images(images_cam1, images_cam2)
                x = torch.cat((f1, f2, st), dim=-1)
            else:
                x = st
            if differentiable == "native":         # forward + backward of the encoder as native programs (mlp_train.py)
                from .mlp_train import encoder_forward
                if not hasattr(self, "_enc_train_cache"):
                    self._enc_train_cache = {}
                return encoder_forward(self.state_encoder, self._enc_train_cache, x)
            return self.state_encoder(x)
        with torch.no_grad():
            return self._encode_observation_native(state, images_cam1, images_cam2, forces)

    def _encode_observation_native(self, state, images_cam1, images_cam2, forces):
        horizon = (self.model_args or {}).get('horizon', 16)
        eng, img1, img2 = self._prep(state, images_cam1, images_cam2, horizon)
        self._load_inputs(eng, state, img1, img2, forces)
        eng.run_ranges(["dino", "enc"])
        return eng.cond.clone()

    @torch.no_grad()
    def predict(self, state, vla_actions, images_cam1=None, images_cam2=None, forces=None):
        """state [B,A], vla_actions [B,T,A] un-normalised, images per camera, forces [B,F] -> refined [B,T,A]
        (bridge_controller.py:149-182)."""
        self.eval()
        if self.stats is None:
            raise RuntimeError("controller.stats is not set (load a checkpoint or assign the normalisation stats)")
        B, T, A = vla_actions.shape
        inject = self.noise_override is not None
        with nv.nvtx_range("vt.predict.inputs"):
            eng, img1, img2 = self._prep(state, images_cam1, images_cam2, T, inject)
            self._load_inputs(eng, state, img1, img2, forces)
            eng.vla.copy_(vla_actions, non_blocking=True)
            eng.set_stats(self.stats)
            if inject:
                eng.noise.copy_(self.noise_override)
            else:
                self._seed += 1
                eng.seed.fill_(self._seed)
        with nv.nvtx_range("vt.predict.program"):
            eng.run_predict(graph=True)
            return eng.out.clone()

    def get_reconstruction_loss(self, batch_data):
        target_force = batch_data['current_force'].to(self.device)
        obs_cond = batch_data['obs_cond'].to(self.device)
        return torch.nn.functional.mse_loss(self.force_decoder(obs_cond), target_force)

    def save(self, path):
        state_dict = {'state_encoder': self.state_encoder.state_dict(), 'model_args': self.model_args, 'stats': self.stats}
        if self.use_force:
            state_dict['force_decoder'] = self.force_decoder.state_dict()
        torch.save(state_dict, f"{path}/controller.pt")
        self.diffusion_model.save_model(path)

    def load(self, path):
        checkpoint = torch.load(f"{path}/controller.pt", map_location=self.device, weights_only=False)
        self.state_encoder.load_state_dict(checkpoint['state_encoder'])
        if self.use_force:
            self.force_decoder.load_state_dict(checkpoint['force_decoder'])
        self.model_args = checkpoint['model_args']
        self.stats = {key: torch.as_tensor(np.asarray(value.cpu() if torch.is_tensor(value) else value), dtype=torch.float32).to(self.device)
                      for key, value in checkpoint['stats'].items()}
        self.diffusion_model.load_model({**self.model_args, 'ckpt_path': path, 'pretrain': True}, self.device)
        self._engines.clear()


def load_bridge_controller(path=None, use_force=True, **kw):
    """Reference defaults (bridge_controller.py:246-273); tolerates the call-site drift `load_bridge_controller(use_force=True)`
    and `load_bridge_controller(path)`: a `path` is loaded after construction."""
    if isinstance(path, bool):
        path, use_force = None, path
    model_args = {
        'interpolant_type': 'linear', 'gamma_type': '2^0.5*t(t-1)', 'epsilon_type': '1-t', 'prior_policy': 'vla',
        'beta_max': 0.03, 'sde_type': 'vs', 'action_dim': 10, 'obs_dim': 256, 'obs_horizon': 1, 'net_type': 'unet1D_si',
        'pretrain': False, 'context_frames': 2, 'horizon': 16,
    }
    controller = DiffusionController(state_dim=10, hidden_dim=256, image_model_path="facebook/dinov2-small", diffusion_steps=10,
                                     model_args=model_args, force_dim=3, use_force=use_force, **kw)
    if path:
        controller.load(path)
    return controller
