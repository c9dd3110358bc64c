"""TEST INFRASTRUCTURE ONLY -- never imported by the product path.

Makes the *unmodified* reference hot-path modules under /root/reference importable in this
container so that golden vectors can be generated from the reference itself
(oracle/gen_golden.py) and the oracle restatement (oracle/vt_oracle.py) can be validated
against it.  /root/reference does not exist on the GPU box, so nothing that runs there may
call `import_reference()`.

Three third-party imports of the reference are absent from this image and are stubbed
(SURVEY.md 8c):
  * torch_ema.ExponentialMovingAverage  (bridge/bridge_model.py:10)  -> restated below from the
    published torch_ema 0.3 algorithm (decay warm-up min(decay,(1+n)/(10+n)), shadow update
    s -= (1-decay)(s-p), store/copy_to/restore context manager, state_dict keys).
  * diffusers.schedulers.scheduling_ddpm (conditional_unet_1D.py:4)  -> only used in __main__.
  * h5py (controller_dataset.py:10)                                   -> only used for dataset IO.
`Dinov2Model.from_pretrained` needs the network; it is replaced by a random-init
`Dinov2Model(Dinov2Config(...))` whose weights the caller then overwrites deterministically.
"""
from __future__ import annotations

import contextlib
import sys
import types

import torch

REF_ROOT = "/root/reference"


class ExponentialMovingAverage:
    """Restatement of torch_ema.ExponentialMovingAverage (v0.3, un-vendored dependency)."""

    def __init__(self, parameters, decay: float, use_num_updates: bool = True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        parameters = list(parameters)
        self.shadow_params = [p.clone().detach() for p in parameters]
        self.collected_params = None
        self._params_refs = parameters

    def _get_parameters(self, parameters):
        if parameters is None:
            return self._params_refs
        return list(parameters)

    def update(self, parameters=None):
        parameters = self._get_parameters(parameters)
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        one_minus_decay = 1.0 - decay
        with torch.no_grad():
            for s_param, param in zip(self.shadow_params, parameters):
                tmp = s_param - param
                tmp.mul_(one_minus_decay)
                s_param.sub_(tmp)

    def copy_to(self, parameters=None):
        parameters = self._get_parameters(parameters)
        for s_param, param in zip(self.shadow_params, parameters):
            param.data.copy_(s_param.data)

    def store(self, parameters=None):
        parameters = self._get_parameters(parameters)
        self.collected_params = [param.clone() for param in parameters]

    def restore(self, parameters=None):
        parameters = self._get_parameters(parameters)
        for c_param, param in zip(self.collected_params, parameters):
            param.data.copy_(c_param.data)

    @contextlib.contextmanager
    def average_parameters(self, parameters=None):
        parameters = self._get_parameters(parameters)
        self.store(parameters)
        self.copy_to(parameters)
        try:
            yield
        finally:
            self.restore(parameters)

    def to(self, device=None, dtype=None):
        self.shadow_params = [p.to(device=device, dtype=dtype) if p.is_floating_point() else p.to(device=device)
                              for p in self.shadow_params]
        return

    def state_dict(self):
        return {"decay": self.decay, "num_updates": self.num_updates,
                "shadow_params": self.shadow_params, "collected_params": self.collected_params}

    def load_state_dict(self, state_dict):
        self.decay = state_dict["decay"]
        self.num_updates = state_dict["num_updates"]
        self.shadow_params = [p.clone() for p in state_dict["shadow_params"]]
        self.collected_params = state_dict["collected_params"]


def _install_stubs():
    if "torch_ema" not in sys.modules:
        m = types.ModuleType("torch_ema")
        m.ExponentialMovingAverage = ExponentialMovingAverage
        sys.modules["torch_ema"] = m
    if "diffusers" not in sys.modules:
        d = types.ModuleType("diffusers")
        ds = types.ModuleType("diffusers.schedulers")
        dd = types.ModuleType("diffusers.schedulers.scheduling_ddpm")
        dd.DDPMScheduler = object
        d.schedulers = ds
        ds.scheduling_ddpm = dd
        sys.modules["diffusers"] = d
        sys.modules["diffusers.schedulers"] = ds
        sys.modules["diffusers.schedulers.scheduling_ddpm"] = dd
    if "h5py" not in sys.modules:
        sys.modules["h5py"] = types.ModuleType("h5py")
    try:
        import torchvision  # noqa: F401  (visual_encoder.py:5 imports torchvision.transforms)
    except Exception:  # pragma: no cover
        tv = types.ModuleType("torchvision")
        tv.transforms = types.ModuleType("torchvision.transforms")
        sys.modules["torchvision"] = tv
        sys.modules["torchvision.transforms"] = tv.transforms


DINO_CFG = {
    "facebook/dinov2-small": dict(hidden_size=384, num_attention_heads=6),
    "facebook/dinov2-base": dict(hidden_size=768, num_attention_heads=12),
}


def make_hf_dinov2(model_name: str, num_hidden_layers: int = 12):
    from transformers import Dinov2Config, Dinov2Model
    c = DINO_CFG[model_name]
    cfg = Dinov2Config(hidden_size=c["hidden_size"], num_attention_heads=c["num_attention_heads"],
                       num_hidden_layers=num_hidden_layers, image_size=518, patch_size=14,
                       layerscale_value=1.0, qkv_bias=True, mlp_ratio=4, hidden_act="gelu",
                       layer_norm_eps=1e-6)
    return Dinov2Model(cfg)


def import_reference(num_dino_layers: int = 12):
    """Returns a namespace with the reference's hot-path modules imported unmodified."""
    import os
    if not os.path.isdir(REF_ROOT):
        raise RuntimeError("/root/reference is not present (GPU box?) -- goldens must come from tests/golden/")
    _install_stubs()
    for p in (REF_ROOT + "/VLA", REF_ROOT + "/VLA/residual_controller"):
        if p not in sys.path:
            sys.path.insert(0, p)
    import transformers

    class _Patched:
        @staticmethod
        def from_pretrained(name, *a, **k):
            return make_hf_dinov2(name, _Patched.layers)
    _Patched.layers = num_dino_layers

    import visual_encoder  # reference
    visual_encoder.Dinov2Model = _Patched  # visual_encoder.py:27 -> no network
    import bridge_controller
    import controller_dataset
    import lstm_step_controller
    from bridge import bridge_model
    from residual_controller.bridge.networks import conditional_unet_1D, conditional_unet_1D_si
    ns = types.SimpleNamespace(
        visual_encoder=visual_encoder, bridge_controller=bridge_controller,
        controller_dataset=controller_dataset, lstm_step_controller=lstm_step_controller,
        bridge_model=bridge_model, conditional_unet_1D=conditional_unet_1D,
        conditional_unet_1D_si=conditional_unet_1D_si, set_dino_layers=lambda n: setattr(_Patched, "layers", n),
        transformers=transformers)
    return ns
