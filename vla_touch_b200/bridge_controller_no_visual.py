"""DiffusionController without the DinoV2 encoder (reference: bridge_controller_no_visual.py:16-140, the image-free ablation):
obs_cond = state_encoder(cat(state, force)); everything else -- normalise, the velocity/score SDE over the two U-Nets,
de-normalise, checkpoints -- is the visual controller's native program with the image stage left out."""
from __future__ import annotations

from .bridge_controller import DiffusionController as _VisualController


class DiffusionController(_VisualController):
    def __init__(self, state_dim=10, hidden_dim=256, image_model_path="facebook/dinov2-small", diffusion_steps=10, device="cuda",
                 model_args=None, use_force=True, force_dim=3, **kw):
        kw.pop("visual", None)
        super().__init__(state_dim=state_dim, hidden_dim=hidden_dim, image_model_path=image_model_path,
                         diffusion_steps=diffusion_steps, device=device, model_args=model_args, use_force=use_force,
                         force_dim=force_dim, visual=False, **kw)


def load_bridge_controller(path=None, use_force=True, **kw):
    """bridge_controller_no_visual.py:204-231 defaults."""
    model_args = {
        'interpolant_type': 'linear', 'gamma_type': '2^0.5*t(t-1)', 'epsilon_type': '1-t', 'prior_policy': 'vla',
        'beta_max': 0.03, 'sde_type': 'vs', 'action_dim': 10, 'obs_dim': 256, 'obs_horizon': 1, 'net_type': 'unet1D_si',
        'pretrain': False, 'context_frames': 2, 'horizon': 16,
    }
    controller = DiffusionController(state_dim=10, hidden_dim=256, diffusion_steps=10, model_args=model_args, force_dim=3,
                                     use_force=use_force, **kw)
    if path:
        controller.load(path)
    return controller
