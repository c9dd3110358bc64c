"""Scalar schedule of the velocity-score SDE sampler, for the one configuration every reference script selects
(interpolant 'linear', gamma '2^0.5*t(t-1)', epsilon '1-t', sde 'vs'; bridge_train.py:633-647).

All coefficients are batch-independent per step, so they are a small host-side table computed with the same fp32
tensor arithmetic and operation order as bridge/bridge_model.py:59-101,334-385 (literal 1.4142, not sqrt(2))."""
from __future__ import annotations

from typing import List, Tuple

import torch

T_MIN = 0.001            # bridge_model.py:43
GAMMA_INV_MAX = 200.0    # bridge_model.py:44

SUPPORTED = {"interpolant_type": "linear", "gamma_type": "2^0.5*t(t-1)", "epsilon_type": "1-t"}


def check_model_args(model_args: dict) -> None:
    """Unknown schedule names raise NotImplementedError like bridge_model.py:71,81,91,101,146."""
    for key, want in SUPPORTED.items():
        got = model_args.get(key, want)
        if got != want:
            raise NotImplementedError(f"{key}={got!r}: only {want!r} is implemented on the B200 path")
    if model_args.get("sde_type", "vs") not in ("vs", "bs"):
        raise NotImplementedError(f"sde_type={model_args.get('sde_type')!r}: 'vs' and 'bs' are implemented (bridge_model.py:268-275)")
    if model_args.get("net_type", "unet1D_si") != "unet1D_si":
        raise NotImplementedError(f"net_type={model_args.get('net_type')!r}")


def sde_schedule(diffuse_step: int) -> Tuple[int, float, List[torch.Tensor]]:
    """(n_steps, delta_t, [t_k]) as derived at bridge_model.py:269,335,347-348:
    delta_t = float(1.0/diffuse_step); n_steps = int(1.0/delta_t) (differs from diffuse_step for e.g. 93, 99)."""
    delta_t = float(1.0 / diffuse_step)
    n_steps = int(1.0 / delta_t)
    ts = []
    for k in range(1, n_steps + 1):
        t = torch.full((1,), k / n_steps).float()
        ts.append(torch.clip(t, T_MIN, 1.0 - T_MIN))
    return n_steps, delta_t, ts


def sde_coefficients(t: torch.Tensor, delta_t: float, sde_type: str = "vs") -> Tuple[float, float, float, float]:
    """(gamma_inv, dot_gamma*gamma, epsilon, delta_t*sqrt(2 epsilon)) at time t (fp32, reference operation order).
    sde_type 'bs' (sde_bs, bridge_model.py:281-332): the drift net's output is used as it is, i.e. the dot_gamma*gamma
    coefficient of the velocity form (b = v - dot_gamma gamma eps s, :369) is zero."""
    gamma = 1.4142 * t * (1 - t)
    dgamma = 1.4142 * (1 - 2 * t)
    ginv = torch.clamp(1 / (1.4142 * t * (1 - t) + 1e-4), 0.0, GAMMA_INV_MAX)
    eps = (1 - t) * 1.0
    nscale = delta_t * torch.sqrt(2 * eps)
    return float(ginv), (0.0 if sde_type == "bs" else float(dgamma * gamma)), float(eps), float(nscale)
