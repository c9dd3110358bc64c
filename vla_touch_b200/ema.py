"""ExponentialMovingAverage with the semantics and state-dict layout of the un-vendored `torch_ema` package the
reference uses (bridge/bridge_model.py:10,267,433-446; bridge_train.py:334): decay warm-up
min(decay, (1+n)/(10+n)), shadow parameters as a flat list in `parameters()` order, store/restore/copy_to and the
`average_parameters()` context manager."""
from __future__ import annotations

import contextlib
import weakref
from typing import Iterable, List, Optional

import torch


class ExponentialMovingAverage:
    def __init__(self, parameters: Iterable[torch.nn.Parameter], decay: float, use_num_updates: bool = True):
        if decay < 0.0 or decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.decay = decay
        self.num_updates = 0 if use_num_updates else None
        parameters = list(parameters)
        self.shadow_params: List[torch.Tensor] = [p.clone().detach() for p in parameters]
        self.collected_params: Optional[List[torch.Tensor]] = None
        self._params_refs = [weakref.ref(p) for p in parameters]
        self.version = 0          # bumped whenever the shadow weights change (engines re-pack lazily)
        self.param_writes = 0     # bumped whenever copy_to / restore overwrite the live parameters (p.data.copy_ does not
                                  # move the autograd version counters the loss programs watch)

    def _get_parameters(self, parameters):
        if parameters is None:
            parameters = [p() for p in self._params_refs]
            if any(p is None for p in parameters):
                raise ValueError("(One of) the parameters with which this ExponentialMovingAverage was initialized "
                                 "no longer exists (was garbage collected); please provide `parameters` explicitly.")
            return parameters
        parameters = list(parameters)
        if len(parameters) != len(self.shadow_params):
            raise ValueError("Number of parameters passed as argument is different from number of shadow parameters "
                             "maintained by this ExponentialMovingAverage")
        return parameters

    @torch.no_grad()
    def update(self, parameters=None) -> None:
        parameters = self._get_parameters(parameters)
        decay = self.decay
        if self.num_updates is not None:
            self.num_updates += 1
            decay = min(decay, (1 + self.num_updates) / (10 + self.num_updates))
        one_minus_decay = 1.0 - decay
        params = [p for p in parameters if p.requires_grad]
        shadows = [s for s, p in zip(self.shadow_params, parameters) if p.requires_grad]
        if params:
            diffs = torch._foreach_sub(shadows, params)
            torch._foreach_mul_(diffs, one_minus_decay)
            torch._foreach_sub_(shadows, diffs)
        self.version += 1

    @torch.no_grad()
    def copy_to(self, parameters=None) -> None:
        for s, p in zip(self.shadow_params, self._get_parameters(parameters)):
            p.data.copy_(s.data)
        self.param_writes += 1

    def store(self, parameters=None) -> None:
        # detached: a clone that carries a grad_fn cannot be deep-copied by load_state_dict / saved in a checkpoint
        self.collected_params = [p.detach().clone() for p in self._get_parameters(parameters)]

    @torch.no_grad()
    def restore(self, parameters=None) -> None:
        if self.collected_params is None:
            raise RuntimeError("This ExponentialMovingAverage has no `store()`ed weights to `restore()`")
        for c, p in zip(self.collected_params, self._get_parameters(parameters)):
            p.data.copy_(c.data)
        self.param_writes += 1

    @contextlib.contextmanager
    def average_parameters(self, parameters=None):
        parameters = self._get_parameters(parameters)
        self.store(parameters)
        self.copy_to(parameters)
        try:
            yield
        finally:
            self.restore(parameters)         # torch_ema keeps `collected_params` afterwards (it is part of state_dict())

    def to(self, device=None, dtype=None) -> None:
        self.shadow_params = [p.to(device=device, dtype=dtype) if p.is_floating_point() else p.to(device=device)
                              for p in self.shadow_params]
        if self.collected_params is not None:
            self.collected_params = [p.to(device=device, dtype=dtype) if p.is_floating_point() else p.to(device=device)
                                     for p in self.collected_params]
        self.version += 1

    def state_dict(self) -> dict:
        return {"decay": self.decay, "num_updates": self.num_updates, "shadow_params": self.shadow_params,
                "collected_params": self.collected_params}

    def load_state_dict(self, state_dict: dict) -> None:
        import copy
        state_dict = copy.deepcopy(state_dict)
        self.decay = state_dict["decay"]
        if self.decay < 0.0 or self.decay > 1.0:
            raise ValueError("Decay must be between 0 and 1")
        self.num_updates = state_dict["num_updates"]
        assert self.num_updates is None or isinstance(self.num_updates, int), "Invalid num_updates"
        shadow = state_dict["shadow_params"]
        assert isinstance(shadow, list) and all(isinstance(p, torch.Tensor) for p in shadow), "shadow_params must be a list of Tensors"
        if len(shadow) != len(self.shadow_params):
            raise ValueError("Tried to `load_state_dict()` with the wrong number of parameters in the saved state.")
        self.shadow_params = [s.to(device=old.device, dtype=old.dtype) for s, old in zip(shadow, self.shadow_params)]
        self.collected_params = state_dict.get("collected_params")
        self.version += 1
