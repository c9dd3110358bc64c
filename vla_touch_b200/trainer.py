"""Training steps of the two controllers as the reference's trainer classes run them, data-parallel over torch.distributed:

* DiffusionControllerTrainer  (bridge_train.py:27-103 construction, :105-164 `_prepare_batch_for_diffusion`, :296-342 the step)
* LSTMControllerTrainer       (lstm_train.py:19-82 construction / `_prepare_batch`, :122-139 the step)

One step = batch -> device, normalise the action chunks, encode the observation (frozen DinoV2 on the native kernels, the
trainable 3-layer encoder with gradients), loss forward + backward as ONE native program, the single data-parallel exchange
(SUM all-reduce of the gradients, NCCL over NVLink / NVSwitch), fused AdamW + EMA + cosine LR.  What differs from running the
reference's loop on the drop-in classes (which also works: `get_loss(...).backward()` is differentiable) is only plumbing:

  - the U-Net gradients never leave the buffers the weight-gradient GEMMs wrote them to: they sit in one contiguous arena
    (LossBackwardProgram.grad_arena), the all-reduce runs IN PLACE on slices of it, launched per bucket on a side stream as
    soon as the backward program has passed the op that completes the bucket, and the optimizer kernel reads them in place;
  - AdamW, the EMA update (bridge_model.py:433, bridge_train.py:334) and the LR schedule are one kernel launch per step;
  - the per-step `.item()` calls (bridge_train.py:340-342) are left to the caller: losses are returned as device tensors.
"""
from __future__ import annotations

from typing import Dict, Optional

import torch

from .controller_dataset import normalize_actions
from .native import nvtx_range
from .optim import FusedAdamWEMA, arena_buckets


def _world(group=None):
    import torch.distributed as dist
    if dist.is_available() and dist.is_initialized():
        return dist.get_world_size(group)
    return 1


class DiffusionControllerTrainer:
    def __init__(self, controller, stats, learning_rate: float = 1e-4, weight_decay: float = 1e-6, device="cuda", group=None,
                 bucket_elems: int = 24 << 20, t_max: int = 100000):
        self.controller, self.device, self.group = controller, device, group
        self.lr, self.weight_decay, self.t_max = learning_rate, weight_decay, t_max
        self.bucket_elems = bucket_elems
        self.use_force = controller.use_force
        controller.stats = stats
        self.stats = {k: torch.as_tensor(v, dtype=torch.float32).to(device) for k, v in stats.items()}
        self.optimizer: Optional[FusedAdamWEMA] = None
        self._prog = None
        self._comm = None
        self.timing: Dict[str, float] = {}
        import os
        self._timing_on = os.environ.get("VT_TRAIN_TIMING") == "1"
        self._graph = os.environ.get("VT_TRAIN_GRAPH", "1") != "0"

    # ---- bridge_train.py:105-164 ----
    def _prepare_batch_for_diffusion(self, batch):
        ctx = self.controller.model_args.get('context_frames', 2)
        states, forces = batch['states'], batch['forces']
        current_state, current_forces = states[:, ctx - 1], forces[:, ctx - 1]
        if 'vla_act' in batch and 'expert_act' in batch:     # DeviceEpisodeStore.gather: normalised by the gather kernel already
            vla_n, exp_n = batch['vla_act'], batch['expert_act']
        else:
            vla_n = normalize_actions(batch['vla_actions'].to(self.device), self.stats, 'vla')
            exp_n = normalize_actions(batch['expert_actions'].to(self.device), self.stats, 'expert')
        cam1, cam2 = batch.get('images_cam1'), batch.get('images_cam2')
        if cam1 is not None and cam2 is not None:
            cam1, cam2 = cam1[:, -1], cam2[:, -1]
        feats = (batch['feat_cam1'], batch['feat_cam2']) if 'feat_cam1' in batch else None   # cached frozen-encoder features
        obs_cond = self.controller.encode_observation(current_state, cam1, cam2, current_forces, image_features=feats)
        return {'obs_cond': obs_cond, 'expert_act': exp_n, 'vla_act': vla_n, 'forces': forces[:, ctx:], 'current_force': current_forces}

    def _ensure(self, B: int, T: int):
        dm = self.controller.diffusion_model
        prog = dm.train_program(B, T)
        if prog is not self._prog:
            names = [n for n, _ in dm.net.named_parameters()]
            params = list(dm.net.parameters())
            src = prog.grad_sources()
            sources = {id(p): src[n] for n, p in zip(names, params)}
            enc = list(self.controller.state_encoder.parameters())
            if self.optimizer is None:
                self.optimizer = FusedAdamWEMA(params + enc, lr=self.lr, weight_decay=self.weight_decay, ema=dm.ema,
                                               ema_params=params, t_max=self.t_max, grad_sources=sources)
            else:
                self.optimizer.set_grad_sources(sources)
            flat, allocs = prog.grad_arena()
            self._arena = flat
            self._buckets = arena_buckets(allocs, flat.numel(), self.bucket_elems)
            self._prog = prog
        return prog

    def _tick(self, name: str, t0: float) -> float:
        """Developer timing (env VT_TRAIN_TIMING=1): host wall-clock per phase, accumulated in self.timing."""
        import time
        t1 = time.perf_counter()
        if self._timing_on:
            self.timing[name] = self.timing.get(name, 0.0) + (t1 - t0) * 1e3
        return t1

    def train_step(self, batch) -> Dict[str, torch.Tensor]:
        """bridge_train.py:296-342 for one minibatch; returns {'loss', 'v_loss', 's_loss', 'b_loss'} as device tensors."""
        import time
        import torch.distributed as dist
        t = time.perf_counter()
        self.controller.train()
        dm = self.controller.diffusion_model
        with nvtx_range("vt.train.prepare_batch"):
            bd = self._prepare_batch_for_diffusion(batch)
        t = self._tick("prepare_batch", t)
        obs = bd['obs_cond']
        x1, x0 = bd['expert_act'].float(), bd['vla_act'].float()
        B, T, A = x1.shape
        prog = self._ensure(B, T)
        self.optimizer.zero_grad()
        step = dm.step_override if dm.step_override is not None else torch.rand(B, device=self.device)
        z = dm.z_override if dm.z_override is not None else torch.randn_like(x1)
        t = self._tick("ensure+zero_grad+rng", t)
        dm.sync_train_program(prog)
        t = self._tick("sync_train_program", t)
        prog.set_inputs(x0, x1, obs.detach().float().flatten(1), step, z)
        world = _world(self.group)
        if world == 1:
            prog.run(graph=self._graph)                     # one CUDA graph launch instead of ~240 (VT_TRAIN_GRAPH=0: eager launches)
        else:
            native = prog.plan.compile()
            prog.runs = getattr(prog, "runs", 0) + 1
            main = torch.cuda.current_stream()
            if self._comm is None:
                self._comm = torch.cuda.Stream()
            works, op0 = [], 0
            for a, b, op_end in self._buckets:
                end = len(prog.plan) if op_end is None else op_end
                if end > op0:
                    native.run(op0, end - op0)
                    op0 = end
                ev = torch.cuda.Event()
                ev.record(main)
                self._comm.wait_event(ev)
                with torch.cuda.stream(self._comm):         # the collective waits for this bucket only, backward keeps running
                    works.append(dist.all_reduce(self._arena[a:b], op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            if op0 < len(prog.plan):
                native.run(op0, len(prog.plan) - op0)
        t = self._tick("program", t)
        out = prog.out.clone()
        if obs.requires_grad:                               # d loss / d obs_cond -> state encoder (torch autograd, 0.33 M parameters)
            obs.backward(prog.d_cond.view_as(obs))
        if world > 1:
            self.optimizer.gather_grads()
            works.append(dist.all_reduce(self.optimizer.grad_flat, op=dist.ReduceOp.SUM, group=self.group, async_op=True))
            for w in works:
                w.wait()
        t = self._tick("encoder backward + collective wait", t)
        with nvtx_range("vt.train.optimizer"):
            self.optimizer.step(grad_scale=1.0 / world)     # AdamW + EMA + cosine LR, one launch
        self._tick("optimizer", t)
        return {'loss': out[0], 'v_loss': out[1], 's_loss': out[2], 'b_loss': out[3]}


class LSTMControllerTrainer:
    def __init__(self, controller, stats, learning_rate: float = 1e-4, weight_decay: float = 1e-6, device="cuda", group=None,
                 t_max: int = 100000):
        self.controller, self.device, self.group = controller, device, group
        controller.stats = {k: torch.as_tensor(v, dtype=torch.float32).to(device) for k, v in stats.items()}
        self.optimizer = FusedAdamWEMA([p for m in controller.trainable_modules for p in m.parameters()], lr=learning_rate,
                                       weight_decay=weight_decay, t_max=t_max)

    # ---- lstm_train.py:57-82 ----
    def _prepare_batch(self, batch, context_frames: int = 2):
        c = self.controller
        current_state = batch['states'][:, context_frames - 1]
        forces = batch['forces'][:, context_frames - 1:-1] if c.use_force else None
        if 'vla_act' in batch and 'expert_act' in batch:     # DeviceEpisodeStore.gather: normalised by the gather kernel already
            vla_n, exp_n = batch['vla_act'], batch['expert_act']
        else:
            vla_n = normalize_actions(batch['vla_actions'].to(self.device), c.stats, 'vla')
            exp_n = normalize_actions(batch['expert_actions'].to(self.device), c.stats, 'expert')
        cam1 = batch['images_cam1'][:, -1] if 'images_cam1' in batch else None
        cam2 = batch['images_cam2'][:, -1] if 'images_cam2' in batch else None
        feats = (batch['feat_cam1'], batch['feat_cam2']) if 'feat_cam1' in batch else None   # cached frozen-encoder features
        obs_cond = c.encode_observation(current_state, cam1, cam2, image_features=feats)
        return {'obs_cond': obs_cond, 'expert_act': exp_n, 'vla_act': vla_n, 'forces': forces}

    def train_step(self, batch) -> torch.Tensor:
        """lstm_train.py:122-139 for one minibatch; returns the loss as a device tensor."""
        import torch.distributed as dist
        self.controller.train()
        bd = self._prepare_batch(batch)
        self.optimizer.zero_grad()
        loss = self.controller.get_loss(bd)
        loss.backward()
        world = _world(self.group)
        if world > 1:
            self.optimizer.gather_grads()
            dist.all_reduce(self.optimizer.grad_flat, op=dist.ReduceOp.SUM, group=self.group)
        self.optimizer.step(grad_scale=1.0 / world)
        return loss.detach()
