"""Training-side pieces of the bridge / LSTM trainers (reference: bridge_train.py:296-342, lstm_train.py:122-139):

* FusedAdamWEMA -- torch.optim.AdamW(lr, weight_decay, default betas/eps) + torch_ema update + CosineAnnealingLR as ONE
  multi-tensor kernel launch per step (csrc/vt_elem.cuh adamw_ema_kernel) instead of ~1300 per-tensor launches.
* allreduce_gradients -- the single data-parallel exchange step of the path: bucketed SUM all-reduce of the gradients over
  torch.distributed (NCCL over NVLink 5 / NVSwitch on the B200 box, gloo in the CPU tests); the 1/world scaling is folded
  into the optimizer kernel (`grad_scale`).

The backward kernels that would produce the gradients are not built yet; both pieces are exercised
with externally supplied gradients."""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Optional, Sequence

import torch

from . import native as nv

CHUNK = 65536


def _bump_version(p: torch.Tensor) -> None:
    try:
        torch._C._autograd._unsafe_set_version_counter(p, p._version + 1)
    except Exception:
        with torch.no_grad():
            p.add_(0)


class FusedAdamWEMA:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-6, ema=None, ema_params: Optional[Sequence[torch.nn.Parameter]] = None,
                 t_max: Optional[int] = None, eta_min: Optional[float] = None):
        """`ema`: a vla_touch_b200.ema.ExponentialMovingAverage whose shadow list lines up with `ema_params` (a subset of
        `params`, e.g. the U-Net parameters but not the state encoder, bridge_train.py:50-57, bridge_model.py:433)."""
        self.params: List[torch.nn.Parameter] = [p for p in params]
        if not self.params:
            raise ValueError("optimizer got an empty parameter list")
        dev = self.params[0].device
        if dev.type != "cuda":
            raise nv.NativeError("FusedAdamWEMA runs on a CUDA (B200) device only")
        self.base_lr, self.lr, self.betas, self.eps, self.weight_decay = lr, lr, betas, eps, weight_decay
        self.t_max, self.eta_min = t_max, (lr / 10 if eta_min is None else eta_min)
        self.ema = ema
        self.step_count = 0
        self.m = [torch.zeros_like(p, dtype=torch.float32) for p in self.params]
        self.v = [torch.zeros_like(p, dtype=torch.float32) for p in self.params]
        shadow = {}
        if ema is not None:
            ema_params = list(ema_params if ema_params is not None else self.params)
            if len(ema_params) != len(ema.shadow_params):
                raise ValueError("ema_params must line up with ema.shadow_params")
            shadow = {id(p): s for p, s in zip(ema_params, ema.shadow_params)}
        self._shadow = [shadow.get(id(p)) for p in self.params]
        self._grads = [torch.zeros_like(p, dtype=torch.float32) for p in self.params]   # fixed gradient buffers
        for p, g in zip(self.params, self._grads):
            p.grad = g
        recs = (nv.OptTensor * len(self.params))()
        chunks = []
        for i, p in enumerate(self.params):
            assert p.dtype == torch.float32 and p.is_contiguous()
            s = self._shadow[i]
            recs[i] = nv.OptTensor(p.data_ptr(), self._grads[i].data_ptr(), self.m[i].data_ptr(), self.v[i].data_ptr(),
                                   s.data_ptr() if s is not None else None, p.numel())
            chunks += [(i, off) for off in range(0, p.numel(), CHUNK)]
        self._recs = torch.frombuffer(bytearray(bytes(recs)), dtype=torch.uint8).to(dev)
        self._chunks = torch.tensor(chunks, dtype=torch.int64).to(dev)
        self.n_chunks = len(chunks)

    def zero_grad(self, set_to_none: bool = False) -> None:
        for p, g in zip(self.params, self._grads):
            g.zero_()
            p.grad = g

    def current_lr(self) -> float:
        if self.t_max is None:
            return self.base_lr
        t = min(self.step_count, self.t_max)      # CosineAnnealingLR(T_max, eta_min), stepped once per optimizer step
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * t / self.t_max)) / 2

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0) -> None:
        for p, g in zip(self.params, self._grads):          # a backward pass may have re-bound .grad
            if p.grad is not None and p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g
        self.lr = self.current_lr()
        self.step_count += 1
        t = self.step_count
        ema_decay = 1.0
        if self.ema is not None:
            decay = self.ema.decay
            if self.ema.num_updates is not None:
                self.ema.num_updates += 1
                decay = min(decay, (1 + self.ema.num_updates) / (10 + self.ema.num_updates))
            ema_decay = decay
            self.ema.version += 1
        d = nv.AdamwDesc()
        d.tensors, d.chunks, d.n_chunks, d.chunk_elems = self._recs.data_ptr(), self._chunks.data_ptr(), self.n_chunks, CHUNK
        d.lr, d.beta1, d.beta2, d.eps, d.weight_decay = self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay
        d.bias_corr1, d.bias_corr2 = 1 - self.betas[0] ** t, 1 - self.betas[1] ** t
        d.ema_decay, d.grad_scale = ema_decay, grad_scale
        nv.check(nv.lib().vt_adamw_ema_step(C.byref(d), C.c_void_p(nv.current_stream_ptr())))
        for p in self.params:          # the kernel wrote the parameters in place: bump their version counters so that
            _bump_version(p)           # engines holding packed copies (BridgeEngine, LossProgram) re-pack lazily


def bucket_plan(numels: Sequence[int], bucket_elems: int = 32 << 20) -> List[List[int]]:
    """Greedy buckets of consecutive tensors (reverse registration order = roughly the order gradients become ready)."""
    buckets, cur, size = [], [], 0
    for i in reversed(range(len(numels))):
        if cur and size + numels[i] > bucket_elems:
            buckets.append(cur)
            cur, size = [], 0
        cur.append(i)
        size += numels[i]
    if cur:
        buckets.append(cur)
    return buckets


def allreduce_gradients(grads: Sequence[torch.Tensor], group=None, bucket_elems: int = 32 << 20, async_op: bool = False):
    """SUM all-reduce of `grads` in flat buckets (the only collective on the refinement path).  Returns the
    world size; scale by 1/world in the optimizer (`FusedAdamWEMA.step(grad_scale=1/world)`)."""
    import torch.distributed as dist
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return 1
    work = []
    for bucket in bucket_plan([g.numel() for g in grads], bucket_elems):
        flat = torch.cat([grads[i].reshape(-1) for i in bucket])
        h = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
        work.append((h, flat, bucket))
    for h, flat, bucket in work:
        h.wait()
        off = 0
        for i in bucket:
            n = grads[i].numel()
            grads[i].copy_(flat[off: off + n].view_as(grads[i]))
            off += n
    return dist.get_world_size(group)
