"""Training-side pieces of the bridge / LSTM trainers (reference: bridge_train.py:296-342, lstm_train.py:122-139):

* FusedAdamWEMA -- torch.optim.AdamW(lr, weight_decay, default betas/eps) + torch_ema update + CosineAnnealingLR as ONE
  multi-tensor kernel launch per step (csrc/vt_elem.cuh adamw_ema_kernel) instead of ~1300 per-tensor launches.  Gradients
  are read either from `.grad` (torch autograd, fixed buffers) or IN PLACE from the buffers the native backward programs
  wrote them to (`grad_sources`: the weight-gradient GEMM's own [rows][taps][c_pad] layout), so the training step has no
  per-parameter unpack / copy pass.
* allreduce_gradients / allreduce_arena -- the single data-parallel exchange step of the path: SUM all-reduce of the gradients
  over torch.distributed (NCCL over NVLink 5 / NVSwitch on the B200 box, gloo in the CPU tests); the 1/world scaling is folded
  into the optimizer kernel (`grad_scale`).  The arena form reduces slices of the contiguous gradient arena in place."""
from __future__ import annotations

import ctypes as C
import math
from typing import Iterable, List, Optional, Sequence

import torch

from . import native as nv

CHUNK = 65536


def _bump_versions(params: Sequence[torch.Tensor]) -> None:
    """Advance the autograd version counters of tensors a native kernel has written in place (one call for all of them)."""
    try:
        torch._C._autograd._unsafe_set_version_counter(tuple(params), tuple(p._version + 1 for p in params))
    except Exception:
        with torch.no_grad():
            torch._foreach_add_(list(params), 0)


def _bump_version(p: torch.Tensor) -> None:
    _bump_versions((p,))


class FusedAdamWEMA:
    def __init__(self, params: Iterable[torch.nn.Parameter], lr: float = 1e-4, betas=(0.9, 0.999), eps: float = 1e-8,
                 weight_decay: float = 1e-6, ema=None, ema_params: Optional[Sequence[torch.nn.Parameter]] = None,
                 t_max: Optional[int] = None, eta_min: Optional[float] = None, grad_sources: Optional[dict] = None):
        """`ema`: a vla_touch_b200.ema.ExponentialMovingAverage whose shadow list lines up with `ema_params` (a subset of
        `params`, e.g. the U-Net parameters but not the state encoder, bridge_train.py:50-57, bridge_model.py:433).
        `grad_sources`: {id(param): (gradient buffer, taps, c, c_pad)} for parameters whose gradient is produced by a native
        program (LossBackwardProgram.grad_sources()); those get no `.grad` buffer and are read in place."""
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
        self._ema_params = None
        if ema is not None:
            self._ema_params = list(ema_params if ema_params is not None else self.params)
            if len(self._ema_params) != len(ema.shadow_params):
                raise ValueError("ema_params must line up with ema.shadow_params")
        self._sources = dict(grad_sources or {})
        # fixed gradient buffers for the parameters torch autograd fills
        # (views of ONE flat buffer, `grad_flat`: the data-parallel all-reduce of these gradients runs in place on it)
        own = [p for p in self.params if id(p) not in self._sources]
        sizes = [(p.numel() + 63) // 64 * 64 for p in own]
        self.grad_flat = torch.zeros(sum(sizes), dtype=torch.float32, device=dev)
        views, off = {}, 0
        for p, n in zip(own, sizes):
            views[id(p)] = self.grad_flat[off: off + p.numel()].view(p.shape)
            off += n
        self._grads = [views.get(id(p)) for p in self.params]
        for p, g in zip(self.params, self._grads):
            if g is not None:
                p.grad = g
        self._dev = dev
        self._shadow_ptrs = None
        self._build_records()

    def _shadows(self):
        if self.ema is None:
            return [None] * len(self.params)
        shadow = {id(p): s for p, s in zip(self._ema_params, self.ema.shadow_params)}
        return [shadow.get(id(p)) for p in self.params]

    def _build_records(self) -> None:
        self._shadow = self._shadows()
        recs = (nv.OptTensor * len(self.params))()
        chunks = []
        for i, p in enumerate(self.params):
            assert p.dtype == torch.float32 and p.is_contiguous()
            s = self._shadow[i]
            if s is not None and (s.device != p.device or s.dtype != torch.float32 or not s.is_contiguous()):
                raise ValueError("EMA shadow parameters must be contiguous fp32 tensors on the parameters' device")
            src = self._sources.get(id(p))
            if src is None:
                g, taps, c, c_pad = self._grads[i], 0, 0, 0
            else:
                g, taps, c, c_pad = src
                assert g.dtype == torch.float32 and g.is_contiguous() and g.device == p.device
                if taps:
                    rows = p.numel() // (c * taps)
                    assert rows * c * taps == p.numel() and g.numel() >= rows * taps * c_pad and c <= c_pad, (tuple(p.shape), taps, c, c_pad)
                else:
                    assert g.numel() >= p.numel()
            recs[i] = nv.OptTensor(p.data_ptr(), g.data_ptr(), self.m[i].data_ptr(), self.v[i].data_ptr(),
                                   s.data_ptr() if s is not None else None, p.numel(), taps, c, c_pad, 0, None)
            chunks += [(i, off) for off in range(0, p.numel(), CHUNK)]
        self._recs = torch.frombuffer(bytearray(bytes(recs)), dtype=torch.uint8).to(self._dev)
        self._chunks = torch.tensor(chunks, dtype=torch.int64).to(self._dev)
        self.n_chunks = len(chunks)
        self._shadow_ptrs = [s.data_ptr() if s is not None else 0 for s in self._shadow]
        self._param_ptrs = [p.data_ptr() for p in self.params]

    def set_grad_sources(self, grad_sources: dict) -> None:
        """Point the records at another program's gradient buffers (a new batch shape built a new training program)."""
        if set(grad_sources) != set(self._sources):
            raise ValueError("the set of natively produced gradients cannot change after construction")
        self._sources = dict(grad_sources)
        self._build_records()

    @torch.no_grad()
    def gather_grads(self) -> None:
        """Make sure the gradients torch autograd produced sit in the fixed flat buffer (autograd may have re-bound `.grad`)."""
        for p, g in zip(self.params, self._grads):
            if g is not None and p.grad is not None and p.grad.data_ptr() != g.data_ptr():
                g.copy_(p.grad)
                p.grad = g

    def zero_grad(self, set_to_none: bool = False) -> None:
        self.grad_flat.zero_()
        for p, g in zip(self.params, self._grads):
            if g is not None:
                p.grad = g

    def current_lr(self) -> float:
        if self.t_max is None:
            return self.base_lr
        # CosineAnnealingLR(T_max, eta_min) stepped once per optimizer step: the closed form, periodic past T_max like torch's
        return self.eta_min + (self.base_lr - self.eta_min) * (1 + math.cos(math.pi * self.step_count / self.t_max)) / 2

    @torch.no_grad()
    def step(self, grad_scale: float = 1.0) -> None:
        self.gather_grads()                                 # a backward pass may have re-bound .grad
        if [p.data_ptr() for p in self.params] != self._param_ptrs:      # p.data was re-pointed (parameter arena of the re-pack)
            self._build_records()
        if self.ema is not None:                            # ema.load_state_dict() / ema.to() replace the shadow tensors
            cur = self._shadows()
            if [s.data_ptr() if s is not None else 0 for s in cur] != self._shadow_ptrs:
                self._build_records()
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
        _bump_versions(self.params)    # the kernel wrote the parameters in place: engines holding packed copies re-pack lazily


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


def arena_buckets(allocs: Sequence, total: int, bucket_elems: int = 24 << 20) -> List[tuple]:
    """Cut the gradient arena (allocations (offset, numel, op index) in production order) into contiguous buckets of about
    `bucket_elems` elements.  -> [(first element, last element + 1, op index after which the bucket is complete)], the last
    bucket's op index is None (= end of the program).  A bucket is complete once the program has reached the op count that was
    current when the NEXT bucket's first tensor was allocated: every builder allocates its outputs and then appends the ops that
    fill them before the next builder runs (a cut is only made where the op count has advanced since the previous allocation)."""
    out, start = [], 0
    for i, (off, n, _) in enumerate(allocs):
        nxt = allocs[i + 1] if i + 1 < len(allocs) else None
        if nxt is not None and nxt[0] - start >= bucket_elems and nxt[2] > allocs[i][2]:
            out.append((start, nxt[0], nxt[2]))
            start = nxt[0]
    out.append((start, total, None))
    return out


def allreduce_arena(flat: torch.Tensor, group=None):
    """In-place SUM all-reduce of a slice of the contiguous gradient arena (no staging copy).  Returns the async work handle."""
    import torch.distributed as dist
    return dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group, async_op=True)
