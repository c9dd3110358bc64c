#!/usr/bin/env python
"""Benchmark of the VLA-Touch action-refinement hot path on B200 (contract: see DESIGN.md 'Measurement').

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload NAME] [--only-headline]

Headline (BASELINE.json `metric`): refined action-chunks/s at batch 256 per GPU = `DiffusionController.predict` rows
(one chunk = 2 camera images + state + tactile + base chunk -> refined [T, A] chunk), whole-job aggregate over N GPUs.
  value : device-resident inputs, K CUDA-graph launches of the whole predict program, CUDA events, max over ranks.
  e2e   : the same metric through the public API with pinned HOST inputs, H2D copies and the D2H read of the result
          inside the timed region.
The same JSON line carries, under "workloads", the other BASELINE configs measured by the same process on the same GPUs:
  cfg2_train : bridge_train.py step (DinoV2 x 2 + state encoder + get_loss forward/backward of 3 U-Nets + NCCL gradient
               all-reduce + fused AdamW/EMA), batch 256 per GPU                                   (BASELINE configs[1])
  cfg3       : 50-step sampler, DinoV2-B/14, global batch 1024 (1024 / N rows per GPU: strong scaling)  (configs[2])
  cfg4_lstm_train : lstm_train.py step, seq_len 128, batch 512 per GPU, NCCL gradient all-reduce        (configs[3])
  cfg5_latency    : predict() latency p50 per call, global batch 64 (64 / N rows per GPU), host inputs  (configs[4])
  cfg2_strong     : the headline at GLOBAL batch 256 (256 / N rows per GPU)
--impl reference : the CPU restatement of the reference (oracle/, "port") on the host cores, bounded sample.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

WORKLOADS = {
    # name: (dino variant, hidden, heads, layers, hw, T, A, F, steps, batch per GPU)
    "cfg2": ("facebook/dinov2-small", 384, 6, 12, 224, 64, 7, 64, 10, 256),
    "cfg3": ("facebook/dinov2-base", 768, 12, 12, 224, 64, 7, 64, 50, 1024),
    "cfg1": ("facebook/dinov2-small", 384, 6, 12, 384, 16, 10, 3, 10, 1),
    "cfg5": ("facebook/dinov2-small", 384, 6, 12, 224, 64, 7, 64, 10, 64),
}
MODEL_ARGS = {'interpolant_type': 'linear', 'gamma_type': '2^0.5*t(t-1)', 'epsilon_type': '1-t', 'prior_policy': 'vla',
              'beta_max': 0.03, 'sde_type': 'vs', 'obs_dim': 256, 'obs_horizon': 1, 'net_type': 'unet1D_si',
              'pretrain': False, 'context_frames': 2}


def dino_flops(hidden, layers, hw):
    n_tok = (hw // 14) ** 2 + 1
    return 2 * (n_tok - 1) * 588 * hidden + layers * (24 * n_tok * hidden ** 2 + 4 * n_tok ** 2 * hidden)


def unet_flops(T):
    return 0.0115e9 + 0.02008e9 * T


def flops_per_chunk(hidden, layers, hw, T, A, F, steps):
    """F = 2 F_dino + F_enc + n * 2 * F_unet(T)."""
    f_enc = 2 * ((2 * hidden + A + F) * 256 + 2 * 256 ** 2)
    return 2 * dino_flops(hidden, layers, hw) + f_enc + steps * 2 * unet_flops(T)


def flops_per_train_sample(hidden, layers, hw, T):
    """2 DinoV2 forwards + 3 U-Net forwards + 3 U-Net backwards (2 x forward) = 2 F_dino + 9 F_unet."""
    return 2 * dino_flops(hidden, layers, hw) + 9 * unet_flops(T)


def synth_weights(hidden, layers, A, F, seed=7):
    from vla_touch_b200 import shapes as shp
    from vla_touch_b200 import synthetic as syn
    dino = syn.synth_state_dict(shp.dinov2_shapes(hidden, layers), seed, prefix="dino.")
    enc = syn.synth_state_dict(shp.mlp_shapes([2 * hidden + A + F, 256, 256, 256]), seed, prefix="enc.")
    net = syn.synth_state_dict(shp.si_net_shapes(A, 256), seed, prefix="net.")
    return dino, enc, net


class ClockSampler(threading.Thread):
    """Samples nvidia-smi SM clocks and throttle reasons during the timed region."""

    Q = "clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown," \
        "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap"

    def __init__(self, index):
        super().__init__(daemon=True)
        self.index, self.samples, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}", "--format=csv,noheader,nounits"],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.samples.append([x.strip() for x in out.split(",")])
            except Exception:
                pass
            time.sleep(0.1)

    def summary(self):
        sm = sorted(int(s[0]) for s in self.samples if s[0].isdigit())
        mx = [int(s[1]) for s in self.samples if s[1].isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(s) > 2 + i and s[2 + i].lower().startswith("active") for s in self.samples)]
        return {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": max(mx) if mx else None, "reasons": reasons,
                "samples": len(self.samples)}


# ----------------------------------------------------------------------------------------------------------------------
# CPU arm: the oracle port of the reference
# ----------------------------------------------------------------------------------------------------------------------
def _oracle_predict_fn(wl, Bs):
    import torch
    from oracle import vt_oracle as orc
    from vla_touch_b200 import synthetic as syn
    name, hidden, heads, layers, hw, T, A, F, steps, batch = wl
    dino, enc, net = synth_weights(hidden, layers, A, F)
    sub = lambda p: {k[len(p):]: v for k, v in net.items() if k.startswith(p)}
    v_sd, s_sd = sub("v_net."), sub("s_net.")
    inp = syn.synth_predict_inputs(Bs, T, A, F, hw, 1234)
    stats = syn.synth_stats(A)
    noise = syn.det_normal("bench.noise", (steps, Bs, T, A), 1)

    def step():
        with torch.no_grad():
            return orc.predict(dino, enc, v_sd, s_sd, stats, heads, inp["state"], inp["vla_actions"], inp["images_cam1"][:, None],
                               inp["images_cam2"][:, None], inp["forces"], steps, 0.03, noise)
    return step


def run_reference(args, wl):
    """--impl reference: the oracle port of the reference CPU path, all host threads, bounded sample per step
    (B = 32 rows per predict() call)."""
    import torch
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, hidden, heads, layers, hw, T, A, F, steps, batch = wl
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bs = args.ref_batch
    step = _oracle_predict_fn(wl, Bs)
    for _ in range(args.warmup):
        step()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        step()
    dt = time.perf_counter() - t0
    val = Bs * args.steps / dt
    line = {"impl": "reference", "metric": "refined action-chunks/sec", "value": val, "unit": "chunks/s", "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": {"workload": f"{args.workload}: predict, {name}, {hw}x{hw} x2 cams, T={T}, A={A}, F={F}, {steps} SDE steps",
                       "batch_per_step": Bs},
            "cpu_baseline": {"value": val, "unit": "chunks/s", "cores": cores, "kind": "port",
                             "sample": f"{args.steps} predict() calls of {Bs} rows (oracle/vt_oracle.py, torch CPU fp32, {cores} threads)"},
            "e2e": {"value": val, "unit": "chunks/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    print(json.dumps(line), flush=True)


def cpu_baseline(wl, seconds=15.0, Bs=8):
    import torch
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    f = _oracle_predict_fn(wl, Bs)
    f()
    n, t0 = 0, time.perf_counter()
    while time.perf_counter() - t0 < seconds:
        f()
        n += 1
    dt = time.perf_counter() - t0
    return {"value": Bs * n / dt, "unit": "chunks/s", "cores": cores, "kind": "port",
            "sample": f"{n} predict() calls of {Bs} rows in {dt:.1f}s (oracle/vt_oracle.py restatement, torch CPU fp32, {cores} threads)"}


# ----------------------------------------------------------------------------------------------------------------------
# GPU arm
# ----------------------------------------------------------------------------------------------------------------------
def gemm_flops(d):
    """FLOPs of one tensor-core launch: an implicit GEMM, or the fused ViT MLP (two GEMMs rows x D x 4D)."""
    if hasattr(d, "w2"):                       # MlpDesc
        return 2.0 * 2.0 * d.rows * d.D * 4 * d.D
    if hasattr(d, "colscale") and hasattr(d, "ln_out"):   # RowprojDesc: rows x D x D
        return 2.0 * d.rows * d.D * d.D
    return 2.0 * d.G * d.M * d.N * d.taps * d.kc * d.passes


def op_kind(d):
    """Kernel family of a program op (the grouping the `roofline` block ranks by time share)."""
    from vla_touch_b200 import native as nv
    if isinstance(d, nv.GemmDesc):
        return "gemm_tc_kernel<GroupNorm+Mish+FiLM epilogue> (U-Net k5 convs)" if d.epi == nv.EPI_GN else \
            ("gemm_tc_kernel<linear epilogue> (U-Net 1x1 / strided convs, FiLM)" if d.taps > 1 or d.t_box != 128 or d.G > 1
             else "gemm_tc_kernel<linear epilogue> (ViT / encoder linears)")
    if isinstance(d, nv.MlpDesc):
        return "mlp_fused_kernel (ViT fc1+GELU+fc2)"
    if isinstance(d, nv.RowprojDesc):
        return "rowproj_kernel (ViT attention out-projection + norm2)"
    if isinstance(d, nv.AttnDesc):
        return "attn_row_kernel (ViT attention)"
    if isinstance(d, nv.PersistDesc):
        return "unet_persist_kernel (sde_vs: all steps x [v_net + s_net evaluation, Euler-Maruyama update], one launch)"
    return type(d).__name__.replace("Desc", "").lower() + "_kernel"


class Ctx:
    """Per-process benchmark context: rank / world, timing helpers."""

    def __init__(self, args):
        import torch
        import torch.distributed as dist
        self.torch, self.dist, self.args = torch, dist, args
        self.rank, self.world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        torch.cuda.set_device(self.local)
        self.dev = f"cuda:{self.local}"
        if self.world > 1:
            dist.init_process_group("nccl", device_id=torch.device("cuda", self.local))

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        self.torch.cuda.synchronize()

    def timed(self, fn, n):
        """ms for n calls: CUDA events on the launching stream, barrier + synchronize on both sides, max over ranks."""
        torch = self.torch
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        self.barrier()
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        self.barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(ms, op=self.dist.ReduceOp.MAX)
        return float(ms)


def make_controller(cx, wl, steps=None):
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.bridge_controller import DiffusionController
    name, hidden, heads, layers, hw, T, A, F, n_steps, batch = wl
    dino_sd, enc_sd, net_sd = synth_weights(hidden, layers, A, F)
    ctl = DiffusionController(state_dim=A, hidden_dim=256, image_model_path=name, diffusion_steps=steps or n_steps, device=cx.dev,
                              model_args=dict(MODEL_ARGS, action_dim=A, horizon=T), use_force=True, force_dim=F,
                              image_state_dict=dino_sd)
    ctl.state_encoder.load_state_dict(enc_sd)
    ctl.diffusion_model.net.load_state_dict(net_sd)
    ctl.diffusion_model.ema = type(ctl.diffusion_model.ema)(ctl.diffusion_model.net.parameters(), decay=0.75)
    ctl.stats = {k: v.to(cx.dev) for k, v in syn.synth_stats(A).items()}
    return ctl


def predict_harness(cx, wl, batch):
    """(controller, engine, api_step, h2d bytes, d2h bytes) for predict() on `batch` rows with pinned host inputs."""
    torch = cx.torch
    from vla_touch_b200 import synthetic as syn
    name, hidden, heads, layers, hw, T, A, F, steps, _ = wl
    ctl = make_controller(cx, wl)
    inp = syn.synth_predict_inputs(batch, T, A, F, hw, 1234 + cx.rank)
    host = {k: v.pin_memory() for k, v in inp.items()}
    host["images_cam1"] = inp["images_cam1"][:, None].contiguous().pin_memory()   # deployment layout [B,1,H,W,3] uint8
    host["images_cam2"] = inp["images_cam2"][:, None].contiguous().pin_memory()
    h2d = sum(host[k].numel() * host[k].element_size() for k in ("state", "forces", "vla_actions", "images_cam1", "images_cam2"))
    out_host = torch.empty(batch, T, A, dtype=torch.float32).pin_memory()

    def api_step():
        out = ctl.predict(host["state"], host["vla_actions"], host["images_cam1"], host["images_cam2"], host["forces"])
        out_host.copy_(out, non_blocking=True)

    api_step()                      # builds the engine, the FiLM time tables and the CUDA graph
    torch.cuda.synchronize()
    eng = next(iter(ctl._engines.values()))
    keys = ("state", "forces", "vla_actions", "images_cam1", "images_cam2")
    land = {k: torch.empty_like(host[k], device=cx.dev) for k in keys}

    def h2d_only():                 # the step's host->device traffic alone (what the e2e number adds on top of `value`)
        for k in keys:
            land[k].copy_(host[k], non_blocking=True)
    api_step.h2d_only = h2d_only
    return ctl, eng, api_step, h2d, out_host.numel() * 4


def bench_predict(cx, wl, batch, steps, warmup):
    ctl, eng, api_step, h2d, d2h = predict_harness(cx, wl, batch)
    for _ in range(warmup):
        eng.run_predict(graph=True)
    ms = cx.timed(lambda: eng.run_predict(graph=True), steps)
    for _ in range(warmup):
        api_step()
    ms_e2e = cx.timed(api_step, steps)
    ms_h2d = cx.timed(api_step.h2d_only, steps)
    return dict(ctl=ctl, eng=eng, ms=ms, ms_e2e=ms_e2e, ms_h2d=ms_h2d, h2d=h2d, d2h=d2h, launches=eng.num_launches())


def peaks():
    try:
        p = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        return p, "MEASURED_PEAKS.json"
    except Exception:
        return {"bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "hbm_gbs": 6650.0}, "fallback (B200_PROFILING.md)"


def roofline_block(cx, eng, wl, workload, value, world):
    """The kernel family with the largest share of the step (each op of the program timed alone, back to back), its achieved
    algorithmic TFLOP/s against the burst bf16 peak, plus the fused ViT MLP / attention kernels and the whole step."""
    torch = cx.torch
    from vla_touch_b200 import native as nv
    name, hidden, heads, layers, hw, T, A, F, steps, batch = wl
    pk, pk_src = peaks()
    peak_tf = pk["bf16_tflops"]
    prog = eng.plan.compile()
    a0, b1 = eng.predict_range()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    kinds = {}
    reps = 3
    for i in range(a0, b1):
        d = eng.plan.descs[i]
        prog.run(i, 1)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(reps):
            prog.run(i, 1)
        ev1.record()
        torch.cuda.synchronize()
        us = ev0.elapsed_time(ev1) * 1e3 / reps
        k = kinds.setdefault(op_kind(d), {"us": 0.0, "flops": 0.0, "launches": 0})
        k["us"] += us
        k["launches"] += 1
        if isinstance(d, nv.PersistDesc):
            k["flops"] += d.algo_flops
        elif isinstance(d, (nv.GemmDesc, nv.MlpDesc, nv.RowprojDesc)):
            k["flops"] += gemm_flops(d)
        elif isinstance(d, nv.AttnDesc):
            k["flops"] += 4.0 * d.tokens * d.tokens * 64 * d.heads * d.images
    total_us = sum(k["us"] for k in kinds.values())
    ranked = sorted(kinds.items(), key=lambda kv: -kv[1]["us"])
    dom_name, dom = ranked[0]
    ach = dom["flops"] / (dom["us"] * 1e-6) / 1e12 if dom["us"] else 0.0
    traffic, traffic_src = None, None
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        rec = table.get(workload, {}).get(dom_name.split(" ")[0])
        if rec:
            traffic, traffic_src = rec["dram_bytes"], rec["source"]
    except Exception:
        pass
    f_chunk = flops_per_chunk(hidden, layers, hw, T, A, F, eng.n_steps)
    step_tf = value / world * f_chunk / 1e12
    by_kind = [{"kernel": n, "share_of_step": k["us"] / total_us, "launches": k["launches"], "us_per_launch": k["us"] / k["launches"],
                "achieved_tflops": (k["flops"] / (k["us"] * 1e-6) / 1e12) if k["flops"] else None,
                "frac_of_peak": (k["flops"] / (k["us"] * 1e-6) / 1e12 / peak_tf) if k["flops"] else None} for n, k in ranked[:6]]
    return {"bound": "tensor", "kernel": dom_name, "selected_by": "largest share of the step's summed per-op times",
            "share_of_step": dom["us"] / total_us, "launches_per_step": dom["launches"],
            "achieved": ach, "peak": peak_tf, "unit": "TFLOP/s", "frac": ach / peak_tf,
            "flops_per_launch": dom["flops"] / dom["launches"], "ms_per_launch": dom["us"] / dom["launches"] * 1e-3,
            "traffic": traffic, "traffic_source": traffic_src,
            "peak_source": f"{pk_src} bf16_tflops (burst: ops timed alone, back to back)",
            "by_kernel": by_kind,
            "whole_step": {"algorithmic_tflops_per_gpu": step_tf, "frac_of_sustained": step_tf / pk.get("bf16_tflops_sustained", 1400.0),
                           "gflop_per_chunk": f_chunk / 1e9}}


def synth_train_batch(cx, wl, batch, device):
    """A ControllerDataset-style minibatch (controller_dataset.py:142-236): ctx = 2 context frames; only the LAST context image is
    read by the trainer (bridge_train.py:143-144), so the image tensors carry that frame alone, uint8 [B,1,H,W,3]."""
    torch = cx.torch
    from vla_touch_b200 import synthetic as syn
    name, hidden, heads, layers, hw, T, A, F, steps, _ = wl
    s = 4321 + cx.rank
    b = {"states": syn.det_normal("tr.states", (batch, 2 + T, A), s), "forces": syn.det_normal("tr.forces", (batch, 2 + T, F), s),
         "vla_actions": syn.det_uniform("tr.vla", (batch, T, A), s, -1.0, 1.0),
         "images_cam1": syn.synth_images_u8("tr.cam1", batch, hw, s)[:, None].contiguous(),
         "images_cam2": syn.synth_images_u8("tr.cam2", batch, hw, s)[:, None].contiguous()}
    b["expert_actions"] = (b["vla_actions"] + 0.1 * syn.det_normal("tr.delta", (batch, T, A), s)).clamp(-1, 1)
    if device == "pinned":
        return {k: v.pin_memory() for k, v in b.items()}
    return {k: v.to(device) for k, v in b.items()}


def bench_train(cx, wl, batch, steps, warmup):
    """bridge_train.py:296-342 at `batch` rows per GPU through vla_touch_b200.trainer.DiffusionControllerTrainer."""
    torch, dist = cx.torch, cx.dist
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.trainer import DiffusionControllerTrainer
    name, hidden, heads, layers, hw, T, A, F, _, _ = wl
    ctl = make_controller(cx, wl)
    tr = DiffusionControllerTrainer(ctl, syn.synth_stats(A), device=cx.dev)
    dev_batch = synth_train_batch(cx, wl, batch, cx.dev)
    host_batch = synth_train_batch(cx, wl, batch, "pinned")
    loss_host = torch.empty(4, dtype=torch.float32).pin_memory()
    losses = []

    def dev_step():
        losses.append(tr.train_step(dict(dev_batch))["loss"])

    def api_step():
        out = tr.train_step(dict(host_batch))          # H2D of the minibatch inside (images, states, forces, chunks)
        loss_host.copy_(torch.stack([out["loss"], out["v_loss"], out["s_loss"], out["b_loss"]]), non_blocking=True)

    for _ in range(max(warmup, 2)):
        dev_step()
    torch.cuda.synchronize()
    l0 = float(losses[0])
    ms = cx.timed(dev_step, steps)
    l1 = float(losses[-1])
    for _ in range(2):
        api_step()
    ms_e2e = cx.timed(api_step, steps)
    prog = tr._prog
    arena, _ = prog.grad_arena()
    ar_ms = None
    if cx.world > 1:                                   # the collective alone (not overlapped), for its share of the step
        for _ in range(2):
            dist.all_reduce(arena)
        ar_ms = cx.timed(lambda: dist.all_reduce(arena), 5) / 5
    if tr.timing:
        print("train_step host ms per phase (accumulated):", {k: round(v, 1) for k, v in tr.timing.items()}, file=sys.stderr)
    f_s = flops_per_train_sample(hidden, layers, hw, T)
    pk, _ = peaks()
    sps = cx.world * batch * steps / (ms * 1e-3)
    h2d = sum(v.numel() * v.element_size() for v in host_batch.values())
    return {"metric": "training samples/sec (bridge_train.py step)", "value": sps, "unit": "samples/s", "ms_per_step": ms / steps,
            "batch_per_gpu": batch, "global_batch": batch * cx.world, "scaling": "weak", "dtype": "bf16",
            "e2e": {"value": cx.world * batch * steps / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": 16},
            "step": "encode_observation (DinoV2 x2 no_grad + state encoder with grad) -> get_loss forward+backward (b/v/s U-Nets, "
                    "one native program) -> gradient all-reduce -> fused AdamW + EMA + cosine LR",
            "gpu_launches_per_step": prog.plan.compile().num_launches() + 2 * 160 + 1,
            "grad_elements": int(arena.numel()), "buckets": len(tr._buckets),
            "collective": None if cx.world == 1 else {
                "op": "NCCL all-reduce (SUM) of the fp32 gradient arena, in place, per bucket on a side stream during backward",
                "ranks": cx.world, "bytes": int(arena.numel()) * 4, "ms_alone": ar_ms, "share_of_step_if_exposed": ar_ms / (ms / steps)},
            "loss_first": l0, "loss_last": l1,
            "roofline": {"bound": "tensor", "gflop_per_sample": f_s / 1e9, "achieved_tflops_per_gpu": sps / cx.world * f_s / 1e12,
                         "frac_of_sustained": sps / cx.world * f_s / 1e12 / pk.get("bf16_tflops_sustained", 1400.0)}}


def bench_train_cached(cx, wl, batch, steps, warmup, episodes=4, frames=168):
    """The same training step fed from the HBM-resident episode store with the DinoV2 feature cache (SURVEY.md 8f N2;
    vla_touch_b200/episode_store.py): per step the host sends `batch` sample numbers, vt_batch_gather assembles the collated,
    normalised minibatch + cached features of both cameras, and the frozen encoder no longer runs.  Episodes follow the reference's
    HDF5 schema (A = 10 pose dims, 3 force dims, 64 x 10 VLA chunks, 224 x 224 uint8 frames), written as .vtep shards to a
    temporary directory, `episodes` x `frames` frames per rank."""
    import shutil
    import tempfile
    import numpy as np
    torch = cx.torch
    from vla_touch_b200 import controller_dataset as cd
    from vla_touch_b200 import episode_store as es
    from vla_touch_b200.synthetic import synth_episode
    from vla_touch_b200.trainer import DiffusionControllerTrainer
    name, hidden, heads, layers, hw, T, _, _, n_steps, _ = wl
    wl10 = (name, hidden, heads, layers, hw, T, 10, 3, n_steps, batch)
    td = tempfile.mkdtemp(prefix="vtep_")
    try:
        for e in range(episodes):
            es.write_episode_shard(synth_episode(1000 * cx.rank + e, frames, hw, still_frames=2), os.path.join(td, f"episode_{e}.vtep"))
        import contextlib
        with contextlib.redirect_stdout(sys.stderr):          # the dataset prints progress like the reference's does
            ds = cd.ControllerDataset(td, context_frames=2, horizon=T, use_images=True, image_size=hw)
        ctl = make_controller(cx, wl10)
        t0 = time.perf_counter()
        store = ds.device_store(cx.dev, image_encoder=ctl.image_encoder, feature_chunk=128)
        torch.cuda.synchronize()
        fill_s = time.perf_counter() - t0
    finally:
        shutil.rmtree(td, ignore_errors=True)
    tr = DiffusionControllerTrainer(ctl, ds.stats, device=cx.dev)
    sampler = cd.EpisodeBatchSampler(len(ds), batch, 0, 1, seed=cx.rank)      # every rank owns its own episodes here
    epoch = [0]
    losses = []

    def next_indices():
        sampler.set_epoch(epoch[0])
        epoch[0] += 1
        return next(iter(sampler))

    fixed = torch.from_numpy(next_indices()).to(cx.dev)

    def dev_step():
        losses.append(tr.train_step(store.gather(fixed))["loss"])

    loss_host = torch.empty(4, dtype=torch.float32).pin_memory()

    def api_step():
        out = tr.train_step(store.gather(next_indices()))     # host -> device: the sample numbers
        loss_host.copy_(torch.stack([out["loss"], out["v_loss"], out["s_loss"], out["b_loss"]]), non_blocking=True)

    for _ in range(max(warmup, 2)):
        dev_step()
    torch.cuda.synchronize()
    ms = cx.timed(dev_step, steps)
    for _ in range(2):
        api_step()
    ms_e2e = cx.timed(api_step, steps)
    reps = 50
    ms_gather = cx.timed(lambda: store.gather(fixed), reps) / reps
    L, A, H = 2 + T, store.A, T
    per_sample = 4 * (L * A + 2 * H * A + L * store.Fd + L * store.Dd + 2 * store.D)          # fp32 elements read from the store
    per_sample_out = per_sample + 4 * 2 * H * A + 4 * H * A                                  # + expert_actions copy, two normalised chunks
    pk, _ = peaks()
    gbs = batch * (per_sample + per_sample_out) / (ms_gather * 1e-3) / 1e9
    sps = cx.world * batch * steps / (ms * 1e-3)
    f_s = 9 * unet_flops(T)
    return {"metric": "training samples/sec (bridge_train.py step fed from the HBM episode store, DinoV2 features cached)",
            "value": sps, "unit": "samples/s", "ms_per_step": ms / steps, "batch_per_gpu": batch, "global_batch": batch * cx.world,
            "scaling": "weak", "dtype": "bf16",
            "e2e": {"value": cx.world * batch * steps / (ms_e2e * 1e-3), "unit": "samples/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": batch * 8, "d2h_bytes_per_step": 16},
            "store": {"episodes": episodes, "frames": store.frames, "samples": len(store), "A": A, "force_dim": store.Fd,
                      "fill_seconds_incl_feature_cache": fill_s, "feature_cache_bytes": int(store.feats.numel()) * 4},
            "collective": None if cx.world == 1 else {
                "op": "NCCL all-reduce (SUM) of the fp32 gradient arena, in place, per bucket on a side stream during backward",
                "ranks": cx.world, "bytes": int(tr._prog.grad_arena()[0].numel()) * 4},
            "gather_kernel": {"ms": ms_gather, "algorithmic_bytes": batch * (per_sample + per_sample_out), "achieved_gbs": gbs,
                              "peak_gbs": pk.get("hbm_gbs"), "note": "launch-latency-bound at this size (10 MB per minibatch)"},
            "loss_first": float(losses[0]), "loss_last": float(losses[-1]),
            "roofline": {"bound": "tensor", "gflop_per_sample": f_s / 1e9, "achieved_tflops_per_gpu": sps / cx.world * f_s / 1e12,
                         "frac_of_sustained": sps / cx.world * f_s / 1e12 / pk.get("bf16_tflops_sustained", 1400.0)}}


def make_lstm_controller(cx, wl):
    from vla_touch_b200 import shapes as shp
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.lstm_step_controller import TactileLSTMController
    name, hidden, heads, layers, hw, T, A, F, _, _ = wl
    dino_sd, _, _ = synth_weights(hidden, layers, A, F)
    lc = TactileLSTMController(state_dim=A, hidden_dim=256, num_layers=2, dropout=0.1, image_model_path=name, device=cx.dev,
                               force_dim=F, use_force=True, image_state_dict=dino_sd)
    for nm, mod in (("obs_encoder", lc.obs_encoder), ("force_encoder", lc.force_encoder), ("lstm", lc.lstm), ("output_head", lc.output_head)):
        syn.fill_named_(mod.named_parameters(), 41, prefix=f"lstm.{nm}.")
    return lc


def bench_lstm_train(cx, wl, batch, T, steps, warmup):
    """lstm_train.py:122-139 (BASELINE configs[3]: seq_len 128, batch 512 per GPU) through trainer.LSTMControllerTrainer."""
    torch, dist = cx.torch, cx.dist
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.trainer import LSTMControllerTrainer
    name, hidden, heads, layers, hw, _, A, F, _, _ = wl
    lc = make_lstm_controller(cx, wl)
    tr = LSTMControllerTrainer(lc, syn.synth_stats(A), device=cx.dev)
    s = 777 + cx.rank
    b = {"states": syn.det_normal("lt.states", (batch, 2 + T, A), s), "forces": syn.det_normal("lt.forces", (batch, 2 + T, F), s),
         "vla_actions": syn.det_uniform("lt.vla", (batch, T, A), s, -1.0, 1.0),
         "images_cam1": syn.synth_images_u8("lt.cam1", batch, hw, s)[:, None].contiguous(),
         "images_cam2": syn.synth_images_u8("lt.cam2", batch, hw, s)[:, None].contiguous()}
    b["expert_actions"] = (b["vla_actions"] + 0.1 * syn.det_normal("lt.delta", (batch, T, A), s)).clamp(-1, 1)
    dev_batch = {k: v.to(cx.dev) for k, v in b.items()}
    host_batch = {k: v.pin_memory() for k, v in b.items()}
    losses = []
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def dev_step():
        losses.append(tr.train_step(dict(dev_batch)))

    def api_step():
        loss_host.copy_(tr.train_step(dict(host_batch)), non_blocking=True)

    for _ in range(max(warmup, 2)):
        dev_step()
    torch.cuda.synchronize()
    ms = cx.timed(dev_step, steps)
    for _ in range(2):
        api_step()
    ms_e2e = cx.timed(api_step, steps)
    # the recurrent part alone: the forward + backward program of the last get_loss call
    prog = next(iter(lc._train_programs.values()))[0]
    native = prog.plan.compile()
    for _ in range(2):
        native.run()
    ms_prog = cx.timed(lambda: native.run(), 5) / 5
    # deployment tick (lstm_step_controller.py:232-286): T = 1 step per control tick, state carried on the device
    cond = torch.zeros(1, 256, device=cx.dev)
    vla1, f1 = torch.zeros(1, A, device=cx.dev), torch.zeros(1, F, device=cx.dev)
    with torch.no_grad():
        lc.eval()
        lc.predict(cond, vla1, f1, initialize=True)
        torch.cuda.synchronize()
        lat = []
        for _ in range(50):
            t0 = time.perf_counter()
            lc.predict(cond, vla1, f1)
            torch.cuda.synchronize()
            lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    sps = cx.world * batch * steps / (ms * 1e-3)
    n_par = sum(p.numel() for m in lc.trainable_modules for p in m.parameters())
    return {"metric": "training sequences/sec (lstm_train.py step)", "value": sps, "unit": "sequences/s", "ms_per_step": ms / steps,
            "seq_len": T, "batch_per_gpu": batch, "global_batch": batch * cx.world, "scaling": "weak", "dtype": "bf16",
            "e2e": {"value": cx.world * batch * steps / (ms_e2e * 1e-3), "unit": "sequences/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": sum(v.numel() * v.element_size() for v in host_batch.values()), "d2h_bytes_per_step": 4},
            "lstm_forward_backward_program_ms": ms_prog, "lstm_steps_per_s": batch * T / (ms_prog * 1e-3),
            "collective": None if cx.world == 1 else {"op": "NCCL all-reduce (SUM) of the flat fp32 gradient buffer", "ranks": cx.world,
                                                      "bytes": n_par * 4},
            "predict_tick_ms_p50": lat[len(lat) // 2], "loss_first": float(losses[0]), "loss_last": float(losses[-1])}


def bench_lstm_train_cached(cx, wl, batch, T, steps, warmup, episodes=4):
    """lstm_train.py:122-139 fed from the HBM episode store with the DinoV2 feature cache: per step the host sends
    `batch` sample numbers; obs_encoder consumes the cached features, the frozen encoder does not run.  Synthetic episodes in the
    reference's schema with `T`-step VLA chunks (A = 10, 3 force dims), `episodes` x (T + 2 + batch / episodes + 8) frames."""
    import contextlib
    import shutil
    import tempfile
    torch = cx.torch
    from vla_touch_b200 import controller_dataset as cd
    from vla_touch_b200 import episode_store as es
    from vla_touch_b200.synthetic import synth_episode
    from vla_touch_b200.trainer import LSTMControllerTrainer
    name, hidden, heads, layers, hw, _, _, _, n_steps, _ = wl
    wl10 = (name, hidden, heads, layers, hw, T, 10, 3, n_steps, batch)
    frames = T + 2 + -(-batch // episodes) + 8
    td = tempfile.mkdtemp(prefix="vtep_")
    try:
        for e in range(episodes):
            es.write_episode_shard(synth_episode(2000 * (cx.rank + 1) + e, frames, hw, vla_T=T, still_frames=2), os.path.join(td, f"episode_{e}.vtep"))
        with contextlib.redirect_stdout(sys.stderr):
            ds = cd.ControllerDataset(td, context_frames=2, horizon=T, use_images=True, image_size=hw)
        lc = make_lstm_controller(cx, wl10)
        t0 = time.perf_counter()
        store = ds.device_store(cx.dev, image_encoder=lc.image_encoder, feature_chunk=128)
        torch.cuda.synchronize()
        fill_s = time.perf_counter() - t0
    finally:
        shutil.rmtree(td, ignore_errors=True)
    tr = LSTMControllerTrainer(lc, ds.stats, device=cx.dev)
    sampler = cd.EpisodeBatchSampler(len(ds), batch, 0, 1, seed=cx.rank)
    epoch = [0]

    def next_indices():
        sampler.set_epoch(epoch[0])
        epoch[0] += 1
        return next(iter(sampler))

    fixed = torch.from_numpy(next_indices()).to(cx.dev)
    losses = []
    loss_host = torch.empty((), dtype=torch.float32).pin_memory()

    def dev_step():
        losses.append(tr.train_step(store.gather(fixed, with_displacements=False)))

    def api_step():
        loss_host.copy_(tr.train_step(store.gather(next_indices(), with_displacements=False)), non_blocking=True)

    for _ in range(max(warmup, 2)):
        dev_step()
    torch.cuda.synchronize()
    ms = cx.timed(dev_step, steps)
    for _ in range(2):
        api_step()
    ms_e2e = cx.timed(api_step, steps)
    sps = cx.world * batch * steps / (ms * 1e-3)
    return {"metric": "training sequences/sec (lstm_train.py step fed from the HBM episode store, DinoV2 features cached)", "value": sps,
            "unit": "sequences/s", "ms_per_step": ms / steps, "seq_len": T, "batch_per_gpu": batch, "global_batch": batch * cx.world,
            "scaling": "weak", "dtype": "bf16",
            "e2e": {"value": cx.world * batch * steps / (ms_e2e * 1e-3), "unit": "sequences/s", "ms_per_step": ms_e2e / steps,
                    "h2d_bytes_per_step": batch * 8, "d2h_bytes_per_step": 4},
            "store": {"episodes": episodes, "frames": store.frames, "samples": len(store), "fill_seconds_incl_feature_cache": fill_s},
            "loss_first": float(losses[0]), "loss_last": float(losses[-1])}


def bench_latency(cx, wl, batch, calls):
    """BASELINE configs[4]: RDT stub (random chunks) + refine, p50 latency of one predict() call with host inputs (wall clock around
    the call + synchronize, per rank; the reported p50 is the max over ranks)."""
    torch = cx.torch
    ctl, eng, api_step, h2d, d2h = predict_harness(cx, wl, batch)
    for _ in range(3):
        api_step()
    torch.cuda.synchronize()
    lat = []
    for _ in range(calls):
        cx.barrier()
        t0 = time.perf_counter()
        api_step()
        torch.cuda.synchronize()
        lat.append((time.perf_counter() - t0) * 1e3)
    lat.sort()
    p = torch.tensor([lat[len(lat) // 2], lat[int(len(lat) * 0.9)]], device=cx.dev)
    if cx.world > 1:
        cx.dist.all_reduce(p, op=cx.dist.ReduceOp.MAX)
    return {"metric": "predict() latency per call, p50", "value": float(p[0]), "unit": "ms", "p90_ms": float(p[1]), "higher_is_better": False,
            "batch_per_gpu": batch, "global_batch": batch * cx.world, "calls": calls, "chunks_per_s": cx.world * batch / (float(p[0]) * 1e-3),
            "h2d_bytes_per_call": h2d, "d2h_bytes_per_call": d2h}


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg1", "cfg2_train", "cfg2_train_cached", "cfg4", "cfg4_cached", "cfg5"])
    ap.add_argument("--batch", type=int, default=0, help="rows per GPU (default: the workload's)")
    ap.add_argument("--ref-batch", type=int, default=32)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--only-headline", action="store_true", help="skip the secondary workloads (cfg2_train, cfg3, cfg4, cfg5)")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)
    base = "cfg2" if args.workload in ("cfg2_train", "cfg2_train_cached", "cfg4", "cfg4_cached") else args.workload
    wl = WORKLOADS[base]
    if args.impl == "reference":
        args.steps = min(args.steps, 4)
        return run_reference(args, wl)

    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"      # the version banner goes to stdout and would precede the JSON line
    # stdout carries exactly ONE JSON line: NCCL writes its version banner to fd 1 when the first communicator is created,
    # so fd 1 points at stderr until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    cx = Ctx(args)
    from vla_touch_b200 import native as nv
    nv.require_b200()
    torch, world, rank = cx.torch, cx.world, cx.rank
    name, hidden, heads, layers, hw, T, A, F, steps, batch = wl
    if args.workload in ("cfg3", "cfg5"):
        batch = max(1, batch // world)          # global batch fixed by BASELINE: strong scaling
    batch = args.batch or batch
    strong = args.workload in ("cfg3", "cfg5")

    def emit(line):
        if rank == 0:
            sys.stdout.flush()
            os.dup2(saved_stdout, 1)
            print(json.dumps(line), flush=True)
        if world > 1:
            cx.dist.destroy_process_group()

    sampler = ClockSampler(cx.local)
    sampler.start()
    # ---- single secondary workloads selected explicitly ----
    if args.workload == "cfg2_train":
        r = bench_train(cx, wl, batch, args.steps, args.warmup)
        sampler.stop_flag = True
        r.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
                  "data": "synthetic", "config": {"workload": "cfg2_train: bridge_train.py step, dinov2-small 224x224 x2 cams, T=64, A=7, F=64"},
                  "clocks": sampler.summary(), "gpu_launches": r["gpu_launches_per_step"] * args.steps})
        return emit(r)
    if args.workload == "cfg2_train_cached":
        r = bench_train_cached(cx, wl, batch, args.steps, args.warmup)
        sampler.stop_flag = True
        r.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
                  "data": "synthetic", "config": {"workload": "cfg2_train_cached: bridge_train.py step from the HBM episode store, T=64, A=10, F=3, DinoV2-S features of 2 x 224x224 cameras cached"},
                  "clocks": sampler.summary()})
        return emit(r)
    if args.workload == "cfg4_cached":
        r = bench_lstm_train_cached(cx, wl, args.batch or 512, 128, args.steps, args.warmup)
        sampler.stop_flag = True
        r.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
                  "data": "synthetic", "config": {"workload": "cfg4_cached: lstm_train.py step from the HBM episode store, seq_len 128, A=10, F=3, DinoV2-S features of 2 x 224x224 cameras cached"},
                  "clocks": sampler.summary()})
        return emit(r)
    if args.workload == "cfg4":
        r = bench_lstm_train(cx, wl, args.batch or 512, 128, args.steps, args.warmup)
        sampler.stop_flag = True
        r.update({"n_gpus": world, "steps": args.steps, "warmup": args.warmup, "higher_is_better": True, "vs_baseline": None,
                  "data": "synthetic", "config": {"workload": "cfg4: lstm_train.py step, seq_len 128, dinov2-small 224x224 x2 cams, A=7, F=64"},
                  "clocks": sampler.summary()})
        return emit(r)
    if args.workload == "cfg5":
        r = bench_latency(cx, wl, batch, max(args.steps, 20))
        sampler.stop_flag = True
        r.update({"n_gpus": world, "steps": max(args.steps, 20), "warmup": 3, "vs_baseline": None, "data": "synthetic", "scaling": "strong",
                  "dtype": "bf16", "config": {"workload": "cfg5: random base chunks + predict(), global batch 64"}, "clocks": sampler.summary()})
        return emit(r)

    # ---- headline: predict ----
    res = bench_predict(cx, wl, batch, args.steps, args.warmup)
    sampler.stop_flag = True
    eng, ms, ms_e2e = res["eng"], res["ms"], res["ms_e2e"]
    value = world * batch * args.steps / (ms * 1e-3)
    e2e = world * batch * args.steps / (ms_e2e * 1e-3)
    line = {
        "metric": "refined action-chunks/sec", "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "strong" if strong else "weak",
        "vs_baseline": None, "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: DiffusionController.predict, {name} ({layers} layers), {hw}x{hw} uint8 x2 cams, "
                               f"T={T}, A={A}, F={F}, {eng.n_steps} SDE steps, batch {batch}/GPU",
                   "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"dp{world} (no collective on the inference path)",
                   "l2": "no flush: per-step working set (activations ~1.3 GB at batch 256) exceeds the 126 MB L2",
                   "weights": "seeded synthetic (no network)", "noise": "in-kernel Philox"},
        "e2e": {"value": e2e, "unit": "chunks/s", "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": res["d2h"],
                "ms_per_step": ms_e2e / args.steps,
                # the upload alone, all ranks at once, max over ranks: with N GPUs behind shared host memory / PCIe switches this is
                # what grows with N and what the e2e figure loses against `value` (the upload of call i+1 overlaps the kernels of
                # call i only while it is shorter than the step)
                "h2d_ms_alone": res["ms_h2d"] / args.steps, "h2d_gbs_per_gpu": res["h2d"] / (res["ms_h2d"] / args.steps * 1e-3) / 1e9},
        "gpu_launches": res["launches"] * args.steps, "launches_per_step": res["launches"],
        "clocks": sampler.summary(),
    }
    if rank == 0:
        line["roofline"] = roofline_block(cx, eng, wl, args.workload, value, world)
    cx.barrier()
    del res, eng
    torch.cuda.empty_cache()
    # ---- the other BASELINE configs, same process, same GPUs ----
    if not args.only_headline and args.workload == "cfg2":
        k = max(3, min(args.steps, 8))
        extra = {}

        def attempt(key, fn):
            try:
                extra[key] = fn()
            except Exception as e:               # a secondary workload must not take the headline down with it
                extra[key] = {"error": f"{type(e).__name__}: {e}"[:300]}
            torch.cuda.synchronize()
            torch.cuda.empty_cache()

        attempt("cfg2_train", lambda: bench_train(cx, wl, 256, k, 3))
        attempt("cfg2_train_cached", lambda: bench_train_cached(cx, wl, 256, k, 3))
        attempt("cfg4_lstm_train_cached", lambda: bench_lstm_train_cached(cx, wl, 512, 128, k, 3))
        attempt("cfg4_lstm_train", lambda: bench_lstm_train(cx, wl, 512, 128, k, 3))
        wl3 = WORKLOADS["cfg3"]

        def cfg3():
            b3 = max(1, wl3[9] // world)
            r3 = bench_predict(cx, wl3, b3, 3, 3)
            v = world * b3 * 3 / (r3["ms"] * 1e-3)
            f3 = flops_per_chunk(wl3[1], wl3[3], wl3[4], wl3[5], wl3[6], wl3[7], wl3[8])
            pk, _ = peaks()
            return {"metric": "refined action-chunks/sec", "value": v, "unit": "chunks/s", "ms_per_step": r3["ms"] / 3, "scaling": "strong",
                    "batch_per_gpu": b3, "global_batch": b3 * world, "sde_steps": 50, "model": wl3[0],
                    "e2e": {"value": world * b3 * 3 / (r3["ms_e2e"] * 1e-3), "unit": "chunks/s", "h2d_bytes_per_step": r3["h2d"],
                            "d2h_bytes_per_step": r3["d2h"]},
                    "gflop_per_chunk": f3 / 1e9, "algorithmic_tflops_per_gpu": v / world * f3 / 1e12,
                    "frac_of_sustained": v / world * f3 / 1e12 / pk.get("bf16_tflops_sustained", 1400.0)}
        attempt("cfg3", cfg3)
        attempt("cfg5_latency", lambda: bench_latency(cx, WORKLOADS["cfg5"], max(1, 64 // world), 20))
        if world > 1:
            def strong_headline():
                bs = max(1, 256 // world)
                r = bench_predict(cx, wl, bs, k, 3)
                return {"metric": "refined action-chunks/sec", "value": world * bs * k / (r["ms"] * 1e-3), "unit": "chunks/s",
                        "ms_per_step": r["ms"] / k, "scaling": "strong", "batch_per_gpu": bs, "global_batch": bs * world,
                        "e2e": {"value": world * bs * k / (r["ms_e2e"] * 1e-3), "unit": "chunks/s"}}
            attempt("cfg2_strong", strong_headline)
        line["workloads"] = extra
    if not args.no_cpu_baseline and world == 1 and rank == 0:
        line["cpu_baseline"] = cpu_baseline(wl)
    emit(line)


if __name__ == "__main__":
    main()
