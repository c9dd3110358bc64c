#!/usr/bin/env python
"""Benchmark of the VLA-Touch action-refinement hot path on B200 (contract: see DESIGN.md 'Measurement').

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl reference] [--workload cfg2|cfg3|cfg1]

metric : refined action-chunks/s (one chunk = one DiffusionController.predict row: 2 camera images + state + tactile
         + base chunk -> refined [T, A] chunk), whole-job aggregate over N GPUs (weak scaling: `batch` rows per GPU).
value  : device-resident inputs, K CUDA-graph launches of the whole predict program, CUDA events, max over ranks.
e2e    : the same metric through the public API (DiffusionController.predict) with pinned HOST inputs, H2D copies and
         the D2H read of the refined chunks inside the timed region.
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
    "cfg3": ("facebook/dinov2-base", 768, 12, 12, 224, 64, 7, 64, 50, 128),
    "cfg1": ("facebook/dinov2-small", 384, 6, 12, 384, 16, 10, 3, 10, 1),
}


def flops_per_chunk(hidden, layers, hw, T, A, F, steps):
    """F = 2 F_dino + F_enc + n * 2 * F_unet(T)."""
    n_tok = (hw // 14) ** 2 + 1
    f_dino = 2 * (n_tok - 1) * 588 * hidden + layers * (24 * n_tok * hidden ** 2 + 4 * n_tok ** 2 * hidden)
    f_enc = 2 * ((2 * hidden + A + F) * 256 + 2 * 256 ** 2)
    f_unet = 0.0115e9 + 0.02008e9 * T
    return 2 * f_dino + f_enc + steps * 2 * f_unet


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


def run_reference(args, wl):
    """--impl reference: the oracle port of the reference CPU path, all host threads, bounded sample per step."""
    import torch
    from oracle import vt_oracle as orc
    from vla_touch_b200 import synthetic as syn
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    name, hidden, heads, layers, hw, T, A, F, steps, batch = wl
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bs = args.ref_batch
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


def cpu_baseline(wl, seconds=15.0):
    import torch
    from oracle import vt_oracle as orc
    from vla_touch_b200 import synthetic as syn
    name, hidden, heads, layers, hw, T, A, F, steps, batch = wl
    cores = os.cpu_count()
    torch.set_num_threads(cores)
    Bs = 4
    dino, enc, net = synth_weights(hidden, layers, A, F)
    sub = lambda p: {k[len(p):]: v for k, v in net.items() if k.startswith(p)}
    inp = syn.synth_predict_inputs(Bs, T, A, F, hw, 1234)
    noise = syn.det_normal("bench.noise", (steps, Bs, T, A), 1)
    f = lambda: orc.predict(dino, enc, sub("v_net."), sub("s_net."), syn.synth_stats(A), heads, inp["state"], inp["vla_actions"],
                            inp["images_cam1"][:, None], inp["images_cam2"][:, None], inp["forces"], steps, 0.03, noise)
    with torch.no_grad():
        f()
        n, t0 = 0, time.perf_counter()
        while time.perf_counter() - t0 < seconds:
            f()
            n += 1
    dt = time.perf_counter() - t0
    return {"value": Bs * n / dt, "unit": "chunks/s", "cores": cores, "kind": "port",
            "sample": f"{n} predict() calls of {Bs} rows in {dt:.1f}s (oracle/vt_oracle.py restatement, torch CPU fp32, {cores} threads)"}


def gemm_flops(d):
    """FLOPs of one tensor-core launch: an implicit GEMM, or the fused ViT MLP (two GEMMs rows x D x 4D)."""
    if hasattr(d, "w2"):                       # MlpDesc
        return 2.0 * 2.0 * d.rows * d.D * 4 * d.D
    return 2.0 * d.G * d.M * d.N * d.taps * d.kc * d.passes


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200")
    ap.add_argument("--workload", default="cfg2", choices=list(WORKLOADS))
    ap.add_argument("--batch", type=int, default=0, help="rows per GPU (default: the workload's)")
    ap.add_argument("--ref-batch", type=int, default=4)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl != "reference" else max(args.warmup, 1)
    wl = WORKLOADS[args.workload]
    if args.impl == "reference":
        args.steps = min(args.steps, 5)
        return run_reference(args, wl)

    if os.environ.get("NCCL_DEBUG", "").upper() == "VERSION":
        os.environ["NCCL_DEBUG"] = "WARN"      # the version banner goes to stdout and would precede the JSON line
    import torch
    import torch.distributed as dist
    from vla_touch_b200 import native as nv
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.bridge_controller import DiffusionController

    name, hidden, heads, layers, hw, T, A, F, steps, batch = wl
    batch = args.batch or batch
    rank, world = int(os.environ.get("RANK", "0")), int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    # stdout carries exactly ONE JSON line: NCCL writes its version banner to fd 1 when the first communicator is created,
    # so fd 1 points at stderr until the line is printed
    sys.stdout.flush()
    saved_stdout = os.dup(1)
    os.dup2(2, 1)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    nv.require_b200()
    dev = f"cuda:{local}"

    # ---- controller with seeded synthetic weights (no network: no checkpoints) ----
    dino_sd, enc_sd, net_sd = synth_weights(hidden, layers, A, F)
    model_args = {'interpolant_type': 'linear', 'gamma_type': '2^0.5*t(t-1)', 'epsilon_type': '1-t', 'prior_policy': 'vla',
                  'beta_max': 0.03, 'sde_type': 'vs', 'action_dim': A, 'obs_dim': 256, 'obs_horizon': 1, 'net_type': 'unet1D_si',
                  'pretrain': False, 'context_frames': 2, 'horizon': T}
    ctl = DiffusionController(state_dim=A, hidden_dim=256, image_model_path=name, diffusion_steps=steps, device=dev,
                              model_args=model_args, use_force=True, force_dim=F, image_state_dict=dino_sd)
    ctl.state_encoder.load_state_dict(enc_sd)
    ctl.diffusion_model.net.load_state_dict(net_sd)
    ctl.diffusion_model.ema = type(ctl.diffusion_model.ema)(ctl.diffusion_model.net.parameters(), decay=0.75)
    ctl.stats = {k: v.to(dev) for k, v in syn.synth_stats(A).items()}

    inp = syn.synth_predict_inputs(batch, T, A, F, hw, 1234 + rank)
    host = {k: v.pin_memory() for k, v in inp.items()}
    host["images_cam1"] = inp["images_cam1"][:, None].contiguous().pin_memory()   # deployment layout [B,1,H,W,3] uint8
    host["images_cam2"] = inp["images_cam2"][:, None].contiguous().pin_memory()
    h2d = sum(host[k].numel() * host[k].element_size() for k in ("state", "forces", "vla_actions", "images_cam1", "images_cam2"))
    out_host = torch.empty(batch, T, A, dtype=torch.float32).pin_memory()

    def api_step():
        out = ctl.predict(host["state"], host["vla_actions"], host["images_cam1"], host["images_cam2"], host["forces"])
        out_host.copy_(out, non_blocking=True)

    # first call builds the engine, the FiLM time tables and the CUDA graph
    api_step()
    torch.cuda.synchronize()
    eng = next(iter(ctl._engines.values()))
    launches = eng.num_launches()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, n):
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        ev0.record()
        for _ in range(n):
            fn()
        ev1.record()
        barrier()
        ms = torch.tensor([ev0.elapsed_time(ev1)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms)

    # ---- value: device-resident inputs, graph replays ----
    for _ in range(args.warmup):
        eng.run_predict(graph=True)
    sampler = ClockSampler(local)
    sampler.start()
    ms = timed(lambda: eng.run_predict(graph=True), args.steps)
    sampler.stop_flag = True
    value = world * batch * args.steps / (ms * 1e-3)
    # ---- e2e: public API, pinned host inputs, H2D + D2H inside the timed region ----
    for _ in range(args.warmup):
        api_step()
    ms_e2e = timed(api_step, args.steps)
    e2e = world * batch * args.steps / (ms_e2e * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    # ---- roofline of the dominant kernel: the GEMM launch with the most FLOPs, timed alone on this stream ----
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak_tf = peaks.get("bf16_tflops", 1590.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops (burst, kernel timed alone)" if peaks else "fallback 1590 (B200_PROFILING.md)"
    prog = eng.plan.compile()
    a0, b1 = eng.predict_range()
    gemms = [(gemm_flops(d), i) for i, d in enumerate(eng.plan.descs) if isinstance(d, (nv.GemmDesc, nv.MlpDesc)) and a0 <= i < b1]
    total_gemm_flops = sum(f for f, _ in gemms)
    fl, idx = max(gemms)
    for _ in range(5):
        prog.run(idx, 1)
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    reps = 50
    ev0.record()
    for _ in range(reps):
        prog.run(idx, 1)
    ev1.record()
    torch.cuda.synchronize()
    k_ms = ev0.elapsed_time(ev1) / reps
    achieved = fl / (k_ms * 1e-3) / 1e12
    # DRAM traffic of that launch from the committed ncu --set full capture (profiles/ncu_traffic.json, bytes per launch)
    traffic, traffic_src = None, None
    try:
        table = json.load(open(os.path.join(ROOT, "profiles", "ncu_traffic.json")))
        for key, rec in table.get(args.workload, {}).items():
            if eng.plan.tags[idx].endswith(key):
                traffic, traffic_src = rec["dram_bytes"], rec["source"]
    except Exception:
        pass
    # the attention kernel of one ViT layer, timed the same way (BASELINE metric: DinoV2 attention TFLOP/s vs peak)
    attn = None
    att_ops = [i for i, d in enumerate(eng.plan.descs) if isinstance(d, nv.AttnDesc) and a0 <= i < b1]
    if att_ops:
        ai = att_ops[-1]
        d = eng.plan.descs[ai]
        for _ in range(5):
            prog.run(ai, 1)
        torch.cuda.synchronize()
        ev0.record()
        for _ in range(reps):
            prog.run(ai, 1)
        ev1.record()
        torch.cuda.synchronize()
        a_ms = ev0.elapsed_time(ev1) / reps
        a_fl = 4.0 * d.tokens * d.tokens * 64 * d.heads * d.images
        a_exp = float(d.tokens) * d.tokens * d.heads * d.images
        attn = {"kernel": f"attention [{eng.plan.tags[ai]}]", "ms_per_launch": a_ms, "flops_per_launch": a_fl,
                "achieved": a_fl / (a_ms * 1e-3) / 1e12, "unit": "TFLOP/s", "frac_of_tensor_peak": a_fl / (a_ms * 1e-3) / 1e12 / peak_tf,
                "exp_per_launch": a_exp,
                "note": "head_dim 64: one exponential per 256 tensor FLOPs; at 16 MUFU ex2 / clock / SM the softmax alone "
                        "needs 2x the cycles of the two MMAs, so the tensor pipe cannot exceed ~50 % in this kernel"}
    f_chunk = flops_per_chunk(hidden, layers, hw, T, A, F, eng.n_steps)
    step_tf = value / world * f_chunk / 1e12
    roofline = {"bound": "tensor",
                "kernel": f"{'mlp_fused_kernel' if isinstance(eng.plan.descs[idx], nv.MlpDesc) else 'gemm_tc_kernel'} [{eng.plan.tags[idx]}]", "achieved": achieved, "peak": peak_tf,
                "unit": "TFLOP/s", "frac": achieved / peak_tf, "traffic": traffic, "traffic_source": traffic_src,
                "peak_source": peak_src, "flops_per_launch": fl, "ms_per_launch": k_ms, "attention": attn,
                "whole_step": {"algorithmic_tflops_per_gpu": step_tf, "frac_of_sustained": step_tf / peaks.get("bf16_tflops_sustained", 1400.0),
                               "gflop_per_chunk": f_chunk / 1e9, "executed_gemm_gflop_per_chunk": total_gemm_flops / batch / 1e9}}
    line = {
        "metric": "refined action-chunks/sec", "value": value, "unit": "chunks/s", "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "bf16", "data": "synthetic",
        "config": {"workload": f"{args.workload}: DiffusionController.predict, {name} ({layers} layers), {hw}x{hw} uint8 x2 cams, "
                               f"T={T}, A={A}, F={F}, {eng.n_steps} SDE steps, batch {batch}/GPU",
                   "batch_per_gpu": batch, "global_batch": batch * world, "parallelism": f"dp{world} (no collective)",
                   "l2": "no flush: per-step working set (activations ~1.3 GB at batch 256) exceeds the 126 MB L2",
                   "weights": "seeded synthetic (no network)", "noise": "in-kernel Philox"},
        "e2e": {"value": e2e, "unit": "chunks/s", "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": out_host.numel() * 4,
                "ms_per_step": ms_e2e / args.steps},
        "gpu_launches": launches * args.steps, "launches_per_step": launches,
        "clocks": sampler.summary(), "roofline": roofline,
    }
    if not args.no_cpu_baseline and world == 1:
        line["cpu_baseline"] = cpu_baseline(wl)
    sys.stdout.flush()
    os.dup2(saved_stdout, 1)
    print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
