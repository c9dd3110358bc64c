"""Developer tool (GPU box): the training step of the bridge (get_loss forward + backward of b_net / v_net / s_net,
unet_train.LossBackwardProgram) at a given batch: CUDA-event time of the whole program (replayed, L2 flushed between
replays), training throughput in samples/s and algorithmic TFLOP/s, and a per-op breakdown (each op replayed alone).

    python tools/train_profile.py [batch=256] [T=64] [A=7]  > gpurun_out/train_profile.txt
    python tools/train_profile.py lstm [batch=512] [T=128] [A=7] [F=64]      (BASELINE configs[3]: the LSTM controller's step)

Algorithmic FLOPs per sample: 3 nets x (forward + backward = 3 x forward) x F_unet(T) with
F_unet(T) = 0.0115e9 + 0.02008e9 * T; the recomputed raw convolutions are NOT counted as useful work."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "tests")]
import torch  # noqa: E402

from vla_touch_b200 import native as nv  # noqa: E402
from vla_touch_b200 import shapes as shp  # noqa: E402
from vla_touch_b200 import synthetic as syn  # noqa: E402
from vla_touch_b200.params import sub_state_dict  # noqa: E402
from vla_touch_b200.unet_train import LossBackwardProgram  # noqa: E402


def gemm_flops(d) -> float:
    return 2.0 * d.G * d.M * d.N * d.taps * d.kc * d.passes if isinstance(d, nv.GemmDesc) else 0.0


def profile(plan, prog, header):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda:0")
    for _ in range(3):
        prog.run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prog.run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    out = [header(ms)]
    rows = []
    for i, d in enumerate(plan.descs):
        tt = []
        for _ in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); prog.run(i, 1); e1.record(); torch.cuda.synchronize()
            tt.append(e0.elapsed_time(e1) * 1e3)
        rows.append((i, plan.tags[i], type(d).__name__, sorted(tt)[1], gemm_flops(d)))
    tot = sum(r[3] for r in rows)
    for i, tag, kind, us, fl in rows:
        out.append(f"{i:4d} {tag:64s} {kind:14s} {us:9.1f} us {100 * us / tot:5.1f}%  {fl / us / 1e6 if fl else 0:8.1f} TFLOP/s")
    print("\n".join(out))


def main_lstm(argv):
    from vla_touch_b200.lstm_train import LstmLossBackwardProgram
    B, T, A, Fd = (int(argv[i]) if len(argv) > i else v for i, v in ((0, 512), (1, 128), (2, 7), (3, 64)))
    mods = {"force_encoder": syn.synth_state_dict(shp.mlp_shapes([Fd, 128, 128]), 41, "lstm.force_encoder."),
            "lstm": syn.synth_state_dict(shp.lstm_shapes(128 + A), 41, "lstm.lstm."),
            "output_head": syn.synth_state_dict(shp.lstm_head_shapes(256, A), 41, "lstm.output_head.")}
    lp = LstmLossBackwardProgram(mods, A, Fd, B, T, torch.device("cuda", 0))
    lp.set_inputs(syn.det_uniform("tp.vla", (B, T, A), 1, -1.0, 1.0), syn.det_normal("tp.f", (B, T, Fd), 1),
                  syn.det_normal("tp.cond", (B, 256), 1), syn.det_uniform("tp.exp", (B, T, A), 1, -1.0, 1.0))
    prog = lp.plan.compile()
    profile(lp.plan, prog, lambda ms: f"LSTM controller get_loss forward+backward, batch {B}, T {T}: {len(lp.plan)} ops, "
            f"{prog.num_launches()} launches, {ms:.3f} ms = {B / ms * 1e3:.0f} sequences/s, {B * T / ms * 1e3:.0f} steps/s; loss {lp.loss():.5f}")


def main():
    if len(sys.argv) > 1 and sys.argv[1] == "lstm":
        return main_lstm(sys.argv[2:])
    B = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    T = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    A = int(sys.argv[3]) if len(sys.argv) > 3 else 7
    dev = torch.device("cuda", 0)
    full = syn.synth_state_dict(shp.si_net_shapes(A, 256), 21, prefix="net.")
    lp = LossBackwardProgram([sub_state_dict(full, p) for p in ("b_net.", "v_net.", "s_net.")], A, B, T, 0.03, dev)
    lp.set_inputs(syn.det_uniform("tp.vla", (B, T, A), 1, -1.0, 1.0), syn.det_uniform("tp.exp", (B, T, A), 1, -1.0, 1.0),
                  syn.det_normal("tp.cond", (B, 256), 1), torch.rand(B, generator=torch.Generator().manual_seed(1)),
                  syn.det_normal("tp.z", (B, T, A), 2))
    prog = lp.plan.compile()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for _ in range(3):
        prog.run()
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); prog.run(); e1.record(); torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    ms = sorted(ts)[len(ts) // 2]
    f_unet = 0.0115e9 + 0.02008e9 * T
    useful = 3 * 3 * f_unet * B
    issued = sum(gemm_flops(d) for d in lp.plan.descs)
    mem = sum(t.numel() * t.element_size() for t in lp.plan._reg) / 2 ** 30
    out = [f"get_loss forward+backward, batch {B}, T {T}, A {A}: {len(lp.plan)} ops, {prog.num_launches()} launches, "
           f"{mem:.2f} GiB of plan tensors", f"whole program {ms:.3f} ms (median of 5, L2 flushed) = {B / ms * 1e3:.0f} samples/s; "
           f"useful {useful / ms / 1e9:.1f} TFLOP/s (algorithmic), issued GEMM work {issued / ms / 1e9:.1f} TFLOP/s; loss {lp.out.tolist()}"]
    rows = []
    for i, d in enumerate(lp.plan.descs):
        tt = []
        for _ in range(3):
            flush.zero_()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record(); prog.run(i, 1); e1.record(); torch.cuda.synchronize()
            tt.append(e0.elapsed_time(e1) * 1e3)
        rows.append((i, lp.plan.tags[i], type(d).__name__, sorted(tt)[1], gemm_flops(d)))
    tot = sum(r[3] for r in rows)
    by_kind = {}
    for _, tag, kind, us, _ in rows:
        k = kind if kind != "GemmDesc" else ("Gemm.wgrad" if "wgrad" in tag else "Gemm.dgrad" if "dgrad" in tag else
                                             "Gemm.recompute" if "recompute" in tag else "Gemm.forward")
        by_kind[k] = by_kind.get(k, 0.0) + us
    out.append("by kind: " + ", ".join(f"{k} {v:.0f} us ({100 * v / tot:.0f} %)" for k, v in sorted(by_kind.items(), key=lambda kv: -kv[1])))
    for i, tag, kind, us, fl in rows:
        out.append(f"{i:4d} {tag:64s} {kind:12s} {us:9.1f} us {100 * us / tot:5.1f}%  {fl / us / 1e6 if fl else 0:8.1f} TFLOP/s")
    print("\n".join(out))


if __name__ == "__main__":
    main()
