"""Developer tool (GPU box): fixed cost of one GEMM launch (tiny problem: one tile, one K atom) for the kernel variants,
stream-serialised back-to-back launches timed with CUDA events, with and without PDL / CTA pairs."""
import os, sys, subprocess
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if len(sys.argv) > 1 and sys.argv[1] == "child":
    sys.path.insert(0, ROOT)
    import torch
    from vla_touch_b200 import native as nv
    from vla_touch_b200.plan import Plan, linear_desc, pack_linear_weight
    dev = torch.device("cuda:0")
    for M, K, N, kind in ((256, 64, 256, "bf16"), (256, 64, 256, "f32res"), (256, 2560, 256, "bf16"), (37888, 64, 256, "bf16")):
        plan = Plan(dev)
        a = plan.buf("a", (M, K), torch.bfloat16)
        wp, n_pad, k_pad = pack_linear_weight(torch.randn(N, K), torch.bfloat16)
        w = plan.reg(wp.to(dev))
        bias = plan.reg(torch.zeros(N, device=dev))
        if kind == "f32res":
            out = plan.buf("out", (M, N), torch.float32)
            kw = dict(res=out, ldres=N)
        else:
            out = plan.buf("out", (M, N), torch.bfloat16)
            kw = {}
        for _ in range(8):
            plan.add(linear_desc(a=a, rows=M, k=k_pad, a_ld=K, w=w, n=N, n_pad=n_pad, w_ld=k_pad, out=out, ldc=N, bias=bias, **kw), "g")
        prog = plan.compile()
        prog.run(0, 8)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(50):
            prog.run(0, 8)
        e1.record()
        torch.cuda.synchronize()
        stream_us = e0.elapsed_time(e1) * 1e3 / 400
        prog.graph_build(0, 8)
        prog.graph_launch()
        torch.cuda.synchronize()
        e0.record()
        for _ in range(50):
            prog.graph_launch()
        e1.record()
        torch.cuda.synchronize()
        graph_us = e0.elapsed_time(e1) * 1e3 / 400
        print(f"  M={M:6d} K={K:5d} N={N} {kind:7s}: {stream_us:6.2f} us/launch (stream)  {graph_us:6.2f} us/launch (graph of 8)", flush=True)
else:
    for env in ({}, {"VT_PDL": "0"}, {"VT_GEMM_PAIR": "0"}, {"VT_GEMM_PAIR": "0", "VT_PDL": "0"}):
        print(env or "default (pairs + PDL)", flush=True)
        r = subprocess.run([sys.executable, os.path.abspath(__file__), "child"], env=dict(os.environ, **env), capture_output=True, text=True, timeout=300)
        print(r.stdout + r.stderr[-600:], flush=True)
