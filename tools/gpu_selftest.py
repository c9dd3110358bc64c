"""Developer tool (GPU box): compares every native kernel against the CPU plan interpreter, op by op.
    python tools/gpu_selftest.py            # runs each section in its own subprocess, log -> gpurun_out/selftest.log
    python tools/gpu_selftest.py <section>  # one section in-process
"""
import os
import subprocess
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

SECTIONS = ["gemm_bf16", "gemm_f32", "elem", "attn", "unet_bf16", "unet_f32", "engine_bf16", "engine_f32", "e2e"]


def cpu_resize(pos, s, nh, nw):
    import torch.nn.functional as F
    D = pos.shape[1]
    return F.interpolate(pos.reshape(1, s, s, D).permute(0, 3, 1, 2), size=(nh, nw), mode="bicubic",
                         align_corners=False).permute(0, 2, 3, 1).reshape(-1, D).contiguous()


def gemm_section(precise):
    import torch
    from vla_touch_b200 import native as nv
    from vla_touch_b200.plan import Plan, linear_desc
    from vla_touch_b200.unet import Mode
    import gpu_diff
    m = Mode(precise)
    cases = [(300, 384, 384, 128), (128, 128, 64, 128), (77, 7, 256, 32), (1000, 1152, 384, 128), (5, 256, 896, 128),
             (257, 384, 1536, 128)]
    for (M, N, K, bn) in cases:
        plans = []
        g = torch.Generator().manual_seed(M * 7 + N)
        a32 = torch.randn(M, K, generator=g)
        w32 = torch.randn(N, K, generator=g) / K ** 0.5
        bias = torch.randn(N, generator=g)
        ls = torch.randn(N, generator=g)
        res = torch.randn(M, N, generator=g)
        for dev in ("cpu", "cuda"):
            p = Plan(dev)
            n_pad = (N + bn - 1) // bn * bn
            a = p.buf("a", (M, m.ld(K)), m.tdt)
            w = p.buf("w", (n_pad, m.ld(K)), m.tdt)
            wp = torch.zeros(n_pad, K)
            wp[:N] = w32
            w.copy_(m.pack_w(wp))
            a.copy_(m.pack_w(a32))
            b = p.buf("bias", (n_pad,), torch.float32)
            b[:N] = bias
            cs = p.buf("ls", (N,), torch.float32)
            cs.copy_(ls)
            r = p.buf("res", (M, N), torch.float32)
            r.copy_(res)
            o1 = p.buf("o_plain", (M, N), torch.float32)
            o2 = p.buf("o_gelu", (M, m.ld(n_pad)), m.tdt)
            o3 = p.buf("o_res", (M, N), torch.float32)
            kw = dict(a=a, rows=M, k=K, a_ld=m.ld(K), w=w, n=N, n_pad=n_pad, w_ld=m.ld(K), bn=bn, passes=m.passes,
                      a_plane=m.plane(K), w_plane=K if precise else 0)
            p.add(linear_desc(out=o1, ldc=N, **kw), f"gemm {M}x{N}x{K} plain->f32")
            p.add(linear_desc(out=o2, ldc=m.ld(n_pad), bias=b, act=nv.ACT_GELU, out_plane=m.plane(n_pad), **kw),
                  f"gemm {M}x{N}x{K} bias+gelu->op")
            p.add(linear_desc(out=o3, ldc=N, bias=b, colscale=cs, res=r, ldres=N, **kw), f"gemm {M}x{N}x{K} bias+ls+res")
            plans.append(p)
        rows = gpu_diff.diff_plans(plans[0], plans[1])
        print(gpu_diff.format_rows(rows, 2e-2 if not precise else 1e-4), flush=True)
        ref = a32 @ w32.t()
        got = plans[1].bufs["o_plain"].cpu()
        print(f"     vs fp32 matmul: err {(got - ref).abs().max():.3e} (ref max {ref.abs().max():.2f})", flush=True)


def elem_section():
    import torch
    import vt_testutil as U
    import gpu_diff
    from vla_touch_b200 import native as nv
    from vla_touch_b200.dino import DinoWeights, DinoProgram
    from vla_touch_b200.plan import Plan
    # dino front end + LN on small shapes: imgstats/patchify/patch-embed/cls/LN with both image dtypes
    for kind, layers in (("u8bright5d", 1), ("f32bchw", 1), ("u8dark", 1)):
        sd = U.dino_sd(384, layers, 5)
        img = U.images_for(kind, "dino.img", 2, 56, 5)
        if img.dim() == 5:
            img = img[:, 0]
        layout = nv.LAYOUT_BCHW if kind == "f32bchw" else nv.LAYOUT_BHWC
        plans = []
        for dev in ("cpu", "cuda"):
            p = Plan(dev)
            dw = DinoWeights(sd, 6, dev, False)
            dp = DinoProgram(p, dw, 1, 2, 56, 56, img.dtype, layout, resize=cpu_resize if dev == "cpu" else
                             __import__("vla_touch_b200.dino", fromlist=["x"]).native_pos_resize)
            dp.img[0].copy_(img.contiguous())
            plans.append(p)
        # position-embedding resize kernel vs torch bicubic
        pos_c = [t for t in plans[0]._reg if t.shape == (17, 384)][0]
        pos_g = [t for t in plans[1]._reg if t.shape == (17, 384)][0]
        print(f"pos-embed bicubic resize: err {(pos_g.cpu() - pos_c).abs().max():.3e}")
        rows = gpu_diff.diff_plans(plans[0], plans[1], resync=True)
        print(kind)
        print(gpu_diff.format_rows(rows, 2e-2), flush=True)


def attn_section():
    import torch
    import gpu_diff
    from vla_touch_b200 import native as nv
    from vla_touch_b200.plan import Plan, ptr
    for (images, tokens, heads, dt) in ((2, 257, 6, torch.bfloat16), (1, 730, 6, torch.bfloat16), (3, 17, 2, torch.bfloat16),
                                        (1, 128, 1, torch.bfloat16), (2, 257, 6, torch.float32)):
        D = heads * 64
        g = torch.Generator().manual_seed(tokens)
        qkv32 = torch.randn(images * tokens, 3 * D, generator=g) * 1.5
        plans = []
        for dev in ("cpu", "cuda"):
            p = Plan(dev)
            qkv = p.buf("qkv", (images * tokens, 3 * D), dt)
            qkv.copy_(qkv32)
            ctx = p.buf("ctx", (images * tokens, D), dt)
            d = nv.AttnDesc()
            d.qkv, d.ctx, d.in_dtype, d.images, d.tokens, d.heads = ptr(qkv), ptr(ctx), nv.VT_BF16 if dt == torch.bfloat16 else nv.VT_F32, images, tokens, heads
            d.ctx_ld, d.ctx_plane = D, 0
            p.add(d, f"attention {images}x{tokens}x{heads} {dt}")
            plans.append(p)
        rows = gpu_diff.diff_plans(plans[0], plans[1])
        print(gpu_diff.format_rows(rows, 3e-2 if dt == torch.bfloat16 else 1e-4), flush=True)


def unet_section(precise):
    import torch
    import vt_testutil as U
    import gpu_diff
    from oracle import vt_oracle as orc
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.unet import UnetProgram
    for (A, T, B) in ((10, 16, 3), (7, 64, 5), (10, 48, 2)):
        v_sd, s_sd = U.net_sd(A, 21, "v_net"), U.net_sd(A, 21, "s_net")
        x = syn.det_uniform("unet.x", (B, T, A), 22, -1.0, 1.0)
        cond = syn.det_normal("unet.cond", (B, 256), 22)
        t = torch.linspace(0.001, 0.999, B)
        ups = []
        for dev in ("cpu", "cuda"):
            up = UnetProgram([v_sd, s_sd], A, B, T, dev, precise=precise)
            up.x.copy_(x); up.t.copy_(t); up.cond.copy_(cond)
            ups.append(up)
        rows = gpu_diff.diff_plans(ups[0].plan, ups[1].plan, resync=True)
        print(f"--- unet A={A} T={T} B={B} precise={precise} (per-op, resynced)")
        print(gpu_diff.format_rows(rows, 3e-2 if not precise else 2e-4), flush=True)
        out = ups[1](x.cuda(), t.cuda(), cond.cuda()).cpu()
        ref_v, ref_s = orc.unet_forward(v_sd, x, t, cond), orc.unet_forward(s_sd, x, t, cond)
        print(f"    end-to-end vs oracle: v err {(out[0] - ref_v).abs().max():.3e} s err {(out[1] - ref_s).abs().max():.3e} "
              f"(ref max {ref_v.abs().max():.2f})", flush=True)


def make_engine(c, dev, precise):
    from vla_touch_b200 import native as nv
    from vla_touch_b200.dino import DinoWeights, native_pos_resize
    from vla_touch_b200.engine import BridgeEngine
    dw = DinoWeights(c["dino"], c["heads"], dev, precise)
    imgs = [c["img1"], c["img2"]]
    if imgs[0].dim() == 5:
        imgs = [i[:, 0] for i in imgs]
    imgs = [i.contiguous() for i in imgs]
    eng = BridgeEngine(dino=dw, enc_sd=c["enc"], v_sd=c["v_ema"], s_sd=c["s_ema"], action_dim=c["A"], state_dim=c["A"],
                       force_dim=c["F"], use_force=True, B=c["B"], T=c["T"], H=c["hw"], W=c["hw"], img_dtype=imgs[0].dtype,
                       layout=nv.LAYOUT_BHWC, diffuse_step=c["steps"], device=dev, precise=precise,
                       resize=cpu_resize if dev == "cpu" else native_pos_resize, inject_noise=True)
    eng.dino_prog.img[0].copy_(imgs[0]); eng.dino_prog.img[1].copy_(imgs[1])
    eng.state.copy_(c["state"]); eng.forces.copy_(c["forces"]); eng.vla.copy_(c["vla"]); eng.noise.copy_(c["gold"]["noise"])
    eng.set_stats(c["stats"])
    return eng


def engine_section(precise):
    import torch
    import vt_testutil as U
    import gpu_diff
    c = U.predict_case("predict_cfg2_B3_dark_varstats")
    e_cpu, e_gpu = make_engine(c, "cpu", precise), make_engine(c, "cuda", precise)
    rows = gpu_diff.diff_plans(e_cpu.setup, e_gpu.setup, resync=True)
    print(gpu_diff.format_rows(rows, 3e-2 if not precise else 2e-4))
    a0, a1 = e_cpu.ranges["dino"][0], e_cpu.ranges["normalize"][1]
    b0, b1 = e_cpu.step_ranges[0][0], e_cpu.step_ranges[1][1]
    for name, t in e_cpu.plan.bufs.items():
        e_gpu.plan.bufs[name].copy_(t)
    rows = gpu_diff.diff_plans(e_cpu.plan, e_gpu.plan, a0, a1 - a0, sync_inputs=False, resync=True)
    rows += gpu_diff.diff_plans(e_cpu.plan, e_gpu.plan, b0, b1 - b0, sync_inputs=False, resync=True)
    print(gpu_diff.format_rows(rows, 3e-2 if not precise else 2e-4), flush=True)


def e2e_section():
    import torch
    import vt_testutil as U
    for tag in ("predict_cfg2_B3_dark_varstats", "predict_T48_f32_varstats", "predict_cfg2_B2", "predict_cfg1", "predict_cfg3_B1_base"):
        for precise in (False, True):
            c = U.predict_case(tag)
            t0 = time.time()
            eng = make_engine(c, "cuda", precise)
            for graph in (False, True):
                eng.run_predict(graph=graph)
                torch.cuda.synchronize()
                g = c["gold"]
                ce = (eng.cond.cpu() - g["cond"]).abs().max().item()
                oe = (eng.out.cpu() - g["out"]).abs().max().item()
                print(f"{tag:32s} {'f32x3' if precise else 'bf16 '} graph={int(graph)} cond err {ce:.3e} out err {oe:.3e} "
                      f"(|out| max {g['out'].abs().max():.2f}) launches {eng.num_launches()}  {time.time() - t0:.1f}s", flush=True)


def main():
    if len(sys.argv) > 1:
        sec = sys.argv[1]
        import torch
        torch.manual_seed(0)
        {"gemm_bf16": lambda: gemm_section(False), "gemm_f32": lambda: gemm_section(True), "elem": elem_section,
         "attn": attn_section, "unet_bf16": lambda: unet_section(False), "unet_f32": lambda: unet_section(True),
         "engine_bf16": lambda: engine_section(False), "engine_f32": lambda: engine_section(True), "e2e": e2e_section}[sec]()
        return
    os.makedirs(os.path.join(ROOT, "gpurun_out"), exist_ok=True)
    log = open(os.path.join(ROOT, "gpurun_out", "selftest.log"), "w")
    for sec in SECTIONS:
        t0 = time.time()
        try:
            r = subprocess.run([sys.executable, os.path.abspath(__file__), sec], capture_output=True, text=True, timeout=420)
            out, rc = r.stdout + "\n" + r.stderr[-3000:], r.returncode
        except subprocess.TimeoutExpired as e:
            out, rc = (e.stdout or b"").decode() + "\nTIMEOUT", -9
        msg = f"===== {sec}: rc={rc} ({time.time() - t0:.0f}s)\n{out}\n"
        log.write(msg)
        log.flush()
        print(msg[-2500:], flush=True)


if __name__ == "__main__":
    main()
