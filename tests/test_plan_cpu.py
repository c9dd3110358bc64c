"""CPU tests of the host-side logic: the plan descriptors (interpreted by tests/plan_emu.py) reproduce the oracle /
the reference golden vectors, the C ABI library loads and exports what include/vt_b200.h declares, ctypes mirrors
match the C structs, and the reference-API helpers behave like the reference."""
import ctypes
import os
import subprocess
import sys
import tempfile

import pytest
import torch
import torch.nn.functional as F

import plan_emu
import vt_testutil as U
from oracle import vt_oracle as orc
from vla_touch_b200 import native as nv
from vla_touch_b200 import schedule as sch
from vla_touch_b200 import shapes as shp
from vla_touch_b200 import synthetic as syn

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def cpu_resize(pos, s, nh, nw):
    D = pos.shape[1]
    return F.interpolate(pos.reshape(1, s, s, D).permute(0, 3, 1, 2), size=(nh, nw), mode="bicubic",
                         align_corners=False).permute(0, 2, 3, 1).reshape(-1, D).contiguous()


def test_library_loads_and_exports_every_declared_symbol():
    L = nv.lib()
    header = open(os.path.join(ROOT, "include", "vt_b200.h")).read()
    import re
    declared = set(re.findall(r"\b(vt_[a-z_0-9]+)\s*\(", header))
    assert declared == set(nv.EXPORTS), declared ^ set(nv.EXPORTS)
    for name in declared:
        assert hasattr(L, name), name
    assert L.vt_abi_version() == nv.ABI_VERSION


def test_no_device_is_an_error_not_a_fallback():
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(nv.NativeError):
        nv.device_info()
    from vla_touch_b200.plan import Plan
    with pytest.raises(nv.NativeError):
        Plan("cpu").compile()
    from vla_touch_b200.controller_dataset import normalize_actions
    with pytest.raises(nv.NativeError):
        normalize_actions(torch.zeros(1, 2, 3), syn.synth_stats(3), "vla")


def test_ctypes_structs_match_the_c_header():
    descs = {"vt_gemm_desc": nv.GemmDesc, "vt_ln_desc": nv.LnDesc, "vt_attn_desc": nv.AttnDesc,
             "vt_imgstats_desc": nv.ImgStatsDesc, "vt_patchify_desc": nv.PatchifyDesc, "vt_cls_desc": nv.ClsDesc,
             "vt_pack_desc": nv.PackDesc, "vt_affine_desc": nv.AffineDesc, "vt_tembed_desc": nv.TembedDesc,
             "vt_sde_desc": nv.SdeDesc, "vt_lstm_desc": nv.LstmDesc, "vt_qsample_desc": nv.QsampleDesc,
             "vt_siloss_desc": nv.SilossDesc, "vt_opt_tensor": nv.OptTensor, "vt_adamw_desc": nv.AdamwDesc,
             "vt_mlp_desc": nv.MlpDesc, "vt_rowproj_desc": nv.RowprojDesc, "vt_tcol_desc": nv.TcolDesc, "vt_gnbwd_desc": nv.GnbwdDesc,
             "vt_colsum_desc": nv.ColsumDesc, "vt_ewise_desc": nv.EwiseDesc, "vt_silossbwd_desc": nv.SilossBwdDesc, "vt_lstm_train_desc": nv.LstmTrainDesc,
             "vt_lstm_bwd_desc": nv.LstmBwdDesc, "vt_lngelubwd_desc": nv.LnGeluBwdDesc, "vt_dropmask_desc": nv.DropmaskDesc,
             "vt_persist_desc": nv.PersistDesc, "vt_wgrad_desc": nv.WgradDesc, "vt_batch_gather_desc": nv.BatchGatherDesc}
    lines = ['#include <stdio.h>', '#include <stddef.h>', f'#include "{ROOT}/include/vt_b200.h"', 'int main(void){']
    probes = []
    for cname, cls in descs.items():
        lines.append(f'printf("%zu\\n", sizeof({cname}));')
        probes.append((cname, None, ctypes.sizeof(cls)))
        for fname, _ in cls._fields_:
            lines.append(f'printf("%zu\\n", offsetof({cname}, {fname}));')
            probes.append((cname, fname, getattr(cls, fname).offset))
    lines.append('return 0;}')
    with tempfile.TemporaryDirectory() as td:
        src, exe = os.path.join(td, "p.c"), os.path.join(td, "p")
        open(src, "w").write("\n".join(lines))
        subprocess.run(["gcc", "-o", exe, src], check=True)
        got = [int(x) for x in subprocess.run([exe], capture_output=True, text=True, check=True).stdout.split()]
    assert len(got) == len(probes)
    for (cname, fname, want), g in zip(probes, got):
        assert g == want, (cname, fname, g, want)


def test_schedule_matches_oracle_and_reference_quirks():
    for ds in (10, 50, 93, 99, 7):
        n, dt, ts = sch.sde_schedule(ds)
        n2, dt2, ts2 = orc.sde_schedule(ds)
        assert n == n2 and dt == dt2 and all(torch.equal(a, b) for a, b in zip(ts, ts2))
        for t in ts:
            got = sch.sde_coefficients(t, dt)
            ref = orc.sde_coefficients(t, dt, 0.03)
            assert got == tuple(float(x) for x in ref)
    assert sch.sde_schedule(93)[0] == 92
    with pytest.raises(NotImplementedError):
        sch.check_model_args({"gamma_type": "(2t(t-1))^0.5"})
    sch.check_model_args({"sde_type": "bs"})                   # sde_bs (bridge_model.py:281-332) = the 'vs' update with b_net as drift
    assert sch.sde_coefficients(sch.sde_schedule(10)[2][3], 0.1, "bs")[1] == 0.0
    with pytest.raises(NotImplementedError):
        sch.check_model_args({"sde_type": "xs"})


@pytest.mark.parametrize("precise,A,T,tol", [(True, 7, 64, 5e-5), (True, 10, 48, 5e-5), (False, 10, 16, 8e-2)])
def test_unet_plan_reproduces_the_oracle(precise, A, T, tol):
    from vla_touch_b200.unet import UnetProgram
    B = 3
    v_sd, s_sd = U.net_sd(A, 21, "v_net"), U.net_sd(A, 21, "s_net")
    x = syn.det_uniform("unet.x", (B, T, A), 22, -1.0, 1.0)
    cond = syn.det_normal("unet.cond", (B, 256), 22)
    t = torch.tensor([0.3, 0.001, 0.999])
    up = UnetProgram([v_sd, s_sd], A, B, T, "cpu", precise=precise)
    up.x.copy_(x); up.t.copy_(t); up.cond.copy_(cond)
    plan_emu.run(up.plan)
    if T in (16, 32, 48, 64):
        g = U.golden(f"unet_A{A}_T{T}")
        assert (up.bufs.out[0] - g["v"]).abs().max() <= tol          # reference golden vector (v_net, per-sample t)
    assert (up.bufs.out[0] - orc.unet_forward(v_sd, x, t, cond)).abs().max() <= tol
    assert (up.bufs.out[1] - orc.unet_forward(s_sd, x, t, cond)).abs().max() <= tol
    assert len(up.plan) == 43 and sum(isinstance(d, nv.GemmDesc) for d in up.plan.descs) == 39


def _engine_for(c, precise):
    from vla_touch_b200.dino import DinoWeights
    from vla_touch_b200.engine import BridgeEngine
    dw = DinoWeights(c["dino"], c["heads"], "cpu", precise)
    imgs = [c["img1"], c["img2"]]
    if imgs[0].dim() == 5:
        imgs = [i[:, 0] for i in imgs]
    imgs = [i.contiguous() for i in imgs]
    eng = BridgeEngine(dino=dw, enc_sd=c["enc"], v_sd=c["v_ema"], s_sd=c["s_ema"], action_dim=c["A"], state_dim=c["A"],
                       force_dim=c["F"], use_force=True, B=c["B"], T=c["T"], H=c["hw"], W=c["hw"], img_dtype=imgs[0].dtype,
                       layout=nv.LAYOUT_BHWC, diffuse_step=c["steps"], device="cpu", precise=precise, resize=cpu_resize,
                       inject_noise=True)
    eng.dino_prog.img[0].copy_(imgs[0]); eng.dino_prog.img[1].copy_(imgs[1])
    eng.state.copy_(c["state"]); eng.forces.copy_(c["forces"]); eng.vla.copy_(c["vla"]); eng.noise.copy_(c["gold"]["noise"])
    eng.set_stats(c["stats"])
    return eng


@pytest.mark.parametrize("tag,precise", [("predict_cfg2_B3_dark_varstats", True), ("predict_T48_f32_varstats", True),
                                         ("predict_cfg2_B3_dark_varstats", False)])
def test_predict_plan_reproduces_the_reference_golden(tag, precise):
    c = U.predict_case(tag)
    eng = _engine_for(c, precise)
    plan_emu.run(eng.setup)
    a0, a1 = eng.ranges["dino"][0], eng.ranges["normalize"][1]
    plan_emu.run(eng.plan, a0, a1 - a0)
    b0, b1 = eng.sampler_range[0], eng.ranges["denormalize"][1]      # bf16: the persistent descriptor, interpreted layer by layer
    plan_emu.run(eng.plan, b0, b1 - b0)
    g = c["gold"]
    scale = float(g["out"].abs().max())
    if precise:
        assert (eng.cond - g["cond"]).abs().max() <= 1e-4
        assert (eng.out - g["out"]).abs().max() <= 1e-3            # north-star fp32 gate: 1e-3 abs
    else:
        assert (eng.out - g["out"]).abs().max() <= 5e-2 * scale    # north-star bf16 gate: 5e-2 rel (to the tensor scale)
    # the two batch-global predicates of visual_encoder.py:78,100, evaluated per camera call
    want = [1, 0] if "dark" in tag else [0, 1]
    assert eng.dino_prog.flags[:, :2].tolist() == [want, want]


@pytest.mark.parametrize("precise,tol", [(True, 2e-5), (False, 5e-2)])
def test_lstm_plan_reproduces_the_reference_golden(precise, tol):
    from vla_touch_b200.lstm_step_controller import LstmEngine
    A, Fd, T = 10, 3, 16
    g = U.golden(f"lstm_A{A}_F{Fd}_T{T}")
    mods = {"force_encoder": syn.synth_state_dict(shp.mlp_shapes([Fd, 128, 128]), 41, "lstm.force_encoder."),
            "lstm": syn.synth_state_dict(shp.lstm_shapes(128 + A), 41, "lstm.lstm."),
            "output_head": syn.synth_state_dict(shp.lstm_head_shapes(256, A), 41, "lstm.output_head.")}
    st = syn.synth_stats_varied(A, 41)
    vla_n = orc.normalize_actions(syn.det_uniform("lstm.vla", (3, T, A), 41, -1.0, 1.0), st, "vla")
    for denorm, key in ((False, "fwd"), (True, "seq")):
        eng = LstmEngine(mods, A, Fd, 256, 2, 3, T, "cpu", precise, denorm)
        eng.vla.copy_(vla_n); eng.forces.copy_(syn.det_normal("lstm.forces", (3, T, Fd), 41))
        eng.cond.copy_(syn.det_normal("lstm.cond", (3, 256), 41))
        eng.stats["action_mins"].copy_(st["action_mins"]); eng.stats["action_maxs"].copy_(st["action_maxs"])
        plan_emu.run(eng.plan)
        assert (eng.out - g[key]).abs().max() <= tol * max(1.0, float(g[key].abs().max()))


def test_ema_matches_torch_ema_semantics():
    from vla_touch_b200.ema import ExponentialMovingAverage
    p = [torch.nn.Parameter(torch.ones(3)), torch.nn.Parameter(torch.zeros(2))]
    ema = ExponentialMovingAverage(p, decay=0.75)
    with torch.no_grad():
        p[0].add_(1.0)
    ema.update()                                   # decay = min(0.75, 2/11)
    d = min(0.75, 2 / 11)
    assert torch.allclose(ema.shadow_params[0], torch.full((3,), 1.0 - (1 - d) * (1.0 - 2.0)))
    with ema.average_parameters():
        assert torch.equal(p[0].data, ema.shadow_params[0])
    assert torch.equal(p[0].data, torch.full((3,), 2.0))
    sd = ema.state_dict()
    assert set(sd) == {"decay", "num_updates", "shadow_params", "collected_params"} and sd["num_updates"] == 1
    ema2 = ExponentialMovingAverage(p, decay=0.5)
    ema2.load_state_dict(sd)
    assert ema2.decay == 0.75 and torch.equal(ema2.shadow_params[0], ema.shadow_params[0])


def test_parameter_trees_follow_the_reference_key_contract():
    from vla_touch_b200.bridge.networks.conditional_unet_1D_si import InterpolantsConditionalUnet1D
    net = InterpolantsConditionalUnet1D(10, 256)
    keys = [k for k, _ in net.named_parameters()]
    assert keys == list(shp.si_net_shapes(10, 256).keys())
    assert len(keys) == 438
    ref_dir = "/root/reference/VLA/residual_controller"
    if not os.path.isdir(ref_dir):
        pytest.skip("reference tree not present (GPU box)")
    sys.path.insert(0, ROOT)
    from oracle.ref_shims import import_reference
    ref = import_reference()
    rnet = ref.bridge_model.InterpolantsConditionalUnet1D(input_dim=10, global_cond_dim=256)
    assert [k for k, _ in rnet.named_parameters()] == keys
    assert [tuple(p.shape) for p in rnet.parameters()] == [tuple(p.shape) for p in net.parameters()]


@pytest.mark.parametrize("A,T", [(10, 16), (7, 64)])
def test_loss_plan_reproduces_the_reference_golden(A, T):
    from vla_touch_b200.unet import LossProgram
    g = U.golden(f"loss_A{A}_T{T}")
    full = U.net_sd(A, 21)
    sub = lambda n: {k[len(n):]: v for k, v in full.items() if k.startswith(n)}
    lp = LossProgram([sub("b_net."), sub("v_net."), sub("s_net.")], A, 3, T, 0.03, "cpu", precise=True)
    lp.x0.copy_(syn.det_uniform("loss.vla", (3, T, A), 24, -1.0, 1.0)); lp.x1.copy_(syn.det_uniform("loss.exp", (3, T, A), 24, -1.0, 1.0))
    lp.cond.copy_(syn.det_normal("loss.cond", (3, 256), 24)); lp.step.copy_(g["step"]); lp.z.copy_(g["z_unit"])
    plan_emu.run(lp.plan)
    for i, key in enumerate(("loss", "v_loss", "s_loss", "b_loss")):
        assert abs(float(lp.out[i]) - float(g[key])) <= 1e-3 * max(1.0, abs(float(g[key]))), key


def test_fused_mlp_descriptor_equals_the_two_gemm_descriptors():
    """vt_mlp_desc (one fused kernel on the GPU) means exactly what the fc1+GELU and fc2+LayerScale+residual GEMM descriptors of
    the unfused plan mean (HF Dinov2MLP, HF:312-328,380-386): both interpreted on the CPU from the same buffers."""
    from vla_touch_b200.plan import Plan, linear_desc, pack_linear_weight, ptr
    D, rows = 384, 150
    g = torch.Generator().manual_seed(5)
    w1, w2 = torch.randn(4 * D, D, generator=g) / D ** 0.5, torch.randn(D, 4 * D, generator=g) / (4 * D) ** 0.5
    outs = []
    for fused in (True, False):
        plan = Plan(torch.device("cpu"))
        xn = plan.buf("xn", (rows, D), torch.bfloat16)
        hid = plan.buf("hid", (rows, 4 * D), torch.bfloat16)
        h = plan.buf("h", (rows, D), torch.float32)
        gg = torch.Generator().manual_seed(6)
        xn.copy_(torch.randn(rows, D, generator=gg))
        h.copy_(torch.randn(rows, D, generator=gg))
        w1p, n1, k1 = pack_linear_weight(w1, torch.bfloat16)
        w2p, n2, k2 = pack_linear_weight(w2, torch.bfloat16)
        t = {k: plan.reg(v) for k, v in dict(w1=w1p, w2=w2p, b1=torch.linspace(-1, 1, 4 * D), b2=torch.linspace(1, -1, D),
                                             ls2=torch.linspace(0.5, 1.5, D)).items()}
        if fused:
            d = nv.MlpDesc()
            d.xn, d.ld_x, d.w1, d.w1_ld, d.b1 = ptr(xn), D, ptr(t["w1"]), k1, ptr(t["b1"])
            d.w2, d.w2_ld, d.b2, d.ls2 = ptr(t["w2"]), k2, ptr(t["b2"]), ptr(t["ls2"])
            d.h, d.ld_h, d.rows, d.D = ptr(h), D, rows, D
            plan.add(d, "mlp")
        else:
            plan.add(linear_desc(a=xn, rows=rows, k=k1, a_ld=D, w=t["w1"], n=4 * D, n_pad=n1, w_ld=k1, out=hid, ldc=4 * D,
                                 bias=t["b1"], act=nv.ACT_GELU), "fc1")
            plan.add(linear_desc(a=hid, rows=rows, k=k2, a_ld=4 * D, w=t["w2"], n=D, n_pad=n2, w_ld=k2, out=h, ldc=D, bias=t["b2"],
                                 colscale=t["ls2"], res=h, ldres=D), "fc2")
        plan_emu.run(plan)
        outs.append(h.clone())
    assert torch.equal(outs[0], outs[1])


def test_rowproj_descriptor_equals_the_gemm_and_layernorm_descriptors():
    """vt_rowproj_desc (one whole-row kernel on the GPU) means exactly what the attention output projection GEMM descriptor
    (bias, LayerScale, residual) followed by the norm2 LayerNorm descriptor mean (HF:238-251,374-381)."""
    from vla_touch_b200.plan import Plan, linear_desc, pack_linear_weight, ptr
    D, rows = 384, 150
    g = torch.Generator().manual_seed(7)
    w = torch.randn(D, D, generator=g) / D ** 0.5
    outs = []
    for fused in (True, False):
        plan = Plan(torch.device("cpu"))
        ctx = plan.buf("ctx", (rows, D), torch.bfloat16)
        xn = plan.buf("xn", (rows, D), torch.bfloat16)
        h = plan.buf("h", (rows, D), torch.float32)
        gg = torch.Generator().manual_seed(8)
        ctx.copy_(torch.randn(rows, D, generator=gg))
        h.copy_(torch.randn(rows, D, generator=gg) * 3)
        wp, n1, k1 = pack_linear_weight(w, torch.bfloat16)
        t = {k: plan.reg(v) for k, v in dict(w=wp, b=torch.linspace(-1, 1, D), ls=torch.linspace(0.5, 1.5, D),
                                             lg=torch.linspace(0.8, 1.2, D), lb=torch.linspace(-0.1, 0.1, D)).items()}
        if fused:
            d = nv.RowprojDesc()
            d.x, d.ld_x, d.w, d.w_ld, d.bias, d.colscale = ptr(ctx), D, ptr(t["w"]), k1, ptr(t["b"]), ptr(t["ls"])
            d.h, d.ld_h, d.rows, d.D = ptr(h), D, rows, D
            d.ln_gamma, d.ln_beta, d.ln_out, d.ln_ld, d.ln_eps = ptr(t["lg"]), ptr(t["lb"]), ptr(xn), D, 1e-6
            plan.add(d, "rowproj")
        else:
            plan.add(linear_desc(a=ctx, rows=rows, k=k1, a_ld=D, w=t["w"], n=D, n_pad=n1, w_ld=k1, out=h, ldc=D, bias=t["b"],
                                 colscale=t["ls"], res=h, ldres=D), "attn_out")
            d = nv.LnDesc()
            d.x, d.in_ld, d.in_row_stride, d.rows, d.D = ptr(h), D, 1, rows, D
            d.gamma, d.beta, d.eps = ptr(t["lg"]), ptr(t["lb"]), 1e-6
            d.out, d.out_dtype, d.out_ld, d.out_plane, d.act = ptr(xn), nv.VT_BF16, D, 0, nv.ACT_NONE
            plan.add(d, "norm2")
        plan_emu.run(plan)
        outs.append((h.clone(), xn.clone()))
    assert torch.equal(outs[0][0], outs[1][0]) and torch.equal(outs[0][1], outs[1][1])


def test_dgrad_descriptors_reproduce_the_explicit_backward():
    """unet_bwd.py: the data gradients of Conv1d(k5) / Conv1d(k3, stride 2) / ConvTranspose1d(k4, stride 2) expressed as forward
    implicit-GEMM descriptors (shifted taps, transposed weight slices, phase-interleaved rows), interpreted on the CPU, against
    oracle/vt_oracle_bwd.py (which is pinned to autograd and to the reference's gradients)."""
    from oracle import vt_oracle_bwd as ob
    from vla_touch_b200 import unet_bwd as ub
    from vla_touch_b200.plan import Plan
    from vla_touch_b200.unet import _View
    G, B, C = 2, 3, 256
    g = torch.Generator().manual_seed(9)
    bf = lambda t: t.to(torch.bfloat16).float()

    def run(kind, T_dy, T_dx, ws):
        plan = Plan(torch.device("cpu"))
        dy = plan.buf("dy", (G, B, T_dy, C), torch.bfloat16)
        dx = plan.buf("dx", (G, B, T_dx, C), torch.bfloat16)
        dy.copy_(torch.randn(G, B, T_dy, C, generator=g))
        ctx = ub.DgradCtx(G, precise=False)
        vy, vx = _View(dy, T_dy, C), _View(dx, T_dx, C)
        if kind == "k5":
            ub.conv_dgrad(plan, ctx, B, vy, vx, ws, pad=2)
        elif kind == "down":
            ub.downsample_dgrad(plan, ctx, B, vy, vx, ws)
        else:
            ub.upsample_dgrad(plan, ctx, B, vy, vx, ws)
        plan_emu.run(plan)
        for n in range(G):
            dyn = dy[n].float().permute(0, 2, 1)                      # [B, C, T]
            w = bf(ws[n])
            if kind == "k5":
                ref = ob.conv1d_bwd(torch.zeros(B, C, T_dx), w, dyn, padding=2)[0]
            elif kind == "down":
                ref = ob.conv1d_bwd(torch.zeros(B, C, T_dx), w, dyn, stride=2, padding=1)[0]
            else:
                ref = ob.convT1d_bwd(torch.zeros(B, C, T_dx), w, dyn)[0]
            got = dx[n].float().permute(0, 2, 1)
            err = (got - ref).abs().max().item()
            assert err <= 1e-2 * ref.abs().max().item(), (kind, n, err, ref.abs().max().item())

    run("k5", 16, 16, [torch.randn(C, C, 5, generator=g) / (5 * C) ** 0.5 for _ in range(G)])
    run("down", 8, 16, [torch.randn(C, C, 3, generator=g) / (3 * C) ** 0.5 for _ in range(G)])
    run("up", 32, 16, [torch.randn(C, C, 4, generator=g) / (2 * C) ** 0.5 for _ in range(G)])


@pytest.mark.parametrize("kind", ["k5", "down", "up", "k1in"])
def test_wgrad_descriptors_reproduce_the_explicit_backward(kind):
    """unet_bwd.conv_wgrad: weight gradients as tcol (transposed im2col) + one plain GEMM descriptor, interpreted on the CPU,
    against conv1d_bwd / convT1d_bwd of oracle/vt_oracle_bwd.py."""
    import bwd_cases
    plan, check = bwd_cases.wgrad_case(kind, torch.device("cpu"))
    plan_emu.run(plan)
    check()


@pytest.mark.parametrize("film", [False, True])
def test_conv_block_backward_plan_reproduces_the_explicit_backward(film):
    """unet_bwd.conv_block_backward (recomputed raw conv -> GroupNorm+Mish(+FiLM) backward -> wgrad -> dgrad) interpreted on the
    CPU against gn_mish_bwd + conv1d_bwd (+ the FiLM reductions of _res_block_bwd) of the oracle."""
    import bwd_cases
    plan, check = bwd_cases.block_case(torch.device("cpu"), film)
    plan_emu.run(plan)
    check()
    plan, check = bwd_cases.colsum_case(torch.device("cpu"))
    plan_emu.run(plan)
    check()


@pytest.mark.parametrize("ci,co", [(256, 256), (256, 512), (1024, 512), (7, 256)])
def test_res_block_backward_plan_reproduces_the_explicit_backward(ci, co):
    """unet_bwd.res_block_backward (ConditionalResidualBlock1D: both Conv1dBlocks, FiLM, identity / 1x1-conv residual, the
    first block's shared padded input) interpreted on the CPU against _res_block_bwd of oracle/vt_oracle_bwd.py."""
    import bwd_cases
    plan, check = bwd_cases.res_block_case(torch.device("cpu"), ci, co)
    plan_emu.run(plan)
    check()


def test_unet_backward_plan_reproduces_the_explicit_backward():
    """unet_train: training forward + build_unet_backward + build_film_time_backward of whole U-Nets (146 parameter gradients
    per net + d global_cond), interpreted on the CPU, against unet_forward_cached / unet_backward of oracle/vt_oracle_bwd.py."""
    import bwd_cases
    plan, check = bwd_cases.unet_case(torch.device("cpu"))
    plan_emu.run(plan)
    res = check()
    assert res["missing"] == [] and res["n"] == 148


@pytest.mark.parametrize("A,T", [(10, 16), (7, 64)])
def test_loss_backward_program_reproduces_the_reference_gradients(A, T):
    """LossBackwardProgram = get_loss(...).backward() of the reference (bridge_model.py:220-246) as one plan, interpreted on the
    CPU: the reference's own gradient digests (tests/golden/loss_grads_*.npz) and every gradient tensor in full."""
    import bwd_cases
    plan, check = bwd_cases.loss_case(torch.device("cpu"), A, T)
    plan_emu.run(plan)
    assert check()["tensors"] == 439


def test_get_loss_backward_through_autograd(monkeypatch):
    """StochasticInterpolants.get_loss -> loss.backward() (the reference training step, bridge_train.py:315-334) with the
    native program replaced by the CPU descriptor interpreter: .grad of all 438 parameters and the gradient reaching the
    producer of obs_cond, before and after an in-place parameter update (re-packed operand copies)."""
    import bwd_cases
    from vla_touch_b200.plan import Plan

    class _Interp:
        def __init__(self, plan):
            self.plan = plan

        def run(self, first=0, count=-1):
            plan_emu.run(self.plan, first, count)

    monkeypatch.setattr(Plan, "compile", lambda self: _Interp(self))
    res = bwd_cases.training_step_case(torch.device("cpu"))()
    assert res["first"] <= 3e-2 and res["after_update"] <= 3e-2
    # the gradients live in the program's buffers: a backward() that comes after a later get_loss() must fail loudly
    si = bwd_cases.interpolant(10, 16, torch.device("cpu"))
    batch = {"obs_cond": syn.det_normal("loss.cond", (3, 256), 24), "expert_act": syn.det_uniform("loss.exp", (3, 16, 10), 24, -1.0, 1.0),
             "vla_act": syn.det_uniform("loss.vla", (3, 16, 10), 24, -1.0, 1.0)}
    first, _ = si.get_loss(batch, "cpu")
    second, _ = si.get_loss(batch, "cpu")
    with pytest.raises(RuntimeError, match="called again before"):
        first.backward()
    second.backward()


def test_training_loop_reduces_the_loss(monkeypatch):
    """The reference's training step verbatim (bridge_train.py:305-337: zero_grad, get_loss, backward, AdamW.step, ema.update)
    on a fixed batch, with the native program replaced by the CPU descriptor interpreter: the loss goes down, the EMA shadow
    follows, and the program picks up the updated weights (re-packed operand copies) at every step."""
    import bwd_cases
    from vla_touch_b200.plan import Plan

    class _Interp:
        def __init__(self, plan):
            self.plan = plan

        def run(self, first=0, count=-1):
            plan_emu.run(self.plan, first, count)

    monkeypatch.setattr(Plan, "compile", lambda self: _Interp(self))
    A, T = 10, 16
    si = bwd_cases.interpolant(A, T, torch.device("cpu"))
    g = U.golden(f"loss_A{A}_T{T}")
    si.step_override, si.z_override = torch.as_tensor(g["step"]), torch.as_tensor(g["z_unit"])
    batch = {"obs_cond": syn.det_normal("loss.cond", (3, 256), 24), "expert_act": syn.det_uniform("loss.exp", (3, T, A), 24, -1.0, 1.0),
             "vla_act": syn.det_uniform("loss.vla", (3, T, A), 24, -1.0, 1.0)}
    opt = torch.optim.AdamW(si.net.parameters(), lr=1e-4, weight_decay=1e-6)
    shadow0 = [s.clone() for s in si.ema.shadow_params]
    losses = []
    for _ in range(3):
        opt.zero_grad()
        loss, info = si.get_loss(batch, "cpu")
        loss.backward()
        opt.step()
        si.ema.update()
        losses.append(float(loss.detach()))
    assert losses[2] < losses[1] < losses[0], losses
    assert any(not torch.equal(a, b) for a, b in zip(shadow0, si.ema.shadow_params))
    with torch.no_grad():                                    # validation path on the updated weights (fp32 forward program)
        val, _ = si.get_loss(batch, "cpu")
    assert abs(float(val) - losses[2]) < abs(losses[0] - losses[2]) and float(val) < losses[1]


@pytest.mark.parametrize("kind", ["k5", "down", "up"])
def test_split_k_wgrad_descriptors(kind, monkeypatch):
    """conv_wgrad(split_k=2): batch slices as extra GEMM groups + one column-sum launch give the same weight gradient; and the
    whole get_loss backward is unchanged with VT_WGRAD_SPLITK=3 (B = 3: one sample per slice; the shared-input convs stay unsplit)."""
    import bwd_cases
    plan, check = bwd_cases.wgrad_case(kind, torch.device("cpu"), B=4, split_k=2)
    assert sum(isinstance(d, nv.ColsumDesc) for d in plan.descs) == 1
    plan_emu.run(plan)
    check()
    if kind == "k5":
        monkeypatch.setenv("VT_WGRAD_SPLITK", "3")
        plan, check = bwd_cases.loss_case(torch.device("cpu"), 10, 16)
        assert sum(isinstance(d, nv.ColsumDesc) and "splitk" in t for d, t in zip(plan.descs, plan.tags)) >= 30
        plan_emu.run(plan)
        assert check()["tensors"] == 439


def test_backward_is_additive_over_the_batch(monkeypatch):
    """The size-independent property the full-size GPU test uses (tests/bwd_cases.batch_additivity_case), here at batch 4 on the
    CPU descriptor interpreter."""
    import bwd_cases
    from vla_touch_b200.plan import Plan

    class _Interp:
        def __init__(self, plan):
            self.plan = plan

        def run(self, first=0, count=-1):
            plan_emu.run(self.plan, first, count)

    monkeypatch.setattr(Plan, "compile", lambda self: _Interp(self))
    res = bwd_cases.batch_additivity_case(torch.device("cpu"), B=4, T=16, A=10)()
    assert res["tensors"] == 439


def test_lstm_layers_training_plans_match_nn_lstm_autograd():
    """lstm_train.lstm_layers_train: training forward (gates and cell states kept) + BPTT of the two stacked LSTM layers as plan
    ops (recurrence kernels + weight-gradient GEMMs + column sums), interpreted on the CPU, against torch.nn.LSTM autograd."""
    import bwd_cases
    plan, check = bwd_cases.lstm_layers_case(torch.device("cpu"))
    plan_emu.run(plan)
    errs = check()
    assert len(errs) == 10


@pytest.mark.parametrize("A,Fd,T", [(10, 3, 16), (7, 64, 32)])
def test_lstm_loss_backward_program_reproduces_the_reference_gradients(A, Fd, T):
    """lstm_train.LstmLossBackwardProgram = TactileLSTMController.get_loss(...).backward() (lstm_step_controller.py:321-337,
    eval-equivalent: dropout off) as one plan, interpreted on the CPU: loss, d obs_cond and all 18 parameter gradients against the
    reference's digests (tests/golden/lstm_grads_*.npz) and the explicit BPTT oracle."""
    import bwd_cases
    plan, check = bwd_cases.lstm_loss_case(torch.device("cpu"), A, Fd, T)
    plan_emu.run(plan)
    assert check()["tensors"] == 18


def test_lstm_get_loss_backward_through_autograd(monkeypatch):
    """TactileLSTMController.get_loss(batch, differentiable=True) -> loss.backward() (lstm_train.py:120-130) with the native
    program replaced by the CPU descriptor interpreter: .grad of the 18 parameters and of obs_cond against the explicit BPTT
    oracle, before and after an in-place parameter update (operands re-packed into the same tensors)."""
    import bwd_cases
    import torch.nn as nn
    from oracle import vt_oracle_bwd as ob
    from vla_touch_b200.lstm_step_controller import TactileLSTMController
    from vla_touch_b200.plan import Plan

    class _Interp:
        def __init__(self, plan):
            self.plan = plan

        def run(self, first=0, count=-1):
            plan_emu.run(self.plan, first, count)

    monkeypatch.setattr(Plan, "compile", lambda self: _Interp(self))
    A, Fd, T, H = 7, 64, 32, 256
    lc = object.__new__(TactileLSTMController)            # the parameter containers only: no DinoV2 encoder on a CPU-only box
    lc.device, lc.state_dim, lc.hidden_dim, lc.force_dim = "cpu", A, H, Fd
    lc.force_encoder = nn.Sequential(nn.Linear(Fd, H // 2), nn.GELU(), nn.Linear(H // 2, H // 2))
    lc.lstm = nn.LSTM(input_size=H // 2 + A, hidden_size=H, num_layers=2, batch_first=True, dropout=0.1)
    lc.output_head = nn.Sequential(nn.Linear(2 * H, H), nn.LayerNorm(H), nn.GELU(), nn.Dropout(0.1), nn.Linear(H, A))
    lc.trainable_modules = [lc.force_encoder, lc.lstm, lc.output_head]
    lc.eval()                                             # deterministic network (train mode adds Philox dropout masks)
    for nm, mod in (("force_encoder", lc.force_encoder), ("lstm", lc.lstm), ("output_head", lc.output_head)):
        syn.fill_named_(mod.named_parameters(), 41, prefix=f"lstm.{nm}.")
    inp = bwd_cases.lstm_fixture_inputs(A, Fd, T)

    def one_step():
        for m in lc.trainable_modules:
            m.zero_grad()
        cond = inp["cond"].clone().requires_grad_(True)
        loss = lc.get_loss({"vla_act": inp["vla_n"], "obs_cond": cond, "forces": inp["forces"], "expert_act": inp["expert"]},
                           differentiable=True)
        (3.0 * loss).backward()
        mods = {"force_encoder": lc.force_encoder.state_dict(), "lstm": lc.lstm.state_dict(), "output_head": lc.output_head.state_dict()}
        mods = {m: {k: v.detach().clone() for k, v in sd.items()} for m, sd in mods.items()}
        ref_loss, ref, dcond = ob.lstm_loss_backward(mods, inp["vla_n"], inp["cond"], inp["forces"], inp["expert"])
        assert abs(float(loss.detach()) - float(ref_loss)) <= 2e-2 * abs(float(ref_loss))
        worst = float((cond.grad - 3.0 * dcond).abs().max() / (3.0 * dcond).abs().max())
        for mname, mod in (("force_encoder", lc.force_encoder), ("lstm", lc.lstm), ("output_head", lc.output_head)):
            for n, p_ in mod.named_parameters():
                r = 3.0 * ref[f"{mname}.{n}"]
                worst = max(worst, float((p_.grad - r).abs().max() / r.abs().max()))
        return worst

    assert one_step() <= 3e-2
    with torch.no_grad():
        for m in lc.trainable_modules:
            for p_ in m.parameters():
                p_.mul_(1.03)
    assert one_step() <= 3e-2


@pytest.mark.parametrize("A,T,B", [(10, 4, 1), (7, 48, 2)])
def test_loss_backward_program_edge_shapes(A, T, B):
    """Smallest horizon (T = 4: one position at the deepest level) with a single row, and the reference's T = 48 variant: every
    gradient tensor in full against autograd through the forward oracle (bf16 gate 4e-2 of the tensor's scale)."""
    from vla_touch_b200.unet_train import LossBackwardProgram
    full = U.net_sd(A, 21)
    sub = lambda n: {k[len(n):]: v for k, v in full.items() if k.startswith(n)}
    g = torch.Generator().manual_seed(3)
    x0, x1 = torch.rand(B, T, A, generator=g) * 2 - 1, torch.rand(B, T, A, generator=g) * 2 - 1
    cond, step, z = torch.randn(B, 256, generator=g), torch.rand(B, generator=g), torch.randn(B, T, A, generator=g)
    lp = LossBackwardProgram([sub("b_net."), sub("v_net."), sub("s_net.")], A, B, T, 0.03, "cpu")
    lp.set_inputs(x0, x1, cond, step, z)
    plan_emu.run(lp.plan)
    sd = {k: v.clone().requires_grad_(True) for k, v in full.items()}
    c = cond.clone().requires_grad_(True)
    loss, *_ = orc.bridge_losses(sd, c, x1, x0, step, z)
    loss.backward()
    assert abs(float(lp.out[0]) - float(loss.detach())) <= 2e-2 * abs(float(loss.detach()))
    grads = lp.grads
    for n, p in sd.items():
        assert float((grads[n].float() - p.grad).abs().max()) <= 4e-2 * float(p.grad.abs().max()), n
    assert float((lp.d_cond - c.grad).abs().max()) <= 4e-2 * float(c.grad.abs().max())
    with pytest.raises(ValueError):
        LossBackwardProgram([sub("b_net."), sub("v_net."), sub("s_net.")], A, B, 6, 0.03, "cpu")      # T must be a multiple of 4


def test_lstm_training_program_with_dropout():
    """LstmLossBackwardProgram(dropout=0.1): inverted-dropout masks between the LSTM layers and in the head (training-mode
    semantics of the reference, lstm_step_controller.py:66-82), injected uniforms, against torch autograd with the same masks."""
    import bwd_cases
    plan, check = bwd_cases.lstm_dropout_case(torch.device("cpu"))
    plan_emu.run(plan)
    assert check()["tensors"] == 18


def test_native_encoder_training_matches_torch_autograd(monkeypatch):
    """mlp_train: the 3-layer GELU state encoder (bridge_controller.py:42-48) forward + backward as plan ops behind a
    torch.autograd.Function, interpreted on the CPU, against torch autograd of the same nn.Sequential -- before and after an
    in-place parameter update."""
    import torch.nn as nn
    from vla_touch_b200 import mlp_train as mt
    from vla_touch_b200.plan import Plan

    class _Interp:
        def __init__(self, plan):
            self.plan = plan

        def run(self, first=0, count=-1):
            plan_emu.run(self.plan, first, count)

    monkeypatch.setattr(Plan, "compile", lambda self: _Interp(self))
    g = torch.Generator().manual_seed(5)
    enc = nn.Sequential(nn.Linear(839, 256), nn.GELU(), nn.Linear(256, 256), nn.GELU(), nn.Linear(256, 256))
    x = torch.randn(6, 839, generator=g)
    dout = torch.randn(6, 256, generator=g)
    cache = {}
    for step in range(2):
        enc.zero_grad()
        out = mt.encoder_forward(enc, cache, x)
        (out * dout).sum().backward()
        got = {n: p.grad.clone() for n, p in enc.named_parameters()}
        enc.zero_grad()
        ref_out = enc(x)
        (ref_out * dout).sum().backward()
        assert float((out - ref_out).detach().abs().max()) <= 2e-2 * float(ref_out.detach().abs().max())
        for n, p in enc.named_parameters():
            assert float((got[n] - p.grad).abs().max()) <= 3e-2 * float(p.grad.abs().max()), (step, n)
        with torch.no_grad():
            for p in enc.parameters():
                p.mul_(1.05)
    assert len(cache) == 1


def test_checkpoint_files_have_the_reference_structure(tmp_path):
    """controller.pt / bridge_model.pt / tactile_controller.pt written by vla_touch_b200 have exactly the key paths, order, shapes
    and dtypes of the files the reference writes (bridge_controller.py:203-244, bridge_model.py:435-447,
    lstm_step_controller.py:351-379; manifest recorded from the reference by oracle/gen_golden_extra.py), incl. the `ema` dict
    (decay, num_updates, shadow_params in net.parameters() order, collected_params)."""
    import json
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    from oracle.gen_golden_extra import manifest
    from vla_touch_b200.bridge_controller import DiffusionController
    from vla_touch_b200.lstm_step_controller import TactileLSTMController
    gold = json.load(open(os.path.join(ROOT, "tests", "golden", "ckpt_manifest.json")))
    A, Fd, T = gold["A"], gold["F"], gold["T"]
    ma = {'interpolant_type': 'linear', 'gamma_type': '2^0.5*t(t-1)', 'epsilon_type': '1-t', 'prior_policy': 'vla', 'beta_max': 0.03,
          'sde_type': 'bs', 'action_dim': A, 'obs_dim': 256, 'obs_horizon': 1, 'net_type': 'unet1D_si', 'pretrain': False,
          'context_frames': 2, 'horizon': T}
    dino = U.dino_sd(384, 1, 1)
    ctl = DiffusionController(state_dim=A, hidden_dim=256, diffusion_steps=10, device="cpu", model_args=ma, use_force=True, force_dim=Fd,
                              image_state_dict=dino)
    ctl.stats = {k: v.numpy() for k, v in syn.synth_stats(A).items()}
    ctl.save(str(tmp_path))
    lc = TactileLSTMController(state_dim=A, hidden_dim=256, num_layers=2, dropout=0.1, device="cpu", force_dim=Fd, image_state_dict=dino)
    lc.stats = {k: torch.as_tensor(v) for k, v in ctl.stats.items()}
    lc.save(str(tmp_path))
    for f, want in gold["files"].items():
        got = manifest(torch.load(os.path.join(str(tmp_path), f), map_location="cpu", weights_only=False))
        strip = lambda m: [[a, b, c, d] for a, b, c, d in m]
        assert len(got) == len(want), (f, len(got), len(want))
        for g_, w_ in zip(strip(got), want):
            assert g_ == w_, (f, g_, w_)
    # and they load back (round trip through the files)
    ctl2 = DiffusionController(state_dim=A, hidden_dim=256, diffusion_steps=10, device="cpu", model_args=ma, use_force=True, force_dim=Fd,
                               image_state_dict=dino)
    ctl2.load(str(tmp_path))
    for a, b in zip(ctl.diffusion_model.net.parameters(), ctl2.diffusion_model.net.parameters()):
        assert torch.equal(a, b)
    assert ctl2.diffusion_model.ema.num_updates == ctl.diffusion_model.ema.num_updates
    lc2 = TactileLSTMController(state_dim=A, hidden_dim=256, num_layers=2, dropout=0.1, device="cpu", force_dim=Fd, image_state_dict=dino)
    lc2.load(str(tmp_path))
    for a, b in zip(lc.lstm.parameters(), lc2.lstm.parameters()):
        assert torch.equal(a, b)


@pytest.mark.skipif(not os.path.isdir("/root/reference"), reason="needs the reference tree (build container only)")
def test_checkpoints_interchange_with_the_reference_classes(tmp_path):
    """Files SAVED BY THE IMPORTED REFERENCE load into the drop-in classes with identical tensors, and files saved by the drop-in
    classes load into the reference classes (bridge_controller.py:203-244, bridge_model.py:421-447, lstm_step_controller.py:351-379)."""
    from oracle.ref_shims import import_reference
    from vla_touch_b200.bridge_controller import DiffusionController
    from vla_touch_b200.lstm_step_controller import TactileLSTMController
    import contextlib, io
    ref = import_reference(num_dino_layers=1)
    A, Fd, T = 7, 64, 16
    ma = {'interpolant_type': 'linear', 'gamma_type': '2^0.5*t(t-1)', 'epsilon_type': '1-t', 'prior_policy': 'vla', 'beta_max': 0.03,
          'sde_type': 'vs', 'action_dim': A, 'obs_dim': 256, 'obs_horizon': 1, 'net_type': 'unet1D_si', 'pretrain': False,
          'context_frames': 2, 'horizon': T}
    with contextlib.redirect_stdout(io.StringIO()):
        rc = ref.bridge_controller.DiffusionController(state_dim=A, hidden_dim=256, image_model_path="facebook/dinov2-small",
                                                       diffusion_steps=10, device="cpu", model_args=dict(ma), use_force=True, force_dim=Fd)
        rl = ref.lstm_step_controller.TactileLSTMController(state_dim=A, hidden_dim=256, num_layers=2, dropout=0.1, device="cpu", force_dim=Fd)
    syn.fill_named_(rc.diffusion_model.net.named_parameters(), 5, prefix="net.")
    syn.fill_named_(rc.state_encoder.named_parameters(), 5, prefix="enc.")
    rc.diffusion_model.ema.update()                       # shadow != live, num_updates = 1
    rc.stats = {k: v.numpy() for k, v in syn.synth_stats_varied(A, 2).items()}
    rl.stats = {k: torch.as_tensor(v) for k, v in rc.stats.items()}
    d_ref = tmp_path / "ref"; d_ref.mkdir()
    rc.save(str(d_ref)); rl.save(str(d_ref))
    dino = U.dino_sd(384, 1, 1)
    ours = DiffusionController(state_dim=A, hidden_dim=256, diffusion_steps=10, device="cpu", model_args=dict(ma), use_force=True,
                               force_dim=Fd, image_state_dict=dino)
    _cuda = torch.Tensor.cuda
    ours.load(str(d_ref))
    for (n, a), b in zip(rc.diffusion_model.net.named_parameters(), ours.diffusion_model.net.parameters()):
        assert torch.equal(a, b), n
    for a, b in zip(rc.diffusion_model.ema.shadow_params, ours.diffusion_model.ema.shadow_params):
        assert torch.equal(a, b)
    assert ours.diffusion_model.ema.num_updates == 1 and ours.diffusion_model.ema.decay == rc.diffusion_model.ema.decay
    for a, b in zip(rc.state_encoder.parameters(), ours.state_encoder.parameters()):
        assert torch.equal(a, b)
    for k, v in rc.stats.items():
        assert torch.equal(ours.stats[k].cpu(), torch.as_tensor(v, dtype=torch.float32))
    lo = TactileLSTMController(state_dim=A, hidden_dim=256, num_layers=2, dropout=0.1, device="cpu", force_dim=Fd, image_state_dict=dino)
    lo.load(str(d_ref))
    for mod in ("obs_encoder", "force_encoder", "lstm", "output_head"):
        for (n, a), b in zip(getattr(rl, mod).named_parameters(), getattr(lo, mod).parameters()):
            assert torch.equal(a, b), (mod, n)
    # the other direction: our files into the reference's state dicts (its own load() hard-codes .cuda(), bridge_controller.py:241)
    d_ours = tmp_path / "ours"; d_ours.mkdir()
    with torch.no_grad():
        for p_ in ours.diffusion_model.net.parameters():
            p_.mul_(1.01)
    ours.save(str(d_ours)); lo.save(str(d_ours))
    ck = torch.load(str(d_ours / "controller.pt"), map_location="cpu", weights_only=False)
    rc.state_encoder.load_state_dict(ck['state_encoder']); rc.force_decoder.load_state_dict(ck['force_decoder'])
    rc.diffusion_model.load_model({**ck['model_args'], 'ckpt_path': str(d_ours), 'pretrain': True}, "cpu")
    for a, b in zip(rc.diffusion_model.net.parameters(), ours.diffusion_model.net.parameters()):
        assert torch.equal(a, b)
    ck = torch.load(str(d_ours / "tactile_controller.pt"), map_location="cpu", weights_only=False)
    for mod in ("obs_encoder", "force_encoder", "lstm", "output_head"):
        getattr(rl, mod).load_state_dict(ck['modules'][mod])
