"""Pins the oracle restatement (oracle/vt_oracle.py) to outputs of the UNMODIFIED reference
(tests/golden/*.npz, produced by oracle/gen_golden.py).  CPU only."""
import os
import sys

import pytest
import torch

from oracle import vt_oracle as orc
from vla_touch_b200 import shapes as shp
from vla_touch_b200 import synthetic as syn
import vt_testutil as U

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

TOL = dict(rtol=0, atol=3e-6)


@pytest.mark.parametrize("A", [7, 10])
def test_normalize_denormalize(A):
    g = U.golden(f"norm_A{A}")
    st = syn.synth_stats_varied(A, seed=3)
    x = syn.det_uniform("norm.x", (3, 16, A), 3, -2.0, 2.0)
    assert torch.equal(orc.normalize_actions(x, st, "vla"), g["vla_n"])
    assert torch.equal(orc.normalize_actions(x, st, "expert"), g["exp_n"])
    assert torch.equal(orc.denormalize_actions(x, st, "expert"), g["exp_dn"])
    with pytest.raises(ValueError):
        orc.normalize_actions(x, st, "nope")


@pytest.mark.parametrize("case", U.DINO_CASES, ids=[c[0] for c in U.DINO_CASES])
def test_dinov2(case):
    tag, hidden, heads, layers, hw, batch, kind, seed = case
    g = U.golden(tag)
    sd = {k[len("dino."):]: v for k, v in U.dino_sd(hidden, layers, seed).items()}
    sd = U.dino_sd(hidden, layers, seed)
    img = U.images_for(kind, "dino.img", batch, hw, seed)
    out = orc.dino_encoder_forward(sd, img, heads)
    torch.testing.assert_close(out, g["cls"], **TOL)
    if "h_emb" in g:
        pv = orc.dinov2_preprocess(img)
        h = orc.dinov2_embeddings(sd, pv[:1])
        idx = U.TOK_IDX(h.shape[1])
        torch.testing.assert_close(h[:, idx], g["h_emb"], **TOL)
        for i in range(layers):
            h = orc.dinov2_layer(sd, i, h, heads)
            torch.testing.assert_close(h[:, idx], g[f"h_l{i}"], rtol=0, atol=2e-5)


@pytest.mark.parametrize("A,Fd", [(10, 3), (7, 64)])
def test_state_encoder_and_unet(A, Fd):
    enc = U.enc_sd(2 * 384 + A + Fd, 21)
    xin = syn.det_normal("enc.in", (4, 2 * 384 + A + Fd), 21)
    torch.testing.assert_close(orc.mlp3_gelu(enc, xin), U.golden(f"enc_A{A}_F{Fd}")["out"], **TOL)
    v_sd, s_sd = U.net_sd(A, 21, "v_net"), U.net_sd(A, 21, "s_net")
    for T in (16, 32, 48, 64):
        g = U.golden(f"unet_A{A}_T{T}")
        x = syn.det_uniform("unet.x", (3, T, A), 22, -1.0, 1.0)
        cond = syn.det_normal("unet.cond", (3, 256), 22)
        t = torch.tensor([0.3, 0.001, 0.999])
        torch.testing.assert_close(orc.unet_forward(v_sd, x, t, cond), g["v"], rtol=0, atol=1e-5)
        torch.testing.assert_close(orc.unet_forward(s_sd, x, t[:1].expand(3), cond), g["s"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("A,T,n", [(10, 16, 10), (7, 64, 10), (7, 64, 50)])
def test_sde_vs_recorded_noise(A, T, n):
    g = U.golden(f"sde_A{A}_T{T}_n{n}")
    v_sd, s_sd = U.net_sd(A, 1021, "v_net"), U.net_sd(A, 1021, "s_net")   # EMA shadow (seed+1000)
    x0 = syn.det_uniform("sde.x0", (2, T, A), 23, -1.0, 1.0)
    cond = syn.det_normal("sde.cond", (2, 256), 23)
    out, traj = orc.sde_vs(v_sd, s_sd, x0, cond, n, 0.03, g["noise"], return_traj=True)
    torch.testing.assert_close(traj[1], g["x1"], rtol=0, atol=1e-5)
    torch.testing.assert_close(out, g["out"], rtol=0, atol=5e-5)
    # live (non-EMA) weights must NOT reproduce it: sample() runs under ema.average_parameters()
    bad = orc.sde_vs(U.net_sd(A, 21, "v_net"), U.net_sd(A, 21, "s_net"), x0, cond, n, 0.03, g["noise"])
    assert (bad - g["out"]).abs().max() > 1e-3


@pytest.mark.parametrize("A,T", [(10, 16), (7, 64)])
def test_sde_vs_beta0(A, T):
    g = U.golden(f"sde_A{A}_T{T}_n10_beta0")
    x0 = syn.det_uniform("sde.x0", (2, T, A), 23, -1.0, 1.0)
    cond = syn.det_normal("sde.cond", (2, 256), 23)
    out = orc.sde_vs(U.net_sd(A, 1021, "v_net"), U.net_sd(A, 1021, "s_net"), x0, cond, 10, 0.0,
                     torch.zeros(10, 2, T, A))
    torch.testing.assert_close(out, g["out"], rtol=0, atol=5e-5)


def test_sde_bs_recorded_noise():
    """sde_type='bs' (bridge_model.py:271-273,281-332): b_net drives the drift; fixture from the reference itself
    (oracle/gen_golden_extra.py), EMA weights."""
    A, T = 7, 64
    g = U.golden(f"sde_bs_A{A}_T{T}_n10")
    x0 = syn.det_uniform("sde.x0", (2, T, A), 23, -1.0, 1.0)
    cond = syn.det_normal("sde.cond", (2, 256), 23)
    out = orc.sde_bs(U.net_sd(A, 1021, "b_net"), U.net_sd(A, 1021, "s_net"), x0, cond, 10, 0.03, g["noise"])
    torch.testing.assert_close(out, g["out"], rtol=0, atol=5e-5)
    bad = orc.sde_vs(U.net_sd(A, 1021, "v_net"), U.net_sd(A, 1021, "s_net"), x0, cond, 10, 0.03, g["noise"])
    assert (bad - g["out"]).abs().max() > 1e-3


def test_sde_schedule_step_count_quirk():
    # n = int(1/float(1/diffuse_step)) differs from diffuse_step for some values
    for ds, n in ((10, 10), (50, 50), (93, 92), (99, 98)):
        assert orc.sde_schedule(ds)[0] == n


@pytest.mark.parametrize("A,T", [(10, 16), (7, 64)])
def test_losses(A, T):
    g = U.golden(f"loss_A{A}_T{T}")
    batch = {"obs_cond": syn.det_normal("loss.cond", (3, 256), 24),
             "expert_act": syn.det_uniform("loss.exp", (3, T, A), 24, -1.0, 1.0),
             "vla_act": syn.det_uniform("loss.vla", (3, T, A), 24, -1.0, 1.0)}
    loss, v, s, b = orc.bridge_losses(U.net_sd(A, 21), batch["obs_cond"], batch["expert_act"], batch["vla_act"],
                                      g["step"], g["z_unit"])
    for got, key in ((loss, "loss"), (v, "v_loss"), (s, "s_loss"), (b, "b_loss")):
        torch.testing.assert_close(got, g[key], rtol=1e-5, atol=1e-5)


@pytest.mark.parametrize("A,T", [(10, 16), (7, 64)])
def test_loss_gradients(A, T):
    """Autograd through the oracle's bridge_losses against the reference's loss.backward() (oracle/gen_golden_grads.py):
    d loss / d obs_cond in full, and norm / sum / first elements of every parameter gradient of net.{b,v,s}_net.  This is the
    parity gate the backward kernels will be held to."""
    g, gg = U.golden(f"loss_A{A}_T{T}"), U.golden(f"loss_grads_A{A}_T{T}")
    sd = {k: v.clone().requires_grad_(True) for k, v in U.net_sd(A, 21).items()}
    cond = syn.det_normal("loss.cond", (3, 256), 24).requires_grad_(True)
    loss, *_ = orc.bridge_losses(sd, cond, syn.det_uniform("loss.exp", (3, T, A), 24, -1.0, 1.0),
                                 syn.det_uniform("loss.vla", (3, T, A), 24, -1.0, 1.0), g["step"], g["z_unit"])
    loss.backward()
    torch.testing.assert_close(cond.grad, gg["d_cond"], rtol=1e-4, atol=1e-5 * float(gg["d_cond"].abs().max()))
    names = [str(n) for n in gg["names"]]
    assert sorted(names) == sorted(sd.keys())
    for i, n in enumerate(names):
        gr = sd[n].grad.flatten().double()
        scale = max(float(gg["norm"][i]), 1e-12)
        assert abs(float(gr.norm()) - float(gg["norm"][i])) <= 1e-4 * scale, n
        assert abs(float(gr.sum()) - float(gg["sum"][i])) <= 1e-3 * scale, n
        k = min(8, gr.numel())
        assert float((gr[:k] - gg["head"][i][:k].double()).abs().max()) <= 1e-4 * scale, n


@pytest.mark.parametrize("tag", list(U.PREDICT_CASES))
def test_predict(tag):
    c = U.predict_case(tag)
    cond = orc.encode_observation(c["dino"], c["enc"], c["heads"], c["state"], c["img1"], c["img2"], c["forces"])
    torch.testing.assert_close(cond, c["gold"]["cond"], rtol=0, atol=2e-5)
    out = orc.predict(c["dino"], c["enc"], c["v_ema"], c["s_ema"], c["stats"], c["heads"], c["state"], c["vla"],
                      c["img1"], c["img2"], c["forces"], c["steps"], 0.03, c["gold"]["noise"])
    torch.testing.assert_close(out, c["gold"]["out"], rtol=0, atol=1e-4)


@pytest.mark.parametrize("A,Fd,T", [(10, 3, 16), (7, 64, 32)])
def test_lstm(A, Fd, T):
    g = U.golden(f"lstm_A{A}_F{Fd}_T{T}")
    mods = {
        "force_encoder": syn.synth_state_dict(shp.mlp_shapes([Fd, 128, 128]), 41, "lstm.force_encoder."),
        "lstm": syn.synth_state_dict(shp.lstm_shapes(128 + A), 41, "lstm.lstm."),
        "output_head": syn.synth_state_dict(shp.lstm_head_shapes(256, A), 41, "lstm.output_head."),
    }
    st = syn.synth_stats_varied(A, 41)
    vla = syn.det_uniform("lstm.vla", (3, T, A), 41, -1.0, 1.0)
    forces = syn.det_normal("lstm.forces", (3, T, Fd), 41)
    cond = syn.det_normal("lstm.cond", (3, 256), 41)
    expert = syn.det_uniform("lstm.exp", (3, T, A), 41, -1.0, 1.0)
    vla_n = orc.normalize_actions(vla, st, "vla")
    fwd = orc.lstm_forward(mods, vla_n, cond, forces)
    torch.testing.assert_close(fwd, g["fwd"], rtol=0, atol=1e-5)
    torch.testing.assert_close(torch.nn.functional.mse_loss(fwd, expert), g["loss"], rtol=1e-5, atol=1e-6)
    # predict_sequence == step-by-step forward + expert de-normalisation (eval mode) :288-319
    torch.testing.assert_close(orc.denormalize_actions(fwd, st, "expert"), g["seq"], rtol=0, atol=1e-5)


@pytest.mark.parametrize("A,Fd,T", [(10, 3, 16), (7, 64, 32)])
def test_lstm_loss_gradients(A, Fd, T):
    """Autograd through the oracle's lstm_forward + MSE against the reference's get_loss(...).backward() in eval mode
    (oracle/gen_golden_grads.py; lstm_step_controller.py:321-337): the parity gate of the LSTM BPTT kernel (row a12)."""
    gg = U.golden(f"lstm_grads_A{A}_F{Fd}_T{T}")
    mods = {
        "force_encoder": syn.synth_state_dict(shp.mlp_shapes([Fd, 128, 128]), 41, "lstm.force_encoder."),
        "lstm": syn.synth_state_dict(shp.lstm_shapes(128 + A), 41, "lstm.lstm."),
        "output_head": syn.synth_state_dict(shp.lstm_head_shapes(256, A), 41, "lstm.output_head."),
    }
    mods = {m: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for m, sd in mods.items()}
    st = syn.synth_stats_varied(A, 41)
    vla_n = orc.normalize_actions(syn.det_uniform("lstm.vla", (3, T, A), 41, -1.0, 1.0), st, "vla")
    forces = syn.det_normal("lstm.forces", (3, T, Fd), 41)
    cond = syn.det_normal("lstm.cond", (3, 256), 41).requires_grad_(True)
    expert = syn.det_uniform("lstm.exp", (3, T, A), 41, -1.0, 1.0)
    loss = torch.nn.functional.mse_loss(orc.lstm_forward(mods, vla_n, cond, forces), expert)
    torch.testing.assert_close(loss.detach(), gg["loss"], rtol=1e-5, atol=1e-6)
    loss.backward()
    torch.testing.assert_close(cond.grad, gg["d_cond"], rtol=1e-4, atol=1e-6)
    for i, n in enumerate(str(x) for x in gg["names"]):
        m, key = n.split(".", 1)
        gr = mods[m][key].grad.flatten().double()
        scale = max(float(gg["norm"][i]), 1e-12)
        assert abs(float(gr.norm()) - float(gg["norm"][i])) <= 1e-4 * scale, n
        assert abs(float(gr.sum()) - float(gg["sum"][i])) <= 1e-3 * scale, n
        k = min(8, gr.numel())
        assert float((gr[:k] - gg["head"][i][:k].double()).abs().max()) <= 1e-4 * scale, n


def test_explicit_unet_backward_matches_autograd():
    """oracle/vt_oracle_bwd.py (per-tap dgrad / wgrad, fused GroupNorm+Mish backward, FiLM, skips: the op decomposition of the
    round-2 kernels) against autograd through the forward oracle, every parameter, the condition and the input."""
    from oracle import vt_oracle_bwd as ob
    A, T, B = 7, 16, 2
    sd = {k: v.double().requires_grad_(True) for k, v in U.net_sd(A, 21, "v_net").items()}
    x = syn.det_uniform("bwd.x", (B, T, A), 5, -1.0, 1.0).double().requires_grad_(True)
    cond = syn.det_normal("bwd.cond", (B, 256), 5).double().requires_grad_(True)
    t = torch.tensor([0.3, 0.9], dtype=torch.float64)
    dout = syn.det_normal("bwd.dout", (B, T, A), 5).double()
    out = orc.unet_forward(sd, x, t, cond)
    out.backward(dout)
    with torch.no_grad():
        out2, cache = ob.unet_forward_cached({k: v.detach() for k, v in sd.items()}, x.detach(), t, cond.detach())
        grads, dcond, dx = ob.unet_backward({k: v.detach() for k, v in sd.items()}, cache, dout)
    torch.testing.assert_close(out2, out.detach(), rtol=1e-9, atol=1e-9)
    torch.testing.assert_close(dcond, cond.grad, rtol=1e-7, atol=1e-9)
    torch.testing.assert_close(dx, x.grad, rtol=1e-7, atol=1e-9)
    assert sorted(grads) == sorted(sd)
    for k in sd:
        torch.testing.assert_close(grads[k], sd[k].grad, rtol=1e-7, atol=1e-8, msg=k)


@pytest.mark.parametrize("A,T", [(10, 16), (7, 64)])
def test_explicit_loss_backward_matches_reference_digests(A, T):
    """The explicit backward of get_loss (three nets, v / s / b targets) against the reference's loss.backward() fixtures."""
    from oracle import vt_oracle_bwd as ob
    g, gg = U.golden(f"loss_A{A}_T{T}"), U.golden(f"loss_grads_A{A}_T{T}")
    with torch.no_grad():
        loss, grads, dcond = ob.bridge_loss_backward(U.net_sd(A, 21), syn.det_normal("loss.cond", (3, 256), 24),
                                                     syn.det_uniform("loss.exp", (3, T, A), 24, -1.0, 1.0),
                                                     syn.det_uniform("loss.vla", (3, T, A), 24, -1.0, 1.0), g["step"], g["z_unit"])
    assert abs(float(loss) - float(g["loss"])) <= 1e-4 * max(1.0, abs(float(g["loss"])))
    torch.testing.assert_close(dcond, gg["d_cond"], rtol=1e-3, atol=1e-4 * float(gg["d_cond"].abs().max()))
    for i, n in enumerate(str(x) for x in gg["names"]):
        gr = grads[n].flatten().double()
        scale = max(float(gg["norm"][i]), 1e-12)
        assert abs(float(gr.norm()) - float(gg["norm"][i])) <= 1e-3 * scale, n
        k = min(8, gr.numel())
        assert float((gr[:k] - gg["head"][i][:k].double()).abs().max()) <= 1e-3 * scale, n


@pytest.mark.parametrize("A,Fd,T", [(10, 3, 16), (7, 64, 32)])
def test_explicit_lstm_bptt_matches_reference_digests(A, Fd, T):
    """oracle/vt_oracle_bwd.lstm_loss_backward (batched input / head GEMMs + sequential recurrence backward) against the
    reference's get_loss().backward() fixtures (lstm_step_controller.py:321-337, eval mode)."""
    from oracle import vt_oracle_bwd as ob
    gg = U.golden(f"lstm_grads_A{A}_F{Fd}_T{T}")
    mods = {
        "force_encoder": syn.synth_state_dict(shp.mlp_shapes([Fd, 128, 128]), 41, "lstm.force_encoder."),
        "lstm": syn.synth_state_dict(shp.lstm_shapes(128 + A), 41, "lstm.lstm."),
        "output_head": syn.synth_state_dict(shp.lstm_head_shapes(256, A), 41, "lstm.output_head."),
    }
    st = syn.synth_stats_varied(A, 41)
    vla_n = orc.normalize_actions(syn.det_uniform("lstm.vla", (3, T, A), 41, -1.0, 1.0), st, "vla")
    with torch.no_grad():
        loss, grads, dcond = ob.lstm_loss_backward(mods, vla_n, syn.det_normal("lstm.cond", (3, 256), 41),
                                                   syn.det_normal("lstm.forces", (3, T, Fd), 41),
                                                   syn.det_uniform("lstm.exp", (3, T, A), 41, -1.0, 1.0))
    torch.testing.assert_close(loss, gg["loss"], rtol=1e-5, atol=1e-6)
    torch.testing.assert_close(dcond, gg["d_cond"], rtol=1e-3, atol=1e-6)
    for i, n in enumerate(str(x) for x in gg["names"]):
        gr = grads[n].flatten().double()
        scale = max(float(gg["norm"][i]), 1e-12)
        assert abs(float(gr.norm()) - float(gg["norm"][i])) <= 1e-3 * scale, n
        k = min(8, gr.numel())
        assert float((gr[:k] - gg["head"][i][:k].double()).abs().max()) <= 1e-3 * scale, n


# ---- pad_and_resize_for_siglip (scripts/utils_eef.py:44-77): the numpy restatement of cv2 INTER_AREA against cv2's own outputs ----
def test_resize_oracle_reproduces_cv2_goldens():
    import hashlib
    import numpy as np
    from oracle import resize_oracle as ro
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import gen_golden_resize as gg
    g = np.load(os.path.join(ROOT, "tests", "golden", "resize_small.npz"))
    for name, h, w, t in gg.CASES:
        got = ro.pad_and_resize_for_siglip(g[f"{name}_in"], t)
        assert got.shape == (t, t, 3) and got.dtype == np.uint8
        assert np.array_equal(got, g[f"{name}_out"]), name                      # bit-exact: byte work
    d = np.load(os.path.join(ROOT, "tests", "golden", "resize_digests.npz"))
    for name, h, w, t, seed in gg.BIG:
        got = ro.pad_and_resize_for_siglip(gg.frame(h, w, seed), t)
        assert np.array_equal(np.frombuffer(hashlib.sha256(got.tobytes()).digest(), dtype=np.uint8), d[name]), name
    assert ro.pad_and_resize_for_siglip(None) is None
