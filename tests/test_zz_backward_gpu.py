"""GPU parity of the backward building blocks (row a10 of SURVEY 8, first slice): the dgrad / wgrad plans of unet_bwd.py and
the GroupNorm + Mish (+ FiLM) backward kernel run through the C ABI on the B200 and are held to oracle/vt_oracle_bwd.py
(which is pinned to autograd and to the reference's own loss.backward() digests).  bf16 operands, fp32 accumulation: the
gate is 2e-2 of the gradient's scale per tensor (north_star: 5e-2 relative on bf16).  Named test_zz_* so that it runs last."""
import os
import sys

import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import bwd_cases  # noqa: E402

pytestmark = pytest.mark.gpu
DEV = torch.device("cuda", 0)


def _run(plan):
    plan.compile().run()
    torch.cuda.synchronize()


@pytest.mark.parametrize("kind", ["k5", "down", "up", "k1in"])
def test_wgrad_on_the_tensor_cores(kind):
    plan, check = bwd_cases.wgrad_case(kind, DEV)
    _run(plan)
    check()


@pytest.mark.parametrize("film", [False, True])
def test_conv_block_backward(film):
    plan, check = bwd_cases.block_case(DEV, film)
    _run(plan)
    check()


def test_conv_block_backward_wide_and_long():
    """512 channels (two channels per thread in gn_mish_bwd_kernel, 64-channel groups) and T = 64 (the BASELINE horizon)."""
    plan, check = bwd_cases.block_case(DEV, True, G=3, B=4, T=64, Ci=256, Co=512, seed=6)
    _run(plan)
    check()


def test_colsum():
    plan, check = bwd_cases.colsum_case(DEV)
    _run(plan)
    check()


@pytest.mark.parametrize("kind", ["k5", "down", "up"])
def test_dgrad_on_the_tensor_cores(kind):
    plan, check = bwd_cases.dgrad_case(kind, DEV)
    _run(plan)
    check()


@pytest.mark.parametrize("ci,co", [(256, 256), (256, 512), (1024, 512), (7, 256)])
def test_res_block_backward(ci, co):
    """ConditionalResidualBlock1D backward (conditional_unet_1D.py:58-105): identity and 1x1-conv residual, concatenated
    (1024-channel) input, the network's first block."""
    plan, check = bwd_cases.res_block_case(DEV, ci, co)
    _run(plan)
    check()


def test_unet_backward():
    """Whole U-Nets: training forward + explicit backward (146 parameter gradients per net, d global_cond) on the B200."""
    plan, check = bwd_cases.unet_case(DEV)
    _run(plan)
    res = check()
    assert res["missing"] == []


@pytest.mark.parametrize("A,T", [(10, 16), (7, 64)])
def test_get_loss_backward_against_the_reference_gradients(A, T):
    """get_loss(...).backward() of the reference (bridge_model.py:220-246) as one program on the B200: loss values, d obs_cond,
    and all 438 parameter gradients of b_net / v_net / s_net against the reference's own digests and in full."""
    plan, check = bwd_cases.loss_case(DEV, A, T)
    _run(plan)
    assert check()["tensors"] == 439


def test_training_step_autograd_contract():
    """loss, info = get_loss(batch); loss.backward() on the B200: .grad of all 438 net parameters, the gradient flowing into the
    producer of obs_cond, and the re-packed operand copies after an in-place parameter update."""
    res = bwd_cases.training_step_case(DEV)()
    assert res["first"] <= 3e-2 and res["after_update"] <= 3e-2


def test_differentiable_encode_observation_feeds_the_state_encoder():
    """encode_observation(differentiable=True) equals the native inference path (bf16 gate) and carries the graph of the
    state encoder, so that get_loss(...).backward() trains it (bridge_train.py:151,315-334)."""
    import vt_testutil as U
    c = U.predict_case("predict_cfg2_B3_dark_varstats")        # 2 ViT layers, batch 3
    ctl = U.make_controller(c, "cuda:0", precise=False)
    args = (c["state"].to(DEV), c["img1"], c["img2"], c["forces"].to(DEV))
    with torch.no_grad():                                      # validation / deployment: the native inference program
        ref = ctl.encode_observation(*args)
    assert not ref.requires_grad
    cond = ctl.encode_observation(*args)                       # grad mode on + trainable encoder: the training path (bridge_train.py:151)
    assert cond.requires_grad and cond.shape == ref.shape
    assert float((cond - ref).abs().max()) <= 5e-2 * float(ref.abs().max())
    si = ctl.diffusion_model
    T, A = c["T"], c["A"]
    batch = {"obs_cond": cond, "expert_act": torch.zeros(cond.shape[0], T, A, device=DEV), "vla_act": c["vla"].to(DEV)}
    loss, _ = si.get_loss(batch, DEV)
    loss.backward()
    g = [p.grad for p in ctl.state_encoder.parameters()]
    assert all(x is not None and torch.isfinite(x).all() for x in g) and float(g[0].abs().max()) > 0
    assert all(p.grad is not None for p in si.net.parameters())


@pytest.mark.parametrize("kind", ["k5", "down", "up"])
def test_split_k_wgrad(kind):
    plan, check = bwd_cases.wgrad_case(kind, DEV, B=4, split_k=2)
    _run(plan)
    check()


def test_full_size_backward_is_additive_over_the_batch():
    """BASELINE batch (256 x T 64 x A 7, three nets): gradients of the full batch == mean of the gradients of its halves."""
    res = bwd_cases.batch_additivity_case(DEV)()
    assert res["tensors"] == 439


def test_lstm_layers_bptt():
    """Stacked nn.LSTM layers: training forward + back-propagation through time against torch.nn.LSTM autograd (CPU)."""
    plan, check = bwd_cases.lstm_layers_case(DEV)
    _run(plan)
    check()


@pytest.mark.parametrize("A,Fd,T", [(10, 3, 16), (7, 64, 32)])
def test_lstm_get_loss_backward_against_the_reference_gradients(A, Fd, T):
    plan, check = bwd_cases.lstm_loss_case(DEV, A, Fd, T)
    _run(plan)
    assert check()["tensors"] == 18


def test_lstm_training_with_dropout():
    plan, check = bwd_cases.lstm_dropout_case(DEV)
    _run(plan)
    check()


def test_native_encoder_training():
    import torch.nn as nn
    from vla_touch_b200 import mlp_train as mt
    g = torch.Generator().manual_seed(5)
    enc = nn.Sequential(nn.Linear(839, 256), nn.GELU(), nn.Linear(256, 256), nn.GELU(), nn.Linear(256, 256)).to(DEV)
    x, dout = torch.randn(6, 839, generator=g).to(DEV), torch.randn(6, 256, generator=g).to(DEV)
    out = mt.encoder_forward(enc, {}, x)
    (out * dout).sum().backward()
    got = {n: p.grad.clone() for n, p in enc.named_parameters()}
    enc.zero_grad()
    ref = enc(x)
    (ref * dout).sum().backward()
    assert float((out - ref).detach().abs().max()) <= 2e-2 * float(ref.detach().abs().max())
    for n, p in enc.named_parameters():
        assert float((got[n] - p.grad).abs().max()) <= 3e-2 * float(p.grad.abs().max()), n


def _train_batch(c, B):
    from vla_touch_b200 import synthetic as syn
    T, A, Fd, hw = c["T"], c["A"], c["F"], c["hw"]
    b = {"states": syn.det_normal("tt.states", (B, 2 + T, A), 3), "forces": syn.det_normal("tt.forces", (B, 2 + T, Fd), 3),
         "vla_actions": syn.det_uniform("tt.vla", (B, T, A), 3, -1.0, 1.0),
         "images_cam1": syn.synth_images_u8("tt.cam1", B, hw, 3)[:, None].contiguous(),
         "images_cam2": syn.synth_images_u8("tt.cam2", B, hw, 3)[:, None].contiguous()}
    b["expert_actions"] = (b["vla_actions"] + 0.1 * syn.det_normal("tt.delta", (B, T, A), 3)).clamp(-1, 1)
    return b


def test_trainer_step_equals_the_reference_loop_on_the_dropin_api():
    """trainer.DiffusionControllerTrainer.train_step (gradients read in place from the backward program's arena) produces the
    same parameters, EMA shadows and loss as the reference's own loop (bridge_train.py:296-337) run on the drop-in classes:
    zero_grad -> get_loss -> loss.backward() -> optimizer.step (+ EMA, cosine LR)."""
    import vt_testutil as U
    from vla_touch_b200.optim import FusedAdamWEMA
    from vla_touch_b200.trainer import DiffusionControllerTrainer
    c = U.predict_case("predict_cfg2_B3_dark_varstats")
    B, T, A = 3, c["T"], c["A"]
    batch = _train_batch(c, B)
    g = torch.Generator().manual_seed(9)
    step, z = torch.rand(B, generator=g).to(DEV), torch.randn(B, T, A, generator=g).to(DEV)
    ctl_a, ctl_b = U.make_controller(c, "cuda:0"), U.make_controller(c, "cuda:0")
    for ctl in (ctl_a, ctl_b):
        ctl.diffusion_model.step_override, ctl.diffusion_model.z_override = step, z
    tr_a = DiffusionControllerTrainer(ctl_a, c["stats"], device="cuda:0")
    out_a = [tr_a.train_step({k: v.clone() for k, v in batch.items()}) for _ in range(2)]
    # the reference loop on the drop-in API
    tr_b = DiffusionControllerTrainer(ctl_b, c["stats"], device="cuda:0")      # only for _prepare_batch_for_diffusion
    dm = ctl_b.diffusion_model
    net_params = list(dm.net.parameters())
    opt = FusedAdamWEMA(net_params + list(ctl_b.state_encoder.parameters()), lr=1e-4, weight_decay=1e-6, ema=dm.ema,
                        ema_params=net_params, t_max=100000)
    out_b = []
    for _ in range(2):
        ctl_b.train()
        bd = tr_b._prepare_batch_for_diffusion({k: v.clone() for k, v in batch.items()})
        opt.zero_grad()
        loss, info = dm.get_loss(bd, "cuda:0")
        loss.backward()
        opt.step()
        out_b.append(loss.detach())
    for a, b in zip(out_a, out_b):
        assert torch.isfinite(a["loss"]) and float((a["loss"] - b).abs()) <= 1e-5 * max(1.0, float(b.abs()))
    assert float(out_a[0]["loss"]) != float(out_a[1]["loss"])                  # the second step saw the updated weights
    worst = 0.0
    for (n, pa), pb in zip(ctl_a.diffusion_model.net.named_parameters(), ctl_b.diffusion_model.net.parameters()):
        worst = max(worst, float((pa - pb).abs().max()))
        assert torch.equal(pa, pb), (n, float((pa - pb).abs().max()))
    for sa, sb in zip(ctl_a.diffusion_model.ema.shadow_params, ctl_b.diffusion_model.ema.shadow_params):
        assert torch.equal(sa, sb)
    for pa, pb in zip(ctl_a.state_encoder.parameters(), ctl_b.state_encoder.parameters()):
        assert float((pa - pb).abs().max()) <= 1e-6
    assert ctl_a.diffusion_model.ema.num_updates == 2 and tr_a.optimizer.step_count == 2


def test_lstm_trainer_steps_reduce_the_loss():
    """lstm_train.py:122-139 on the drop-in controller: `loss = controller.get_loss(batch)` is differentiable in training mode
    without any flag, obs_encoder trains through d obs_cond, and a few optimizer steps on one minibatch reduce the loss."""
    import vt_testutil as U
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.lstm_step_controller import TactileLSTMController
    from vla_touch_b200.trainer import LSTMControllerTrainer
    c = U.predict_case("predict_cfg2_B3_dark_varstats")
    A, Fd = c["A"], c["F"]
    lc = TactileLSTMController(state_dim=A, hidden_dim=256, num_layers=2, dropout=0.1, device="cuda:0", force_dim=Fd,
                               image_state_dict=c["dino"])
    for nm, mod in (("obs_encoder", lc.obs_encoder), ("force_encoder", lc.force_encoder), ("lstm", lc.lstm), ("output_head", lc.output_head)):
        syn.fill_named_(mod.named_parameters(), 41, prefix=f"lstm.{nm}.")
    tr = LSTMControllerTrainer(lc, c["stats"], learning_rate=1e-3, device="cuda:0")
    batch = _train_batch(dict(c, T=16), 3)
    w0 = lc.obs_encoder[0].weight.detach().clone()
    losses = [float(tr.train_step({k: v.clone() for k, v in batch.items()})) for _ in range(8)]
    assert all(l == l for l in losses) and losses[-1] < losses[0], losses
    assert not torch.equal(w0, lc.obs_encoder[0].weight)                       # the observation encoder received gradients
    # encode_force (lstm_step_controller.py:148-168): sequence and single-step inputs
    lc.eval()
    f = syn.det_normal("tt.f", (2, 5, Fd), 1).to(DEV)
    with torch.no_grad():
        ref = lc.force_encoder(f.reshape(-1, Fd)).reshape(2, 5, -1)
    got = lc.encode_force(f)
    assert got.shape == ref.shape and float((got - ref).abs().max()) <= 3e-2 * float(ref.abs().max())
    assert lc.encode_force(f[:, 0]).shape == (2, 128)


def test_no_visual_controller_predicts_and_trains():
    """bridge_controller_no_visual.DiffusionController (reference :16-140): obs = cat(state, force), images optional / ignored."""
    import vt_testutil as U
    from oracle import vt_oracle as orc
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.bridge_controller_no_visual import DiffusionController
    from vla_touch_b200.dropin import bridge_controller_no_visual as shim
    assert shim.DiffusionController is DiffusionController
    A, Fd, T, B = 7, 64, 16, 4
    ma = {'interpolant_type': 'linear', 'gamma_type': '2^0.5*t(t-1)', 'epsilon_type': '1-t', 'prior_policy': 'vla', 'beta_max': 0.03,
          'sde_type': 'vs', 'action_dim': A, 'obs_dim': 256, 'obs_horizon': 1, 'net_type': 'unet1D_si', 'pretrain': False,
          'context_frames': 2, 'horizon': T}
    ctl = DiffusionController(state_dim=A, hidden_dim=256, diffusion_steps=10, device="cuda:0", model_args=ma, use_force=True, force_dim=Fd)
    assert ctl.image_encoder is None and ctl.obs_dim == A + Fd
    enc = U.enc_sd(A + Fd, 5)
    ctl.state_encoder.load_state_dict(enc)
    ctl.diffusion_model.net.load_state_dict(U.net_sd(A, 5))
    ctl.diffusion_model.ema = type(ctl.diffusion_model.ema)(ctl.diffusion_model.net.parameters(), decay=0.75)
    ctl.stats = {k: v.to(DEV) for k, v in syn.synth_stats(A).items()}
    state, forces = syn.det_normal("nv.s", (B, A), 1), syn.det_normal("nv.f", (B, Fd), 1)
    vla = syn.det_uniform("nv.v", (B, T, A), 1, -1.0, 1.0)
    noise = syn.det_normal("nv.n", (10, B, T, A), 1)
    ctl.noise_override = noise.to(DEV)
    out = ctl.predict(state.to(DEV), vla.to(DEV), None, None, forces.to(DEV)).cpu()
    with torch.no_grad():                                      # CPU oracle of the same path: encoder over cat(state, force), sde_vs
        cond = orc.mlp3_gelu(enc, torch.cat((state, forces), -1))
        xn = orc.normalize_actions(vla, syn.synth_stats(A), "vla")
        ref = orc.denormalize_actions(orc.sde_vs(U.net_sd(A, 5, "v_net"), U.net_sd(A, 5, "s_net"), xn, cond, 10, 0.03, noise),
                                      syn.synth_stats(A), "expert")
    assert float((out - ref).abs().max()) <= 5e-2 * float(ref.abs().max())
    assert out.shape == (B, T, A) and torch.isfinite(out).all()
    cond_t = ctl.encode_observation(state.to(DEV), None, None, forces.to(DEV))
    assert cond_t.requires_grad and cond_t.shape == (B, 256)


@pytest.mark.parametrize("B,T", [(40, 24), (256, 8)])
def test_lstm_tensor_core_recurrence_equals_the_cuda_core_kernels(B, T, monkeypatch):
    """csrc/vt_lstm_tc.cuh (cluster of 8 CTAs per 128 rows, W_hh resident in shared memory, tcgen05 MMA per step) against the
    CUDA-core recurrence kernels of csrc/vt_lstm.cuh on the same program: loss, every one of the 18 parameter gradients and
    d obs_cond of get_loss forward + BPTT, and the inference forward.  Ragged row block (40 of 128) and two clusters (256 rows).
    The kernels differ in operand precision (bf16 h / d gates on the tensor cores) and in the MUFU tanh: gate 2e-2 of scale."""
    from vla_touch_b200 import shapes as shp
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.lstm_step_controller import LstmEngine
    from vla_touch_b200.lstm_train import LstmLossBackwardProgram
    A, Fd = 7, 64
    mods = {"force_encoder": syn.synth_state_dict(shp.mlp_shapes([Fd, 128, 128]), 41, "lstm.force_encoder."),
            "lstm": syn.synth_state_dict(shp.lstm_shapes(128 + A), 41, "lstm.lstm."),
            "output_head": syn.synth_state_dict(shp.lstm_head_shapes(256, A), 41, "lstm.output_head.")}
    vla, f = syn.det_uniform("lt.vla", (B, T, A), 1, -1.0, 1.0), syn.det_normal("lt.f", (B, T, Fd), 1)
    cond, exp = syn.det_normal("lt.cond", (B, 256), 1), syn.det_uniform("lt.exp", (B, T, A), 1, -1.0, 1.0)
    res = {}
    for tc in ("1", "0"):
        monkeypatch.setenv("VT_LSTM_TC", tc)
        lp = LstmLossBackwardProgram(mods, A, Fd, B, T, DEV)
        lp.set_inputs(vla, f, cond, exp)
        lp.run()
        torch.cuda.synchronize()
        eng = LstmEngine(mods, A, Fd, 256, 2, B, T, DEV, False, False)
        eng.vla.copy_(vla); eng.forces.copy_(f); eng.cond.copy_(cond)
        eng.run()
        torch.cuda.synchronize()
        res[tc] = (lp.loss(), {k: v.detach().float().clone() for k, v in lp.grads.items()}, lp.d_cond.clone(), eng.out.clone())
    (l1, g1, d1, o1), (l0, g0, d0, o0) = res["1"], res["0"]
    rel = lambda a, b: float((a - b).abs().max()) / max(float(b.abs().max()), 1e-12)
    assert abs(l1 - l0) <= 2e-2 * abs(l0)
    assert rel(o1, o0) <= 2e-2 and rel(d1, d0) <= 2e-2
    worst = {k: rel(g1[k], g0[k]) for k in g0}
    assert len(worst) == 18 and max(worst.values()) <= 2e-2, worst


def test_gather_repack_equals_the_tensor_op_repack(monkeypatch):
    """After an optimizer step the packed GEMM operands are rebuilt by ONE gather launch over index maps (unet_train.setup_gather,
    csrc gather_repack_kernel).  Every operand must equal what the packing functions produce from the new parameter values, bit for
    bit; the parameters must still be the same Parameter objects with the same values; a parameter that leaves the arena
    (p.data replaced) switches the program back to the tensor-op re-pack."""
    import vt_testutil as U
    c = U.predict_case("predict_cfg2_B3_dark_varstats")
    ctl = U.make_controller(c, "cuda:0", precise=False)
    si = ctl.diffusion_model
    B, T, A = 3, c["T"], c["A"]
    cond = torch.randn(B, 256, device=DEV)
    batch = {"obs_cond": cond, "expert_act": torch.zeros(B, T, A, device=DEV), "vla_act": c["vla"].to(DEV)}
    params = list(si.net.parameters())
    before = [p.detach().clone() for p in params]
    loss, _ = si.get_loss(batch, DEV)
    loss.backward()
    with torch.no_grad():                                    # "optimizer step": in-place update bumps the version counters
        for p in params:
            p.add_(0.01 * torch.randn_like(p))
    want = [p.detach().clone() for p in params]
    loss2, _ = si.get_loss(batch, DEV)                       # first re-pack: sets the gather up (verified against the tensor ops)
    prog = si.train_program(B, T)
    assert getattr(prog, "_gather", None) is not None and prog.gather_valid()
    assert all(p is q for p, q in zip(params, si.net.parameters()))
    assert all(torch.equal(p.detach(), w) for p, w in zip(params, want)) and not torch.equal(want[0], before[0])
    with torch.no_grad():
        for p in params:
            p.mul_(1.0 + 0.05 * torch.rand_like(p))
    loss3, _ = si.get_loss(batch, DEV)                       # second re-pack: the gather launch alone
    sds = si._net_state_dicts()
    fresh = prog.W._pack(sds)
    for k, t in prog.W.t.items():
        assert torch.equal(t, fresh[k]), k
    for dest, fn in prog.W.repack:
        assert torch.equal(dest, fn(sds))
    assert torch.isfinite(loss3) and float(loss3) != float(loss2)
    with torch.no_grad():                                    # a parameter leaves the arena
        params[5].data = params[5].data.clone()
        params[5].add_(0.5)
    assert not prog.gather_valid()
    loss4, _ = si.get_loss(batch, DEV)
    assert getattr(prog, "_gather", None) is None
    fresh = prog.W._pack(si._net_state_dicts())
    for k, t in prog.W.t.items():
        assert torch.equal(t, fresh[k]), k


def test_saved_raw_conv_outputs_equal_the_recomputed_ones(monkeypatch):
    """vt_gemm_desc.raw_out (ABI v6): the training forward's GroupNorm epilogue also stores conv + bias as fp32, so the backward needs
    no recomputation GEMM.  The stored values are the same accumulators + bias the recompute produced: every gradient, d obs_cond
    and the losses are bit-identical between the two forms of the program, and the new form has 25 GEMM ops fewer."""
    from vla_touch_b200 import shapes as shp
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.params import sub_state_dict
    from vla_touch_b200.unet_train import LossBackwardProgram
    A, T, B = 7, 32, 5
    full = syn.synth_state_dict(shp.si_net_shapes(A, 256), 77, prefix="net.")
    sds = [sub_state_dict(full, p) for p in ("b_net.", "v_net.", "s_net.")]
    g = torch.Generator().manual_seed(77)
    inputs = (torch.rand(B, T, A, generator=g) * 2 - 1, torch.rand(B, T, A, generator=g) * 2 - 1, torch.randn(B, 256, generator=g),
              torch.rand(B, generator=g), torch.randn(B, T, A, generator=g))
    res = {}
    for mode in ("1", "0"):
        monkeypatch.setenv("VT_TRAIN_RECOMPUTE", mode)
        lp = LossBackwardProgram(sds, A, B, T, 0.03, DEV)
        lp.set_inputs(*inputs)
        out = lp.run().clone()
        res[mode] = (out, {k: v.clone() for k, v in lp.grads.items()}, lp.d_cond.clone(), len(lp.plan))
    (o1, g1, d1, n1), (o0, g0, d0, n0) = res["1"], res["0"]
    assert n1 - n0 == 25 * 1, (n1, n0)                      # 12 blocks x 2 + final_conv.0, one grouped GEMM each
    assert torch.equal(o1, o0) and torch.equal(d1, d0)
    for k in g1:
        assert torch.equal(g1[k], g0[k]), k
