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
    ref = ctl.encode_observation(*args)
    assert not ref.requires_grad
    cond = ctl.encode_observation(*args, differentiable=True)
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


@pytest.mark.xfail(strict=False, reason="first run of the backward kernels at the BASELINE batch (256 x 64 x 7): written after the "
                                        "round's GPU budget ended, never run on a B200 at this size")
def test_full_size_backward_is_additive_over_the_batch():
    """BASELINE batch (256 x T 64 x A 7, three nets): gradients of the full batch == mean of the gradients of its halves."""
    res = bwd_cases.batch_additivity_case(DEV)()
    assert res["tensors"] == 439


@pytest.mark.xfail(strict=False, reason="lstm_seq_train_kernel / lstm_bwd_kernel were written after the round's GPU budget ended: "
                                        "compiled for sm_100a and checked on the CPU descriptor interpreter, never run on a B200 yet")
def test_lstm_layers_bptt():
    """Stacked nn.LSTM layers: training forward + back-propagation through time against torch.nn.LSTM autograd (CPU)."""
    plan, check = bwd_cases.lstm_layers_case(DEV)
    _run(plan)
    check()


@pytest.mark.xfail(strict=False, reason="uses lstm_seq_train_kernel / lstm_bwd_kernel / ln_gelu_bwd_kernel and the new ewise ops, all "
                                        "written after the round's GPU budget ended: never run on a B200 yet")
@pytest.mark.parametrize("A,Fd,T", [(10, 3, 16), (7, 64, 32)])
def test_lstm_get_loss_backward_against_the_reference_gradients(A, Fd, T):
    plan, check = bwd_cases.lstm_loss_case(DEV, A, Fd, T)
    _run(plan)
    assert check()["tensors"] == 18


@pytest.mark.xfail(strict=False, reason="LSTM training kernels + dropmask_kernel: written after the round's GPU budget ended, never run on a B200")
def test_lstm_training_with_dropout():
    plan, check = bwd_cases.lstm_dropout_case(DEV)
    _run(plan)
    check()


@pytest.mark.xfail(strict=False, reason="mlp_train uses the ewise gelu' op added after the round's GPU budget ended: never run on a B200")
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
