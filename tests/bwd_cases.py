"""TEST INFRASTRUCTURE ONLY -- the backward-plan cases shared by the CPU (descriptor interpreter) and the GPU (native kernels)
tests: each builder returns (plan, check) where check() compares the plan's output buffers with oracle/vt_oracle_bwd.py."""
import torch

from oracle import vt_oracle_bwd as ob
from vla_touch_b200 import unet_bwd as ub
from vla_touch_b200.plan import Plan
from vla_touch_b200.unet import _View

bf = lambda t: t.to(torch.bfloat16).float()


def _rel(got, ref):
    return (got - ref).abs().max().item() / max(ref.abs().max().item(), 1e-12)


def wgrad_case(kind: str, device, G=2, B=3, C=256, seed=11, split_k=None):
    """conv_wgrad for 'k5' (Conv1d k5 p2), 'down' (Conv1d k3 s2 p1), 'up' (ConvTranspose1d k4 s2 p1), 'k1in' (1x1 conv on a
    64-channel padded 7-channel input shared by all nets: the first block's residual_conv)."""
    g = torch.Generator().manual_seed(seed)
    plan = Plan(device)
    ctx = ub.DgradCtx(G, precise=False)
    T = 16
    if kind == "k1in":
        Cx, Tx, Ty, shared = 64, T, T, True
        x = plan.buf("x", (B, Tx, Cx), torch.bfloat16)
        x[:, :, :7] = torch.randn(B, Tx, 7, generator=g).to(device)
    else:
        Cx, Tx, shared = C, T, False
        Ty = {"k5": T, "down": T // 2, "up": 2 * T}[kind]
        x = plan.buf("x", (G, B, Tx, Cx), torch.bfloat16)
        x.copy_(torch.randn(G, B, Tx, Cx, generator=g))
    dy = plan.buf("dy", (G, B, Ty, C), torch.bfloat16)
    dy.copy_(torch.randn(G, B, Ty, C, generator=g))
    vx, vy = _View(x, Tx, Cx, shared=shared), _View(dy, Ty, C)
    if kind == "k5":
        dw = ub.conv_wgrad(plan, ctx, B, vy, vx, tap_off=[k - 2 for k in range(5)], t_out=T, split_k=split_k)
    elif kind == "k1in":
        dw = ub.conv_wgrad(plan, ctx, B, vy, vx, tap_off=[0], t_out=T)
    elif kind == "down":
        dw = ub.conv_wgrad(plan, ctx, B, vy, vx, tap_off=[k - 1 for k in range(3)], stride=2, t_out=Ty, split_k=split_k)
    else:
        dw = ub.conv_wgrad(plan, ctx, B, vx, vy, tap_off=[k - 1 for k in range(4)], stride=2, t_out=Tx, split_k=split_k)

    def check(tol=1e-2):
        for n in range(G):
            xn = (x if shared else x[n]).float().cpu().permute(0, 2, 1)
            dyn = dy[n].float().cpu().permute(0, 2, 1)
            if kind == "k5":
                ref = ob.conv1d_bwd(xn, torch.zeros(C, Cx, 5), dyn, padding=2)[1]
                got = ub.unpack_wgrad(dw, Cx, 5)[n]
            elif kind == "k1in":
                ref = ob.conv1d_bwd(xn[:, :7], torch.zeros(C, 7, 1), dyn)[1]
                got = ub.unpack_wgrad(dw, 7, 1)[n]
            elif kind == "down":
                ref = ob.conv1d_bwd(xn, torch.zeros(C, Cx, 3), dyn, stride=2, padding=1)[1]
                got = ub.unpack_wgrad(dw, Cx, 3)[n]
            else:
                ref = ob.convT1d_bwd(xn, torch.zeros(Cx, C, 4), dyn)[1]
                got = ub.unpack_wgrad(dw, C, 4)[n]
            e = _rel(got.float().cpu(), ref)
            assert e <= tol, (kind, n, e)
    return plan, check


def block_case(device, film: bool, G=2, B=3, T=16, Ci=256, Co=256, seed=5):
    """conv_block_backward of Conv1d(k5) -> GroupNorm(8) -> Mish [-> FiLM] against gn_mish_bwd + conv1d_bwd of the oracle."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    plan = Plan(device)
    ctx = ub.DgradCtx(G, precise=False)
    ws = [r(Co, Ci, 5) / (5 * Ci) ** 0.5 for _ in range(G)]
    bs = [0.1 * r(Co) for _ in range(G)]
    gam = [1 + 0.2 * r(Co) for _ in range(G)]
    bet = [0.2 * r(Co) for _ in range(G)]
    x = plan.buf("x", (G, B, T, Ci), torch.bfloat16)
    x.copy_(r(G, B, T, Ci))
    dout = plan.buf("dout", (G, B, T, Co), torch.float32)
    dout.copy_(r(G, B, T, Co))
    dx = plan.buf("dx", (G, B, T, Ci), torch.bfloat16)
    fl = None
    if film:
        LD, OFF = 4 * Co, Co          # a table wider than the block's slice, like the stacked FiLM table of the 12 blocks
        ft = plan.buf("film", (G, B, LD), torch.float32)
        ft.copy_(1 + 0.3 * r(G, B, LD))
        dft = plan.buf("dfilm", (G, B, LD), torch.float32)
        fl = (ft, dft, OFF)
    out = ub.conv_block_backward(plan, ctx, B, _View(x, T, Ci), ws, bs, gam, bet, dout, _View(dx, T, Ci), film=fl, tag="blk")

    def check(tol=2e-2):
        errs = {}
        for n in range(G):
            xn = x[n].float().cpu().permute(0, 2, 1)
            w = bf(ws[n])
            raw = ob.conv1d_fwd(xn, w, bs[n], padding=2)
            do = dout[n].float().cpu().permute(0, 2, 1)
            if film:
                scale = ft[n, :, OFF: OFF + Co].float().cpu()[:, :, None]
                y0 = ob.gn_mish_fwd(raw, gam[n], bet[n])
                demb = torch.cat([(do * y0).sum(-1), do.sum(-1)], dim=1)
                errs["dfilm"] = _rel(dft[n, :, OFF: OFF + 2 * Co].float().cpu(), demb)
                assert float(dft[n, :, :OFF].abs().max()) == 0 and float(dft[n, :, OFF + 2 * Co:].abs().max()) == 0
                do = do * scale
            draw, dgam, dbet = ob.gn_mish_bwd(raw, gam[n], bet[n], do)
            dxr, dwr, dbr = ob.conv1d_bwd(xn, w, draw, padding=2)
            errs["raw"] = _rel(out["raw"][n].float().cpu().permute(0, 2, 1), raw)
            errs["draw"] = _rel(out["draw"][n].float().cpu().permute(0, 2, 1), draw)
            errs["dgamma"] = _rel(out["dgamma"][n].float().cpu(), dgam)
            errs["dbeta"] = _rel(out["dbeta"][n].float().cpu(), dbet)
            errs["dbias"] = _rel(out["dbias"][n].float().cpu(), dbr)
            errs["dw"] = _rel(ub.unpack_wgrad(out["dw"], Ci, 5)[n].float().cpu(), dwr)
            errs["dx"] = _rel(dx[n].float().cpu().permute(0, 2, 1), dxr)
            bad = {k: v for k, v in errs.items() if not v <= tol}
            assert not bad, (n, bad, errs)
        return errs
    return plan, check


def colsum_case(device, G=2, rows=100, C=70, ld=96):
    from vla_touch_b200 import native as nv
    from vla_touch_b200.plan import ptr
    plan = Plan(device)
    x = plan.buf("x", (G, rows, ld), torch.float32)
    x.copy_(torch.randn(G, rows, ld, generator=torch.Generator().manual_seed(3)))
    out = plan.buf("out", (G, 128), torch.float32)
    d = nv.ColsumDesc()
    d.x, d.ld, d.x_g, d.G, d.rows, d.C, d.out, d.out_ld = ptr(x), ld, rows * ld, G, rows, C, ptr(out), 128
    plan.add(d, "colsum")

    def check(tol=1e-5):
        ref = x[:, :, :C].float().cpu().sum(dim=1)
        assert _rel(out[:, :C].float().cpu(), ref) <= tol
        assert float(out[:, C:].abs().max()) == 0
    return plan, check


def dgrad_case(kind: str, device, G=2, B=3, C=256, seed=9):
    """conv_dgrad / downsample_dgrad / upsample_dgrad (forward implicit-GEMM descriptors) against the oracle's dX."""
    g = torch.Generator().manual_seed(seed)
    T_dy, T_dx, K = {"k5": (16, 16, 5), "down": (8, 16, 3), "up": (32, 16, 4)}[kind]
    ws = [torch.randn(C, C, K, generator=g) / (K * C) ** 0.5 for _ in range(G)]
    plan = Plan(device)
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

    def check(tol=1e-2):
        for n in range(G):
            dyn = dy[n].float().cpu().permute(0, 2, 1)
            w = bf(ws[n])
            z = torch.zeros(B, C, T_dx)
            if kind == "k5":
                ref = ob.conv1d_bwd(z, w, dyn, padding=2)[0]
            elif kind == "down":
                ref = ob.conv1d_bwd(z, w, dyn, stride=2, padding=1)[0]
            else:
                ref = ob.convT1d_bwd(z, w, dyn)[0]
            e = _rel(dx[n].float().cpu().permute(0, 2, 1), ref)
            assert e <= tol, (kind, n, e)
    return plan, check


def res_block_case(device, Ci: int, Co: int, G=2, B=3, T=16, seed=21, need_dx=True):
    """res_block_backward against _res_block_bwd of the oracle.  Ci != Co -> the block has a 1x1 residual_conv.
    Ci == 7: the network's first block (64-channel padded input shared by all nets, no input gradient)."""
    import torch.nn.functional as F
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.randn(*s, generator=g)
    p = "blk."
    first = Ci == 7
    sds = []
    for _ in range(G):
        sd = {p + "blocks.0.block.0.weight": bf(r(Co, Ci, 5) / (5 * Ci) ** 0.5), p + "blocks.0.block.0.bias": 0.1 * r(Co),
              p + "blocks.0.block.1.weight": 1 + 0.2 * r(Co), p + "blocks.0.block.1.bias": 0.2 * r(Co),
              p + "blocks.1.block.0.weight": bf(r(Co, Co, 5) / (5 * Co) ** 0.5), p + "blocks.1.block.0.bias": 0.1 * r(Co),
              p + "blocks.1.block.1.weight": 1 + 0.2 * r(Co), p + "blocks.1.block.1.bias": 0.2 * r(Co),
              p + "cond_encoder.1.weight": r(2 * Co, 512) / 512 ** 0.5, p + "cond_encoder.1.bias": 1 + 0.1 * r(2 * Co)}
        if Ci != Co:
            sd[p + "residual_conv.weight"] = bf(r(Co, Ci, 1) / Ci ** 0.5)
            sd[p + "residual_conv.bias"] = 0.1 * r(Co)
        sds.append(sd)
    mgf = r(B, 512)
    xs = bf(r(B, Ci, T)) if first else bf(r(G, B, Ci, T))                       # oracle layout [B, C, T]
    douts = r(G, B, Co, T)
    caches, embs = [], []
    for n in range(G):
        cache = {}
        ob._res_block_fwd(sds[n], p, xs if first else xs[n], mgf, cache)
        caches.append(cache)
        embs.append(F.linear(mgf, sds[n][p + "cond_encoder.1.weight"], sds[n][p + "cond_encoder.1.bias"]))
    plan = Plan(device)
    ctx = ub.DgradCtx(G, precise=False)
    Cx = 64 if first else Ci
    if first:
        x = plan.buf("x", (B, T, Cx), torch.bfloat16)
        x[:, :, :Ci] = xs.permute(0, 2, 1).to(device)
    else:
        x = plan.buf("x", (G, B, T, Cx), torch.bfloat16)
        x.copy_(xs.permute(0, 1, 3, 2))
    y1 = plan.buf("y1", (G, B, T, Co), torch.bfloat16)
    y1.copy_(torch.stack([c[p + "blocks.1."][0] for c in caches]).permute(0, 1, 3, 2))
    dout = plan.buf("dout", (G, B, T, Co), torch.float32)
    dout.copy_(douts.permute(0, 1, 3, 2))
    ft = plan.buf("film", (G, B, 2 * Co), torch.float32)
    ft.copy_(torch.stack(embs))
    dft = plan.buf("dfilm", (G, B, 2 * Co), torch.float32)
    dx = None if first or not need_dx else plan.buf("dx", (G, B, T, Cx), torch.float32)
    out = ub.res_block_backward(plan, ctx, B, sds, p, _View(x, T, Cx, shared=first), _View(y1, T, Co), dout,
                                None if dx is None else _View(dx, T, Cx), (ft, dft, 0))

    def check(tol=2e-2):
        errs = {}
        for n in range(G):
            grads = {}
            dxr, dmgf = ob._res_block_bwd(sds[n], p, douts[n], mgf, caches[n], grads)
            for k, v in out.items():
                ref = grads[p + k]
                if isinstance(v, tuple):
                    got = ub.unpack_wgrad(v[0], ref.shape[1], v[1])[n].float().cpu()
                else:
                    got = v[n].float().cpu()
                errs[k] = max(errs.get(k, 0.0), _rel(got, ref))
            demb = dft[n].float().cpu()
            errs["cond_encoder.1.bias"] = max(errs.get("cond_encoder.1.bias", 0.0), _rel(demb.sum(0), grads[p + "cond_encoder.1.bias"]))
            errs["cond_encoder.1.weight"] = max(errs.get("cond_encoder.1.weight", 0.0),
                                                _rel(demb.t() @ mgf, grads[p + "cond_encoder.1.weight"]))
            if dx is not None:
                errs["dx"] = max(errs.get("dx", 0.0), _rel(dx[n].float().cpu().permute(0, 2, 1), dxr))
        bad = {k: v for k, v in errs.items() if not v <= tol}
        assert not bad, (bad, errs)
        return errs
    return plan, check


def unet_case(device, A=7, T=16, B=2, G=2, seed=31):
    """Training forward + build_unet_backward of G whole U-Nets against unet_forward_cached / unet_backward of the oracle:
    every parameter gradient the plan produces, per net."""
    import vt_testutil as U
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200 import unet_train as ut
    from vla_touch_b200.unet import FILM_ROWS, UnetWeights, xpad_desc
    names = ["b_net", "v_net", "s_net"][:G]
    sds = [{k: (bf(v) if v.dim() >= 2 else v.clone()) for k, v in U.net_sd(A, seed, n).items()} for n in names]
    x = syn.det_uniform("bwd.x", (B, T, A), seed, -1.0, 1.0)
    cond = syn.det_normal("bwd.cond", (B, 256), seed)
    t = torch.linspace(0.2, 0.9, B)
    dout = syn.det_normal("bwd.dout", (G, B, T, A), seed)
    plan = Plan(device)
    W = UnetWeights(sds, A, device, precise=False)
    W.register(plan)
    xb = plan.buf("in.x", (B, T, A), torch.float32); xb.copy_(x)
    tbuf = plan.buf("in.t", (B,), torch.float32); tbuf.copy_(t)
    cb = plan.buf("in.cond", (B, 256), torch.float32); cb.copy_(cond)
    dvs = plan.buf("in.dvs", (G, B, T, A), torch.float32); dvs.copy_(dout)
    film = plan.buf("film", (G, B, FILM_ROWS), torch.float32)
    dfilm = plan.buf("dfilm", (G, B, FILM_ROWS), torch.float32)
    tb = ut.UnetTrainBuffers(plan, W, B, T)
    plan.add(xpad_desc(W, xb, B * T, tb), "xpad")
    tf = ut.build_time_film_train(plan, W, tbuf, B, cb, film)
    ut.build_unet_train_forward(plan, W, tb, film)
    grads = ut.build_unet_backward(plan, W, sds, tb, dvs, film, dfilm)
    extra = ut.build_film_time_backward(plan, W, sds, tf, B, film, dfilm, grads)

    def check(tol=4e-2):
        errs = {}
        for n in range(G):
            out_ref, cache = ob.unet_forward_cached(sds[n], x, t, cond)
            ref, dcond, _ = ob.unet_backward(sds[n], cache, dout[n])
            errs["forward"] = max(errs.get("forward", 0.0), _rel(tb.out[n].float().cpu(), out_ref))
            for k, v in grads.items():
                got = ut.grad_tensor(v, ref[k].shape)[n].float().cpu()
                errs[k] = max(errs.get(k, 0.0), _rel(got.reshape(ref[k].shape), ref[k]))
            errs["d_cond"] = max(errs.get("d_cond", 0.0), _rel(extra["dcond"][n].float().cpu(), dcond))
        missing = set(ref) - set(grads)
        bad = {k: v for k, v in errs.items() if not v <= tol}
        assert not bad, (bad,)
        return dict(worst=max(errs.values()), n=len(errs), missing=sorted(missing))
    return plan, check


def loss_case(device, A=7, T=64):
    """LossBackwardProgram against the REFERENCE's own get_loss(...).backward() digests (tests/golden/loss_grads_*.npz, made by
    oracle/gen_golden_grads.py from the unmodified reference): loss values, d loss / d obs_cond in full, and norm / sum / first
    elements of all 438 parameter gradients.  bf16 operands: 3e-2 of each tensor's scale."""
    import vt_testutil as U
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.unet_train import LossBackwardProgram
    g, gg = U.golden(f"loss_A{A}_T{T}"), U.golden(f"loss_grads_A{A}_T{T}")
    full = U.net_sd(A, 21)
    sub = lambda n: {k[len(n):]: v for k, v in full.items() if k.startswith(n)}
    lp = LossBackwardProgram([sub("b_net."), sub("v_net."), sub("s_net.")], A, 3, T, 0.03, device)
    lp.set_inputs(syn.det_uniform("loss.vla", (3, T, A), 24, -1.0, 1.0), syn.det_uniform("loss.exp", (3, T, A), 24, -1.0, 1.0),
                  syn.det_normal("loss.cond", (3, 256), 24), torch.as_tensor(g["step"]), torch.as_tensor(g["z_unit"]))

    def check(tol=3e-2):
        """(1) reference digests: loss values, d_cond in full, norm and leading elements of all 438 gradients (the `sum` digest
        is not used at bf16: for gradients with a coherent factor, e.g. dW = d_emb^T Mish(gf) with mean(Mish) > 0, it amplifies
        the operand rounding by sqrt(numel)); (2) every gradient tensor IN FULL against autograd through the forward oracle,
        which tests/test_oracle_golden.py::test_loss_gradients holds to all three reference digests at 1e-4."""
        from oracle import vt_oracle as orc
        for i, key in enumerate(("loss", "v_loss", "s_loss", "b_loss")):
            assert abs(float(lp.out[i]) - float(g[key])) <= 2e-2 * max(1.0, abs(float(g[key]))), (key, float(lp.out[i]), float(g[key]))
        dc = torch.as_tensor(gg["d_cond"])
        worst = {"d_cond": _rel(lp.d_cond.float().cpu(), dc)}
        grads = lp.grads
        names = [str(n) for n in gg["names"]]
        assert sorted(names) == sorted(grads), "every reference parameter has a gradient"
        sd = {k: v.clone().requires_grad_(True) for k, v in full.items()}
        loss, *_ = orc.bridge_losses(sd, syn.det_normal("loss.cond", (3, 256), 24), syn.det_uniform("loss.exp", (3, T, A), 24, -1.0, 1.0),
                                     syn.det_uniform("loss.vla", (3, T, A), 24, -1.0, 1.0), torch.as_tensor(g["step"]), torch.as_tensor(g["z_unit"]))
        loss.backward()
        for i, n in enumerate(names):
            got = grads[n].float().cpu()
            gr = got.flatten().double()
            scale = max(float(gg["norm"][i]), 1e-12)
            k = min(8, gr.numel())
            worst[n] = max(abs(float(gr.norm()) - float(gg["norm"][i])) / scale,
                           float((gr[:k] - torch.as_tensor(gg["head"][i][:k]).double()).abs().max()) / scale,
                           _rel(got, sd[n].grad))
        bad = {k: v for k, v in worst.items() if not v <= tol}
        assert not bad, bad
        return dict(worst=max(worst.values()), tensors=len(worst))
    return lp.plan, check


def interpolant(A, T, device, precise=False, beta=0.03):
    """vla_touch_b200's StochasticInterpolants with the fixture weights (the model the reference gradient digests were made on)."""
    import vt_testutil as U
    from vla_touch_b200.bridge.bridge_model import StochasticInterpolants
    args = {'interpolant_type': 'linear', 'gamma_type': '2^0.5*t(t-1)', 'epsilon_type': '1-t', 'prior_policy': 'vla',
            'beta_max': beta, 'sde_type': 'vs', 'action_dim': A, 'obs_dim': 256, 'obs_horizon': 1, 'net_type': 'unet1D_si',
            'pretrain': False, 'context_frames': 2, 'horizon': T}
    si = StochasticInterpolants(precise=precise)
    si.load_model(args, device)
    si.net.load_state_dict(U.net_sd(A, 21))
    return si


def training_step_case(device, A=10, T=16):
    """The reference training step's autograd contract (bridge_train.py:315-334): loss, info = get_loss(batch);
    loss.backward() fills .grad of every net parameter and back-propagates into obs_cond; after an in-place parameter update the
    next call uses the new weights (operand copies are re-packed).  Checked against autograd through the forward oracle."""
    import vt_testutil as U
    from oracle import vt_oracle as orc
    from vla_touch_b200 import synthetic as syn
    g = U.golden(f"loss_A{A}_T{T}")
    si = interpolant(A, T, device)
    si.step_override, si.z_override = torch.as_tensor(g["step"]).to(device), torch.as_tensor(g["z_unit"]).to(device)
    cond0 = syn.det_normal("loss.cond", (3, 256), 24)
    exp, vla = syn.det_uniform("loss.exp", (3, T, A), 24, -1.0, 1.0), syn.det_uniform("loss.vla", (3, T, A), 24, -1.0, 1.0)

    def one_step(tol=3e-2):
        pre = torch.nn.Linear(256, 256, bias=False).to(device)          # stands for the state encoder in front of obs_cond
        with torch.no_grad():
            pre.weight.copy_(torch.eye(256))
        for p in si.net.parameters():
            p.grad = None
        loss, info = si.get_loss({"obs_cond": pre(cond0.to(device)), "expert_act": exp.to(device), "vla_act": vla.to(device)}, device)
        assert loss.requires_grad and not info["v_loss"].requires_grad
        (2.0 * loss).backward()                                          # a scaled loss: gradients scale with it
        sd = {k: v.detach().cpu().clone().requires_grad_(True) for k, v in si.net.state_dict().items()}
        c = cond0.clone().requires_grad_(True)
        ref, *_ = orc.bridge_losses(sd, c, exp, vla, torch.as_tensor(g["step"]), torch.as_tensor(g["z_unit"]))
        (2.0 * ref).backward()
        assert abs(float(loss.detach()) - float(ref.detach())) <= 2e-2 * max(1.0, abs(float(ref.detach())))
        worst = {"d_cond(through the producer of obs_cond)": _rel(pre.weight.grad.float().cpu(), c.grad.t() @ cond0)}
        for n, p in si.net.named_parameters():
            assert p.grad is not None and p.grad.shape == p.shape, n
            worst[n] = _rel(p.grad.float().cpu(), sd[n].grad)
        bad = {k: v for k, v in worst.items() if not v <= tol}
        assert not bad, bad
        return max(worst.values())

    def check():
        w1 = one_step()
        with torch.no_grad():                                            # an "optimizer step": in-place update of every parameter
            for i, p in enumerate(si.net.parameters()):
                p.mul_(1.0 + 0.05 * ((i % 3) - 1)).add_(0.01 if p.dim() == 1 else 0.0)
        w2 = one_step()
        return dict(first=w1, after_update=w2)
    return check


def batch_additivity_case(device, B=256, T=64, A=7, seed=41):
    """Size-independent property at the BASELINE batch (no oracle needed): the loss is a batch MEAN of per-sample terms and no
    statistic crosses samples (GroupNorm is per sample), so the gradients of the full batch equal the mean of the gradients of
    its two halves, and so do the loss values."""
    from vla_touch_b200 import shapes as shp
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.params import sub_state_dict
    from vla_touch_b200.unet_train import LossBackwardProgram
    full = syn.synth_state_dict(shp.si_net_shapes(A, 256), seed, prefix="net.")
    sds = [sub_state_dict(full, p) for p in ("b_net.", "v_net.", "s_net.")]
    g = torch.Generator().manual_seed(seed)
    x0, x1 = torch.rand(B, T, A, generator=g) * 2 - 1, torch.rand(B, T, A, generator=g) * 2 - 1
    cond, step, z = torch.randn(B, 256, generator=g), torch.rand(B, generator=g), torch.randn(B, T, A, generator=g)
    h = B // 2

    def run(lo, hi):
        lp = LossBackwardProgram(sds, A, hi - lo, T, 0.03, device)
        lp.set_inputs(x0[lo:hi], x1[lo:hi], cond[lo:hi], step[lo:hi], z[lo:hi])
        out = lp.run().clone()
        return out, {k: v.clone() for k, v in lp.grads.items()}, lp.d_cond.clone()

    def check(tol=3e-2):
        o, gr, dc = run(0, B)
        oa, ga, dca = run(0, h)
        ob_, gb, dcb = run(h, B)
        assert torch.isfinite(o).all()
        assert float((o - 0.5 * (oa + ob_)).abs().max()) <= 1e-3 * float(o.abs().max())
        worst = {"d_cond": _rel(dc.float().cpu(), 0.5 * torch.cat((dca, dcb)).float().cpu())}
        for k in gr:
            assert torch.isfinite(gr[k]).all(), k
            worst[k] = _rel(gr[k].float().cpu(), 0.5 * (ga[k] + gb[k]).float().cpu())
        bad = {k: v for k, v in worst.items() if not v <= tol}
        assert not bad, bad
        return dict(worst=max(worst.values()), tensors=len(worst))
    return check


def lstm_layers_case(device, B=5, T=12, k_in=138, layers=2, seed=51):
    """lstm_train.lstm_layers_train (training forward + BPTT of the stacked nn.LSTM layers) against torch.nn.LSTM autograd --
    the module the reference itself trains (lstm_step_controller.py:66-73)."""
    from vla_touch_b200 import lstm_train as lt
    from vla_touch_b200.plan import round_up
    g = torch.Generator().manual_seed(seed)
    H = 256
    ref = torch.nn.LSTM(input_size=k_in, hidden_size=H, num_layers=layers, batch_first=True)
    with torch.no_grad():
        for p_ in ref.parameters():
            p_.copy_(torch.randn(p_.shape, generator=g) * (0.06 if p_.dim() == 2 else 0.1))
    sd = {k: v.detach().clone() for k, v in ref.state_dict().items()}
    x = bf(torch.randn(B, T, k_in, generator=g))
    dy = torch.randn(B, T, H, generator=g)
    plan = Plan(device)
    kp = round_up(k_in, 64)
    xb = plan.buf("x", (B, T, kp), torch.bfloat16)
    xb[:, :, :k_in] = x.to(device)
    dyb = plan.buf("dy", (B, T, H), torch.float32)
    dyb.copy_(dy)
    lays = lt.lstm_layers_train(plan, xb, k_in, sd, B, T, layers)
    d = dyb
    for lay in reversed(lays):
        d = lay.backward(d)
        if lay is not lays[0]:
            d = d[:, :, :H] if d.shape[-1] != H else d

    def check(tol=3e-2):
        xr = x.clone().requires_grad_(True)
        y, _ = ref(xr)
        (y * dy).sum().backward()
        errs = {"y": _rel(lays[-1].y.float().cpu(), y.detach()), "dx": _rel(lays[0].dx[:, :, :k_in].float().cpu(), xr.grad)}
        for l, lay in enumerate(lays):
            for k, v in lay.grads.items():
                errs[f"{k}_l{l}"] = _rel(v.float().cpu(), getattr(ref, f"{k}_l{l}").grad)
        bad = {k: v for k, v in errs.items() if not v <= tol}
        assert not bad, (bad, errs)
        return errs
    return plan, check


def lstm_fixture_inputs(A, Fd, T):
    """The inputs oracle/gen_golden_grads.py::lstm_grads fed to the reference."""
    from oracle import vt_oracle as orc
    from vla_touch_b200 import synthetic as syn
    st = syn.synth_stats_varied(A, 41)
    return dict(vla_n=orc.normalize_actions(syn.det_uniform("lstm.vla", (3, T, A), 41, -1.0, 1.0), st, "vla"),
                forces=syn.det_normal("lstm.forces", (3, T, Fd), 41), cond=syn.det_normal("lstm.cond", (3, 256), 41),
                expert=syn.det_uniform("lstm.exp", (3, T, A), 41, -1.0, 1.0))


def lstm_loss_case(device, A=7, Fd=64, T=32):
    """lstm_train.LstmLossBackwardProgram against the REFERENCE's own TactileLSTMController.get_loss(...).backward() digests
    (tests/golden/lstm_grads_*.npz, oracle/gen_golden_grads.py) and, tensor by tensor in full, against the explicit BPTT oracle
    `lstm_loss_backward` (which tests/test_oracle_golden.py holds to those digests at 1e-4)."""
    import vt_testutil as U
    from vla_touch_b200 import shapes as shp
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.lstm_train import LstmLossBackwardProgram
    gg = U.golden(f"lstm_grads_A{A}_F{Fd}_T{T}")
    mods = {"force_encoder": syn.synth_state_dict(shp.mlp_shapes([Fd, 128, 128]), 41, "lstm.force_encoder."),
            "lstm": syn.synth_state_dict(shp.lstm_shapes(128 + A), 41, "lstm.lstm."),
            "output_head": syn.synth_state_dict(shp.lstm_head_shapes(256, A), 41, "lstm.output_head.")}
    inputs = lstm_fixture_inputs(A, Fd, T)
    B = inputs["vla_n"].shape[0]
    lp = LstmLossBackwardProgram(mods, A, Fd, B, T, device)
    lp.set_inputs(inputs["vla_n"], inputs["forces"], inputs["cond"], inputs["expert"])

    def check(tol=3e-2):
        loss_ref, ref, dcond = ob.lstm_loss_backward(mods, inputs["vla_n"], inputs["cond"], inputs["forces"], inputs["expert"])
        assert abs(lp.loss() - float(loss_ref)) <= 2e-2 * abs(float(loss_ref)), (lp.loss(), float(loss_ref))
        assert abs(float(loss_ref) - float(gg["loss"])) <= 1e-4 * abs(float(gg["loss"]))
        worst = {"d_cond": _rel(lp.d_cond.float().cpu(), dcond), "d_cond(reference)": _rel(lp.d_cond.float().cpu(), torch.as_tensor(gg["d_cond"]))}
        assert sorted(lp.grads) == sorted(ref), sorted(set(lp.grads) ^ set(ref))
        for k, v in lp.grads.items():
            worst[k] = _rel(v.float().cpu().reshape(ref[k].shape), ref[k])
        names = [str(n) for n in gg["names"]]
        for i, n in enumerate(names):
            gr = lp.grads[n].float().cpu().flatten().double()
            scale = max(float(gg["norm"][i]), 1e-12)
            worst[n + "(reference norm)"] = abs(float(gr.norm()) - scale) / scale
        bad = {k: v for k, v in worst.items() if not v <= tol}
        assert not bad, bad
        return dict(worst=max(worst.values()), tensors=len(lp.grads))
    return lp.plan, check


def lstm_dropout_case(device, A=7, Fd=64, T=32, p_drop=0.1):
    """LstmLossBackwardProgram(dropout=0.1) with injected uniforms against torch autograd of the same network with the same two
    inverted-dropout masks (between the LSTM layers, nn.LSTM(dropout=p) semantics, and in the head, nn.Dropout(p))."""
    import torch.nn.functional as F
    from vla_touch_b200 import shapes as shp
    from vla_touch_b200 import synthetic as syn
    from vla_touch_b200.lstm_train import LstmLossBackwardProgram
    H = 256
    mods = {"force_encoder": syn.synth_state_dict(shp.mlp_shapes([Fd, 128, 128]), 41, "lstm.force_encoder."),
            "lstm": syn.synth_state_dict(shp.lstm_shapes(128 + A), 41, "lstm.lstm."),
            "output_head": syn.synth_state_dict(shp.lstm_head_shapes(256, A), 41, "lstm.output_head.")}
    inp = lstm_fixture_inputs(A, Fd, T)
    B = inp["vla_n"].shape[0]
    g = torch.Generator().manual_seed(77)
    u0, u1 = torch.rand(B * T, H, generator=g), torch.rand(B * T, H, generator=g)
    lp = LstmLossBackwardProgram(mods, A, Fd, B, T, device, dropout=p_drop, inject_uniforms=True)
    lp.set_inputs(inp["vla_n"], inp["forces"], inp["cond"], inp["expert"])
    lp.u[0].copy_(u0); lp.u[1].copy_(u1)

    def check(tol=3e-2):
        P = {m: {k: v.clone().requires_grad_(True) for k, v in sd.items()} for m, sd in mods.items()}
        fe, ls, hd = P["force_encoder"], P["lstm"], P["output_head"]
        cond = inp["cond"].clone().requires_grad_(True)
        m0 = ((u0 >= p_drop).float() / (1 - p_drop)).reshape(B, T, H)
        m1 = ((u1 >= p_drop).float() / (1 - p_drop)).reshape(B, T, H)
        fenc = F.linear(F.gelu(F.linear(inp["forces"], fe["0.weight"], fe["0.bias"])), fe["2.weight"], fe["2.bias"])
        x = torch.cat([fenc, inp["vla_n"]], dim=-1)

        def layer(xin, l):
            h, c, ys = torch.zeros(B, H), torch.zeros(B, H), []
            for t in range(T):
                gt = F.linear(xin[:, t], ls[f"weight_ih_l{l}"], ls[f"bias_ih_l{l}"]) + F.linear(h, ls[f"weight_hh_l{l}"], ls[f"bias_hh_l{l}"])
                i_, f_, g_, o_ = gt.chunk(4, dim=-1)
                c = torch.sigmoid(f_) * c + torch.sigmoid(i_) * torch.tanh(g_)
                h = torch.sigmoid(o_) * torch.tanh(c)
                ys.append(h)
            return torch.stack(ys, dim=1)
        y1 = layer(layer(x, 0) * m0, 1)
        comb = torch.cat([y1, cond[:, None].expand(B, T, H)], dim=-1)
        zn = F.gelu(F.layer_norm(F.linear(comb, hd["0.weight"], hd["0.bias"]), (H,), hd["1.weight"], hd["1.bias"], 1e-5)) * m1
        out = inp["vla_n"] + F.linear(zn, hd["4.weight"], hd["4.bias"])
        loss = F.mse_loss(out, inp["expert"])
        loss.backward()
        assert abs(lp.loss() - float(loss.detach())) <= 2e-2 * abs(float(loss.detach()))
        keep = float((lp.masks[0] > 0).float().mean())
        assert abs(keep - (1 - p_drop)) < 0.02 and abs(float(lp.masks[0].max()) - 1 / (1 - p_drop)) < 1e-6
        worst = {"d_cond": _rel(lp.d_cond.float().cpu(), cond.grad)}
        for k, v in lp.grads.items():
            m, key = k.split(".", 1)
            worst[k] = _rel(v.float().cpu().reshape(P[m][key].shape), P[m][key].grad)
        bad = {k: v for k, v in worst.items() if not v <= tol}
        assert not bad, bad
        return dict(worst=max(worst.values()), tensors=len(lp.grads))
    return lp.plan, check
