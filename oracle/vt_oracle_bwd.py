"""TEST INFRASTRUCTURE ONLY (see oracle/vt_oracle.py) -- explicit BACKWARD of the interpolant U-Net and of the three bridge
losses, written op by op WITHOUT autograd, in the decomposition the round-2 kernels will use:

    conv / conv-transpose   dgrad = per-tap implicit GEMM with the transposed weight slice, wgrad = per-tap X_shift^T dY,
                            bias grad = column sum                       (conditional_unet_1D.py:22-55)
    GroupNorm(8) + Mish     one fused elementwise + per-(sample, group) reduction pass from the saved raw conv output
    FiLM                    d scale = sum_t dOut * y, d shift = sum_t dOut (per sample, channel), then the cond-encoder Linear
    Linear / Mish / time embedding MLP, residual fan-out, skip concatenation (:86-105, :194-247)
    losses                  d v_loss / dv = (v - (x1 - x0)) / B, ... (bridge_model.py:183-218)

`unet_forward_cached` restates oracle.vt_oracle.unet_forward and keeps what a training forward has to save; `unet_backward`
returns the gradients of every parameter, of the global condition and of the input.  tests/test_oracle_golden.py holds it to
(a) autograd through the forward oracle on small random cases and (b) the reference's own loss.backward() digests
(oracle/gen_golden_grads.py).
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn.functional as F

from oracle.vt_oracle import T_MIN, sinusoidal_pos_emb

SD = Dict[str, torch.Tensor]


# ---------------------------------------------------------------------------------------------- primitives
def mish_grad(x: torch.Tensor) -> torch.Tensor:
    """d mish / dx = tanh(sp) + x * (1 - tanh(sp)^2) * sigmoid(x), sp = softplus(x)"""
    tsp = torch.tanh(F.softplus(x))
    return tsp + x * (1 - tsp * tsp) * torch.sigmoid(x)


def conv1d_fwd(x, w, b, stride=1, padding=0):
    """y[n,co,t] = b[co] + sum_{ci,k} w[co,ci,k] x[n,ci,t*stride + k - padding]   (one GEMM per tap k)"""
    N, Ci, T = x.shape
    Co, _, K = w.shape
    To = (T + 2 * padding - K) // stride + 1
    xp = F.pad(x, (padding, padding))
    y = b[None, :, None].expand(N, Co, To).clone()
    for k in range(K):
        xs = xp[:, :, k: k + (To - 1) * stride + 1: stride]          # [N,Ci,To]: the tap's shifted view
        y = y + torch.einsum("oc,nct->not", w[:, :, k], xs)
    return y


def conv1d_bwd(x, w, dy, stride=1, padding=0):
    """-> dx, dw, db.  wgrad: dw[:, :, k] = sum_{n,t} dy[n,:,t] (x) x_shift_k[n,:,t];  dgrad: dx scatter of w[:, :, k]^T dy."""
    N, Ci, T = x.shape
    Co, _, K = w.shape
    To = dy.shape[-1]
    xp = F.pad(x, (padding, padding))
    dxp = torch.zeros_like(xp)
    dw = torch.zeros_like(w)
    for k in range(K):
        sl = slice(k, k + (To - 1) * stride + 1, stride)
        dw[:, :, k] = torch.einsum("not,nct->oc", dy, xp[:, :, sl])
        dxp[:, :, sl] += torch.einsum("oc,not->nct", w[:, :, k], dy)
    dx = dxp[:, :, padding: padding + T]
    return dx, dw, dy.sum(dim=(0, 2))


def convT1d_fwd(x, w, b):
    """ConvTranspose1d(k4, s2, p1): y[n,co,2t + k - 1] += sum_ci w[ci,co,k] x[n,ci,t]  (two 2-tap GEMMs, one per output phase)"""
    N, Ci, T = x.shape
    _, Co, K = w.shape
    yp = torch.zeros(N, Co, 2 * T + 2, dtype=x.dtype)                                 # index u + 1, u = 2t + k - 1 in [-1, 2T]
    for k in range(K):
        yp[:, :, k: k + 2 * T: 2] += torch.einsum("co,nct->not", w[:, :, k], x)
    return yp[:, :, 1: 2 * T + 1] + b[None, :, None]


def convT1d_bwd(x, w, dy):
    N, Ci, T = x.shape
    dyp = F.pad(dy, (1, 1))
    dx = torch.zeros_like(x)
    dw = torch.zeros_like(w)
    for k in range(w.shape[-1]):
        sl = dyp[:, :, k: k + 2 * T: 2]                                # dy[n,co,2t + k - 1]
        dx += torch.einsum("co,not->nct", w[:, :, k], sl)
        dw[:, :, k] = torch.einsum("nct,not->co", x, sl)
    return dx, dw, dy.sum(dim=(0, 2))


def gn_mish_fwd(raw, gamma, beta, groups=8, eps=1e-5):
    N, C, T = raw.shape
    r = raw.reshape(N, groups, -1)
    mean = r.mean(dim=-1, keepdim=True)
    rstd = torch.rsqrt(r.var(dim=-1, unbiased=False, keepdim=True) + eps)
    xhat = ((r - mean) * rstd).reshape(N, C, T)
    return F.mish(xhat * gamma[None, :, None] + beta[None, :, None])


def gn_mish_bwd(raw, gamma, beta, dout, groups=8, eps=1e-5):
    """From the saved raw conv output: recompute x_hat and the GroupNorm output, then
       dy = dout * mish'(y);  d gamma = sum dy x_hat;  d beta = sum dy;
       d raw = rstd * (dxh - mean_g(dxh) - x_hat * mean_g(dxh * x_hat)),  dxh = dy * gamma."""
    N, C, T = raw.shape
    r = raw.reshape(N, groups, -1)
    mean = r.mean(dim=-1, keepdim=True)
    rstd = torch.rsqrt(r.var(dim=-1, unbiased=False, keepdim=True) + eps)
    xhat = ((r - mean) * rstd).reshape(N, C, T)
    y = xhat * gamma[None, :, None] + beta[None, :, None]
    dy = dout * mish_grad(y)
    dgamma = (dy * xhat).sum(dim=(0, 2))
    dbeta = dy.sum(dim=(0, 2))
    dxh = (dy * gamma[None, :, None]).reshape(N, groups, -1)
    xh = xhat.reshape(N, groups, -1)
    draw = rstd * (dxh - dxh.mean(dim=-1, keepdim=True) - xh * (dxh * xh).mean(dim=-1, keepdim=True))
    return draw.reshape(N, C, T), dgamma, dbeta


def linear_bwd(x, w, dy):
    return dy @ w, dy.t() @ x, dy.sum(dim=0)


# ---------------------------------------------------------------------------------------------- blocks
def _conv_block_fwd(sd, p, x, cache):
    w, b = sd[p + "block.0.weight"], sd[p + "block.0.bias"]
    raw = conv1d_fwd(x, w, b, padding=w.shape[-1] // 2)
    cache[p] = (x, raw)                                                # what the training forward stores
    return gn_mish_fwd(raw, sd[p + "block.1.weight"], sd[p + "block.1.bias"])


def _conv_block_bwd(sd, p, dout, cache, grads):
    x, raw = cache[p]
    w = sd[p + "block.0.weight"]
    draw, grads[p + "block.1.weight"], grads[p + "block.1.bias"] = gn_mish_bwd(raw, sd[p + "block.1.weight"],
                                                                               sd[p + "block.1.bias"], dout)
    dx, grads[p + "block.0.weight"], grads[p + "block.0.bias"] = conv1d_bwd(x, w, draw, padding=w.shape[-1] // 2)
    return dx


def _res_block_fwd(sd, p, x, mgf, cache):
    y0 = _conv_block_fwd(sd, p + "blocks.0.", x, cache)
    emb = F.linear(mgf, sd[p + "cond_encoder.1.weight"], sd[p + "cond_encoder.1.bias"])
    C = y0.shape[1]
    scale, shift = emb[:, :C, None], emb[:, C:, None]
    y1 = scale * y0 + shift
    y2 = _conv_block_fwd(sd, p + "blocks.1.", y1, cache)
    cache[p + "film"] = (y0, scale)
    if p + "residual_conv.weight" in sd:
        res = conv1d_fwd(x, sd[p + "residual_conv.weight"], sd[p + "residual_conv.bias"])
    else:
        res = x
    cache[p + "x"] = x
    return y2 + res


def _res_block_bwd(sd, p, dout, mgf, cache, grads):
    """-> (dx, d mish(gf)) ; parameter gradients go to `grads`"""
    y0, scale = cache[p + "film"]
    x = cache[p + "x"]
    dy1 = _conv_block_bwd(sd, p + "blocks.1.", dout, cache, grads)
    demb = torch.cat([(dy1 * y0).sum(dim=-1), dy1.sum(dim=-1)], dim=1)                     # [B, 2C]: d scale | d shift
    dmgf, grads[p + "cond_encoder.1.weight"], grads[p + "cond_encoder.1.bias"] = linear_bwd(mgf, sd[p + "cond_encoder.1.weight"], demb)
    dx = _conv_block_bwd(sd, p + "blocks.0.", dy1 * scale, cache, grads)
    if p + "residual_conv.weight" in sd:
        dxr, grads[p + "residual_conv.weight"], grads[p + "residual_conv.bias"] = conv1d_bwd(x, sd[p + "residual_conv.weight"], dout)
        dx = dx + dxr
    else:
        dx = dx + dout
    return dx, dmgf


def unet_forward_cached(sd: SD, sample, timestep, global_cond) -> Tuple[torch.Tensor, dict]:
    cache: dict = {}
    x = sample.moveaxis(-1, -2)
    t = timestep.expand(sample.shape[0])
    pe = sinusoidal_pos_emb(t, sd["diffusion_step_encoder.1.weight"].shape[1])
    t1 = F.linear(pe, sd["diffusion_step_encoder.1.weight"], sd["diffusion_step_encoder.1.bias"])
    temb = F.linear(F.mish(t1), sd["diffusion_step_encoder.3.weight"], sd["diffusion_step_encoder.3.bias"])
    gf = torch.cat([temb, global_cond], dim=-1)
    mgf = F.mish(gf)
    cache["time"] = (pe, t1, gf, mgf)
    skips: List[torch.Tensor] = []
    order: List[tuple] = []
    L = 0
    while f"down_modules.{L}.0.blocks.0.block.0.weight" in sd:
        for j in range(2):
            x = _res_block_fwd(sd, f"down_modules.{L}.{j}.", x, mgf, cache)
            order.append(("res", f"down_modules.{L}.{j}."))
        skips.append(x)
        order.append(("skip_push",))
        if f"down_modules.{L}.2.conv.weight" in sd:
            cache[f"down_modules.{L}.2."] = x
            x = conv1d_fwd(x, sd[f"down_modules.{L}.2.conv.weight"], sd[f"down_modules.{L}.2.conv.bias"], stride=2, padding=1)
            order.append(("down", f"down_modules.{L}.2."))
        L += 1
    for m in range(2):
        x = _res_block_fwd(sd, f"mid_modules.{m}.", x, mgf, cache)
        order.append(("res", f"mid_modules.{m}."))
    U = 0
    while f"up_modules.{U}.0.blocks.0.block.0.weight" in sd:
        sk = skips.pop()
        order.append(("cat", x.shape[1]))
        x = torch.cat((x, sk), dim=1)
        for j in range(2):
            x = _res_block_fwd(sd, f"up_modules.{U}.{j}.", x, mgf, cache)
            order.append(("res", f"up_modules.{U}.{j}."))
        if f"up_modules.{U}.2.conv.weight" in sd:
            cache[f"up_modules.{U}.2."] = x
            x = convT1d_fwd(x, sd[f"up_modules.{U}.2.conv.weight"], sd[f"up_modules.{U}.2.conv.bias"])
            order.append(("up", f"up_modules.{U}.2."))
        U += 1
    x = _conv_block_fwd(sd, "final_conv.0.", x, cache)
    cache["final_in"] = x
    x = conv1d_fwd(x, sd["final_conv.1.weight"], sd["final_conv.1.bias"])
    cache["order"] = order
    cache["n_skips_unused"] = len(skips)          # the level-0 skip is pushed but never consumed
    return x.moveaxis(-1, -2), cache


def unet_backward(sd: SD, cache: dict, dout: torch.Tensor):
    """dout [B,T,A] -> (grads of every parameter, d global_cond [B,256], d sample [B,T,A])"""
    grads: Dict[str, torch.Tensor] = {}
    pe, t1, gf, mgf = cache["time"]
    dmgf = torch.zeros_like(mgf)
    dx = dout.moveaxis(-1, -2)
    dx, grads["final_conv.1.weight"], grads["final_conv.1.bias"] = conv1d_bwd(cache["final_in"], sd["final_conv.1.weight"], dx)
    dx = _conv_block_bwd(sd, "final_conv.0.", dx, cache, grads)
    dskips: List[torch.Tensor] = []               # gradients of the consumed skip tensors (filled at the "cat" ops)
    for op in reversed(cache["order"]):
        if op[0] == "res":
            dx, d = _res_block_bwd(sd, op[1], dx, mgf, cache, grads)
            dmgf = dmgf + d
        elif op[0] == "up":
            dx, grads[op[1] + "conv.weight"], grads[op[1] + "conv.bias"] = convT1d_bwd(cache[op[1]], sd[op[1] + "conv.weight"], dx)
        elif op[0] == "down":
            dx, grads[op[1] + "conv.weight"], grads[op[1] + "conv.bias"] = conv1d_bwd(cache[op[1]], sd[op[1] + "conv.weight"], dx,
                                                                                      stride=2, padding=1)
        elif op[0] == "cat":                      # x = cat(x, skip): split the gradient, the skip part waits for its push
            dskips.append(dx[:, op[1]:])
            dx = dx[:, :op[1]]
        elif op[0] == "skip_push":                # walking backwards, the pushes appear deepest level first: that skip was consumed
            if dskips:                            # by the FIRST up level, whose "cat" gradient was appended last; the level-0
                dx = dx + dskips.pop()            # skip is never consumed and gets no gradient
    dgf = dmgf * mish_grad(gf)
    dtemb, dcond = dgf[:, : gf.shape[1] - 256], dgf[:, gf.shape[1] - 256:]
    dm1, grads["diffusion_step_encoder.3.weight"], grads["diffusion_step_encoder.3.bias"] = linear_bwd(
        F.mish(t1), sd["diffusion_step_encoder.3.weight"], dtemb)
    _, grads["diffusion_step_encoder.1.weight"], grads["diffusion_step_encoder.1.bias"] = linear_bwd(
        pe, sd["diffusion_step_encoder.1.weight"], dm1 * mish_grad(t1))
    return grads, dcond, dx.moveaxis(-1, -2)


# ---------------------------------------------------------------------------------------------- losses
def bridge_loss_backward(net_sd: SD, obs_cond, expert_act, vla_act, step, z_unit, beta_max: float = 0.03):
    """get_loss (bridge_model.py:220-246) forward + explicit backward -> (loss, grads with 'b_net.'/'v_net.'/'s_net.' prefixes,
    d obs_cond)."""
    x0, x1 = vla_act, expert_act
    B = x0.shape[0]
    tb = torch.clip(step[:, None, None], T_MIN, 1.0 - T_MIN)
    z = beta_max * z_unit
    xt = (1 - tb) * x0 + tb * x1 + 1.4142 * tb * (1 - tb) * z
    t = torch.clip(step, T_MIN, 1.0 - T_MIN)
    gd = (1.4142 * (1 - 2 * t))[:, None, None]
    targets = {"v_net.": x1 - x0, "s_net.": -z, "b_net.": (x1 - x0) + gd * z}       # loss_n = mean_b(0.5 |o|^2 - <target, o>)
    loss = 0.0
    grads: Dict[str, torch.Tensor] = {}
    dcond = torch.zeros_like(obs_cond)
    for n, tgt in targets.items():
        sd = {k[len(n):]: v for k, v in net_sd.items() if k.startswith(n)}
        out, cache = unet_forward_cached(sd, xt, t, obs_cond)
        loss = loss + torch.mean(0.5 * out.flatten(1).pow(2).sum(-1) - (tgt * out).flatten(1).sum(-1))
        g, dc, _ = unet_backward(sd, cache, (out - tgt) / B)
        grads.update({n + k: v for k, v in g.items()})
        dcond = dcond + dc
    return loss, grads, dcond


# ---------------------------------------------------------------------------------------------- LSTM controller (row a12)
def lstm_loss_backward(mods: Dict[str, SD], vla_act_n, obs_cond, forces, expert_act):
    """TactileLSTMController.get_loss (lstm_step_controller.py:321-337, eval mode: dropout off) forward + explicit
    back-propagation through time, in the decomposition of the planned kernels: the input / head linear maps of ALL time steps
    are batched GEMMs (their weight gradients one GEMM each over B*T rows), only the recurrence itself is sequential
    (per step: gate derivatives, dh_{t-1} = W_hh^T dgates, W_hh gradient accumulated as an outer product).
    -> (loss, grads {module.param}, d obs_cond)."""
    fe, lstm, head = mods["force_encoder"], mods["lstm"], mods["output_head"]
    B, T, A = vla_act_n.shape
    L = 0
    while f"weight_ih_l{L}" in lstm:
        L += 1
    H = lstm["weight_hh_l0"].shape[1]
    # ---- forward, keeping what BPTT needs ----
    f0 = forces.reshape(B * T, -1)
    a1 = F.linear(f0, fe["0.weight"], fe["0.bias"])
    g1 = F.gelu(a1)
    fenc = F.linear(g1, fe["2.weight"], fe["2.bias"]).reshape(B, T, -1)
    x = torch.cat([fenc, vla_act_n], dim=-1)
    inps = [x]
    gates, cs, hs = [], [], []
    for l in range(L):
        xin = inps[l]
        pre_x = F.linear(xin.reshape(B * T, -1), lstm[f"weight_ih_l{l}"], lstm[f"bias_ih_l{l}"]).reshape(B, T, 4 * H)   # batched
        h, c = torch.zeros(B, H), torch.zeros(B, H)
        gl, cl, hl = [], [], []
        for t in range(T):
            g = pre_x[:, t] + F.linear(h, lstm[f"weight_hh_l{l}"], lstm[f"bias_hh_l{l}"])
            i_, f_, g_, o_ = g.chunk(4, dim=-1)
            i_, f_, g_, o_ = torch.sigmoid(i_), torch.sigmoid(f_), torch.tanh(g_), torch.sigmoid(o_)
            c_prev = c
            c = f_ * c + i_ * g_
            h = o_ * torch.tanh(c)
            gl.append((i_, f_, g_, o_, c_prev))
            cl.append(c)
            hl.append(h)
        gates.append(gl)
        cs.append(cl)
        hs.append(torch.stack(hl, dim=1))
        inps.append(hs[-1])
    y = hs[-1]
    comb = torch.cat([y, obs_cond.unsqueeze(1).expand(B, T, obs_cond.shape[-1])], dim=-1).reshape(B * T, -1)
    z0 = F.linear(comb, head["0.weight"], head["0.bias"])
    mu = z0.mean(dim=-1, keepdim=True)
    rstd = torch.rsqrt(z0.var(dim=-1, unbiased=False, keepdim=True) + 1e-5)
    zh = (z0 - mu) * rstd
    z1 = zh * head["1.weight"] + head["1.bias"]
    g2 = F.gelu(z1)
    delta = F.linear(g2, head["4.weight"], head["4.bias"]).reshape(B, T, A)
    out = vla_act_n + delta
    loss = torch.mean((out - expert_act) ** 2)
    # ---- backward ----
    grads: Dict[str, torch.Tensor] = {}
    dout = (2.0 / out.numel()) * (out - expert_act)
    ddelta = dout.reshape(B * T, A)
    dg2, grads["output_head.4.weight"], grads["output_head.4.bias"] = linear_bwd(g2, head["4.weight"], ddelta)
    gelu_grad = lambda v: 0.5 * (1 + torch.erf(v / 2 ** 0.5)) + v * torch.exp(-0.5 * v * v) / (2 * torch.pi) ** 0.5
    dz1 = dg2 * gelu_grad(z1)
    grads["output_head.1.weight"] = (dz1 * zh).sum(dim=0)
    grads["output_head.1.bias"] = dz1.sum(dim=0)
    dzh = dz1 * head["1.weight"]
    dz0 = rstd * (dzh - dzh.mean(dim=-1, keepdim=True) - zh * (dzh * zh).mean(dim=-1, keepdim=True))
    dcomb, grads["output_head.0.weight"], grads["output_head.0.bias"] = linear_bwd(comb, head["0.weight"], dz0)
    dcomb = dcomb.reshape(B, T, -1)
    dcond = dcomb[:, :, H:].sum(dim=1)
    dy = dcomb[:, :, :H]
    for l in reversed(range(L)):
        w_hh = lstm[f"weight_hh_l{l}"]
        dgates_all = torch.zeros(B, T, 4 * H)
        dh_next, dc_next = torch.zeros(B, H), torch.zeros(B, H)
        dw_hh = torch.zeros_like(w_hh)
        for t in reversed(range(T)):
            i_, f_, g_, o_, c_prev = gates[l][t]
            tc = torch.tanh(cs[l][t])
            dh = dy[:, t] + dh_next
            dc = dc_next + dh * o_ * (1 - tc * tc)
            dg = torch.cat([dc * g_ * i_ * (1 - i_), dc * c_prev * f_ * (1 - f_), dc * i_ * (1 - g_ * g_), dh * tc * o_ * (1 - o_)], dim=-1)
            dgates_all[:, t] = dg
            h_prev = hs[l][:, t - 1] if t > 0 else torch.zeros(B, H)
            dw_hh += dg.t() @ h_prev
            dh_next = dg @ w_hh
            dc_next = dc * f_
        grads[f"lstm.weight_hh_l{l}"] = dw_hh
        grads[f"lstm.bias_hh_l{l}"] = dgates_all.sum(dim=(0, 1))
        grads[f"lstm.bias_ih_l{l}"] = grads[f"lstm.bias_hh_l{l}"].clone()
        dxin, grads[f"lstm.weight_ih_l{l}"], _ = linear_bwd(inps[l].reshape(B * T, -1), lstm[f"weight_ih_l{l}"],
                                                            dgates_all.reshape(B * T, 4 * H))
        dy = dxin.reshape(B, T, -1)
    dfenc = dy[:, :, : fenc.shape[-1]].reshape(B * T, -1)
    dg1, grads["force_encoder.2.weight"], grads["force_encoder.2.bias"] = linear_bwd(g1, fe["2.weight"], dfenc)
    _, grads["force_encoder.0.weight"], grads["force_encoder.0.bias"] = linear_bwd(f0, fe["0.weight"], dg1 * gelu_grad(a1))
    return loss, grads, dcond
