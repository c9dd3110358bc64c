"""TEST INFRASTRUCTURE ONLY -- a CPU interpreter of plan descriptors.

Interprets the very structs that cross the C ABI (include/vt_b200.h) with plain torch ops on CPU tensors,
following the documented semantics of each op.  It lets the host-side plan logic (tile boxes, conv taps,
weight packing, buffer wiring, FiLM offsets ...) be validated against the oracle without a GPU; on the
B200 the kernels are then checked against the same expectations.  Never imported by the product package.
"""
import math

import torch
import torch.nn.functional as F

from vla_touch_b200 import native as nv
from vla_touch_b200.plan import TORCH_DT, Plan, tf32_round


def _flat(plan: Plan, address, dtype):
    return plan.resolve(address, dtype)


def _act(x, act, fast=False):
    if act == nv.ACT_GELU:
        return F.gelu(x)
    if act == nv.ACT_MISH:
        return F.mish(x)
    return x


def _store(plan, address, dtype_code, idx, vals, plane=0):
    """scatter vals (float32) to flat element indices idx of the buffer at `address`"""
    dt = TORCH_DT[dtype_code]
    flat = _flat(plan, address, dt)
    if dtype_code == nv.VT_F32 and plane > 0:
        hi = tf32_round(vals.float().contiguous())
        flat[idx.reshape(-1)] = hi.reshape(-1)
        flat[idx.reshape(-1) + plane] = (vals.float() - hi).reshape(-1)
    else:
        flat[idx.reshape(-1)] = vals.reshape(-1).to(dt)


def emu_gemm(plan: Plan, d: nv.GemmDesc):
    in_dt = TORCH_DT[d.in_dtype]
    es_k = 64 if d.in_dtype == nv.VT_BF16 else 32
    assert d.kc % es_k == 0
    G, M, N = d.G, d.M, d.N
    a_flat = _flat(plan, d.a, in_dt)
    w_flat = _flat(plan, d.w, in_dt)
    t_out = M // d.a_B
    m = torch.arange(M)
    b_idx, t_idx = m // t_out, m % t_out
    acc = torch.zeros(G, M, N, dtype=torch.float32)
    passes = [(0, 0)] if d.passes == 1 else [(0, 0), (d.a_plane, 0), (0, d.w_plane)]
    for g in range(G):
        ag = g * d.a_sG if d.a_G > 1 else 0
        w0 = g * d.n_pad * d.w_ld
        W = w_flat[w0: w0 + N * d.w_ld].reshape(N, d.w_ld).float()
        for (pa, pw) in passes:
            cols = []
            for tp in range(d.taps):
                p, dt_ = d.tap_p[tp], d.tap_t[tp]
                tq = t_idx + dt_
                ok = (tq >= 0) & (tq < d.a_T)
                base = ag + b_idx * d.a_sB + (tq.clamp(0, d.a_T - 1) * d.a_P + p) * d.a_ld
                ch = d.a_c0 + pa + torch.arange(d.kc)
                ch_ok = ch < d.a_C
                idx = base[:, None] + ch.clamp(max=d.a_C - 1)[None, :]
                vals = a_flat[idx.reshape(-1)].reshape(M, d.kc).float()
                vals = vals * (ok[:, None] & ch_ok[None, :])
                cols.append(vals)
            A = torch.cat(cols, dim=1)                                  # [M, taps*kc]
            Wp = W[:, pw: pw + d.taps * d.kc]
            if d.in_dtype == nv.VT_F32:                                 # the MMA datapath truncates fp32 to tf32
                A = (A.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
                Wp = (Wp.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
            acc[g] += A @ Wp.t()
    out_dt = TORCH_DT[d.out_dtype]
    q, rem = m // d.row_div, m % d.row_div
    out_rows = q * d.out_q + rem * d.out_r + d.out_off
    res_rows = q * d.res_q + rem * d.res_r + d.res_off
    ncol = torch.arange(N)
    for g in range(G):
        x = acc[g]
        if d.bias:
            x = x + _flat(plan, d.bias, torch.float32)[g * d.n_pad: g * d.n_pad + N]
        if d.epi == nv.EPI_LINEAR:
            x = _act(x, d.act)
            if d.colscale:
                x = x * _flat(plan, d.colscale, torch.float32)[:N]
            if d.res:
                rflat = _flat(plan, d.res, torch.float32)
                x = x + rflat[(g * d.res_g + res_rows[:, None] * d.ldres + ncol[None, :]).reshape(-1)].reshape(M, N)
        else:
            gamma = _flat(plan, d.gn_gamma, torch.float32)[g * d.n_pad: g * d.n_pad + N]
            beta = _flat(plan, d.gn_beta, torch.float32)[g * d.n_pad: g * d.n_pad + N]
            assert d.row_div == d.t_box == t_out
            if d.raw_out:                                               # training forward: conv + bias as GroupNorm sees it
                _store(plan, d.raw_out, nv.VT_F32, g * d.raw_g + m[:, None] * d.raw_ld + ncol[None, :], x, 0)
            xs = x.reshape(d.a_B, t_out, N).permute(0, 2, 1)          # [B, C, T]
            xs = F.group_norm(xs, N // d.gn_group_ch, gamma, beta, d.gn_eps)
            xs = F.mish(xs).permute(0, 2, 1).reshape(M, N)
            if d.film_c:
                fc = _flat(plan, d.film_c, torch.float32)
                sc_idx = g * d.film_g + q[:, None] * d.film_ld + d.film_off + ncol[None, :]
                scale = fc[sc_idx.reshape(-1)].reshape(M, N)
                shift = fc[(sc_idx + d.film_C).reshape(-1)].reshape(M, N)
                if d.film_t:
                    ft = _flat(plan, d.film_t, torch.float32)
                    o = g * d.film_tg + d.film_off
                    scale = scale + ft[o: o + N]
                    shift = shift + ft[o + d.film_C: o + d.film_C + N]
                xs = scale * xs + shift
            if d.res:
                rflat = _flat(plan, d.res, out_dt)
                ridx = (g * d.res_g + res_rows[:, None] * d.ldres + ncol[None, :]).reshape(-1)
                xs = xs + rflat[ridx].reshape(M, N).float()
                if d.res_plane > 0:
                    xs = xs + rflat[ridx + d.res_plane].reshape(M, N).float()
            x = xs
        idx = g * d.out_g + out_rows[:, None] * d.ldc + ncol[None, :]
        _store(plan, d.out, d.out_dtype, idx, x, d.out_plane)


def emu_layernorm(plan, d: nv.LnDesc):
    x = _flat(plan, d.x, torch.float32)
    rows = torch.arange(d.rows)
    idx = rows[:, None] * d.in_row_stride * d.in_ld + torch.arange(d.D)[None, :]
    v = x[idx.reshape(-1)].reshape(d.rows, d.D)
    y = F.layer_norm(v, (d.D,), _flat(plan, d.gamma, torch.float32)[: d.D], _flat(plan, d.beta, torch.float32)[: d.D], d.eps)
    if d.act == nv.ACT_GELU:
        y = F.gelu(y)
    _store(plan, d.out, d.out_dtype, rows[:, None] * d.out_ld + torch.arange(d.D)[None, :], y, d.out_plane)


def emu_attention(plan, d: nv.AttnDesc):
    dt = TORCH_DT[d.in_dtype]
    D = d.heads * 64
    n = d.images * d.tokens
    qkv = _flat(plan, d.qkv, dt)[: n * 3 * D].reshape(d.images, d.tokens, 3, d.heads, 64).float()
    q, k, v = (qkv[:, :, i].permute(0, 2, 1, 3) for i in range(3))
    att = torch.softmax(q @ k.transpose(-1, -2) * 0.125, dim=-1)
    if d.in_dtype == nv.VT_BF16:
        att = att.to(torch.bfloat16).float()   # P is handed to the tensor cores as bf16 (un-normalised in the kernel)
    ctx = (att @ v).permute(0, 2, 1, 3).reshape(n, D)
    _store(plan, d.ctx, d.in_dtype, torch.arange(n)[:, None] * d.ctx_ld + torch.arange(D)[None, :], ctx, d.ctx_plane)


def emu_mlp(plan, d: nv.MlpDesc):
    """h += ls2 * (bf16(GELU(xn W1^T + b1)) W2^T + b2): the hidden activation is handed to the second MMA as bf16."""
    D, H = d.D, 4 * d.D
    r = torch.arange(d.rows)
    xn = _flat(plan, d.xn, torch.bfloat16)[(r[:, None] * d.ld_x + torch.arange(D)[None, :]).reshape(-1)].reshape(d.rows, D).float()
    w1 = _flat(plan, d.w1, torch.bfloat16)[(torch.arange(H)[:, None] * d.w1_ld + torch.arange(D)[None, :]).reshape(-1)].reshape(H, D).float()
    w2 = _flat(plan, d.w2, torch.bfloat16)[(torch.arange(D)[:, None] * d.w2_ld + torch.arange(H)[None, :]).reshape(-1)].reshape(D, H).float()
    b1, b2 = _flat(plan, d.b1, torch.float32)[:H], _flat(plan, d.b2, torch.float32)[:D]
    hid = F.gelu(xn @ w1.t() + b1).to(torch.bfloat16).float()
    y = hid @ w2.t() + b2
    if d.ls2:
        y = y * _flat(plan, d.ls2, torch.float32)[:D]
    hf = _flat(plan, d.h, torch.float32)
    idx = (r[:, None] * d.ld_h + torch.arange(D)[None, :]).reshape(-1)
    hn = y + hf[idx].reshape(d.rows, D)
    hf[idx] = hn.reshape(-1)
    if d.ln_out:                                   # the next block's norm1, written by the same kernel
        g, b = _flat(plan, d.ln_gamma, torch.float32)[:D], _flat(plan, d.ln_beta, torch.float32)[:D]
        yn = F.layer_norm(hn, (D,), g, b, d.ln_eps).to(torch.bfloat16)
        of = _flat(plan, d.ln_out, torch.bfloat16)
        of[(r[:, None] * d.ln_ld + torch.arange(D)[None, :]).reshape(-1)] = yn.reshape(-1)


def emu_rowproj(plan, d: nv.RowprojDesc):
    """h += colscale * (x W^T + bias), then (optionally) LayerNorm of the updated rows as bf16: what the attn_out GEMM descriptor
    followed by the norm2 LayerNorm descriptor compute."""
    D = d.D
    r = torch.arange(d.rows)
    x = _flat(plan, d.x, torch.bfloat16)[(r[:, None] * d.ld_x + torch.arange(D)[None, :]).reshape(-1)].reshape(d.rows, D).float()
    w = _flat(plan, d.w, torch.bfloat16)[(torch.arange(D)[:, None] * d.w_ld + torch.arange(D)[None, :]).reshape(-1)].reshape(D, D).float()
    y = x @ w.t() + _flat(plan, d.bias, torch.float32)[:D]
    if d.colscale:
        y = y * _flat(plan, d.colscale, torch.float32)[:D]
    hf = _flat(plan, d.h, torch.float32)
    idx = (r[:, None] * d.ld_h + torch.arange(D)[None, :]).reshape(-1)
    hn = y + hf[idx].reshape(d.rows, D)
    hf[idx] = hn.reshape(-1)
    if d.ln_out:
        g, b = _flat(plan, d.ln_gamma, torch.float32)[:D], _flat(plan, d.ln_beta, torch.float32)[:D]
        yn = F.layer_norm(hn, (D,), g, b, d.ln_eps).to(torch.bfloat16)
        of = _flat(plan, d.ln_out, torch.bfloat16)
        of[(r[:, None] * d.ln_ld + torch.arange(D)[None, :]).reshape(-1)] = yn.reshape(-1)


def emu_imgstats(plan, d: nv.ImgStatsDesc):
    img = _flat(plan, d.img, TORCH_DT[d.dtype])[: d.count]
    mx = float(img.max())
    mean = float(img.double().mean())
    if mx > 1.0:
        mean /= 255.0
    flags = _flat(plan, d.flags, torch.int32)
    flags[0] = 1 if mx > 1.0 else 0
    flags[1] = 0 if torch.tensor(mean, dtype=torch.float32) < 0.5 else 1


def emu_patchify(plan, d: nv.PatchifyDesc):
    n = d.images * d.H * d.W * 3
    img = _flat(plan, d.img, TORCH_DT[d.dtype])[:n]
    img = img.reshape(d.images, d.H, d.W, 3).permute(0, 3, 1, 2) if d.layout == nv.LAYOUT_BHWC else img.reshape(d.images, 3, d.H, d.W)
    img = img.float()[:, :, : d.H // d.patch * d.patch, : d.W // d.patch * d.patch]
    flags = _flat(plan, d.flags, torch.int32)
    if int(flags[0]):
        img = img / 255.0
    if int(flags[1]):
        img = (img - torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)) / torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    cols = F.unfold(img, d.patch, stride=d.patch).transpose(1, 2)      # [B, np, 3*p*p], k = c*p*p + i*p + j
    rows = cols.shape[0] * cols.shape[1]
    out = torch.zeros(rows, d.out_cols)
    out[:, : cols.shape[2]] = cols.reshape(rows, -1)
    _store(plan, d.out, d.out_dtype, torch.arange(rows)[:, None] * d.out_ld + torch.arange(d.out_cols)[None, :], out, d.out_plane)


def emu_cls(plan, d: nv.ClsDesc):
    h = _flat(plan, d.h, torch.float32)
    row = _flat(plan, d.cls, torch.float32)[: d.D] + _flat(plan, d.pos, torch.float32)[: d.D]
    for b in range(d.images):
        h[b * d.tokens * d.D: b * d.tokens * d.D + d.D] = row


def emu_pack(plan, d: nv.PackDesc):
    src = _flat(plan, d.src, torch.float32)
    rows = torch.arange(d.rows)
    srows = rows // d.src_row_div if d.src_row_div > 1 else rows
    v = src[(srows[:, None] * d.src_ld + torch.arange(d.cols)[None, :]).reshape(-1)].reshape(d.rows, d.cols)
    v = _act(v, d.act)
    width = max(d.cols, d.zero_to)
    full = torch.zeros(d.rows, width)
    full[:, : d.cols] = v
    _store(plan, d.out, d.out_dtype, rows[:, None] * d.out_ld + d.dst_c0 + torch.arange(width)[None, :], full, d.out_plane)


def emu_affine(plan, d: nv.AffineDesc):
    n = d.rows * d.A
    x = _flat(plan, d.x, torch.float32)[:n].reshape(d.rows, d.A).clone()
    if d.add:
        x = x + _flat(plan, d.add, torch.float32)[:n].reshape(d.rows, d.A)
    mins = _flat(plan, d.mins, torch.float32)[: d.A]
    maxs = _flat(plan, d.maxs, torch.float32)[: d.A]
    padded = (maxs - mins) * torch.tensor(d.pad, dtype=torch.float32)
    center = (mins + maxs) / 2
    pmin, pmax = center - padded / 2, center + padded / 2
    rng = pmax - pmin
    if d.denorm == 2:
        y = x
    elif not d.denorm:
        rng = torch.where(rng < 1e-6, torch.ones_like(rng), rng)
        y = 2.0 * (x - pmin) / rng - 1.0
    else:
        y = (x + 1.0) / 2.0 * rng + pmin
    if d.out:
        _flat(plan, d.out, torch.float32)[:n] = y.reshape(-1)
    if d.xpad:
        rows = torch.arange(d.rows)
        _store(plan, d.xpad, d.xpad_dtype, rows[:, None] * d.xpad_ld + torch.arange(d.A)[None, :], y, d.xpad_plane)


def emu_tembed(plan, d: nv.TembedDesc):
    t = _flat(plan, d.t, torch.float32)[: d.rows]
    half = d.dim // 2
    e = math.log(10000) / (half - 1)
    f = torch.exp(torch.arange(half) * -e)
    arg = t[:, None] * f[None, :]
    out = torch.cat((arg.sin(), arg.cos()), dim=-1)
    rows = torch.arange(d.rows)
    _store(plan, d.out, d.out_dtype, rows[:, None] * d.out_ld + torch.arange(d.dim)[None, :], out, d.out_plane)


def emu_sde(plan, d: nv.SdeDesc):
    n = d.rows * d.A
    x = _flat(plan, d.x, torch.float32)[:n]
    v = _flat(plan, d.v, torch.float32)[:n]
    s = _flat(plan, d.s, torch.float32)[:n]
    assert d.noise, "the emulator needs injected noise"
    z = _flat(plan, d.noise, torch.float32)[:n]
    f = lambda c: torch.tensor(c, dtype=torch.float32)
    sv = s * f(d.ginv)
    b = v - f(d.dgg) * sv * f(d.eps)
    nx = x + (b + f(d.eps) * sv) * f(d.dt)
    nx = nx + f(d.nscale) * (f(d.d) * z)
    x[:] = nx
    if d.xpad:
        rows = torch.arange(d.rows)
        _store(plan, d.xpad, d.xpad_dtype, rows[:, None] * d.xpad_ld + torch.arange(d.A)[None, :], nx.reshape(d.rows, d.A), d.xpad_plane)


def emu_lstm(plan, d: nv.LstmDesc):
    H = d.H
    xw = _flat(plan, d.xw, torch.float32)[: d.B * d.T * 4 * H].reshape(d.B, d.T, 4 * H)
    w = _flat(plan, d.w_hh, torch.float32)[: 4 * H * H].reshape(H, 4 * H)     # transposed: [H][4H]
    h = _flat(plan, d.h, torch.float32)[: d.B * H].reshape(d.B, H)
    c = _flat(plan, d.c, torch.float32)[: d.B * H].reshape(d.B, H)
    hh, cc = h.clone(), c.clone()
    rows = torch.arange(d.B)
    for t in range(d.T):
        g = xw[:, t] + hh @ w
        i_, f_, g_, o_ = g.chunk(4, dim=-1)
        cc = torch.sigmoid(f_) * cc + torch.sigmoid(i_) * torch.tanh(g_)
        hh = torch.sigmoid(o_) * torch.tanh(cc)
        _store(plan, d.y, d.y_dtype, (rows[:, None] * d.T + t) * d.y_ld + torch.arange(H)[None, :], hh, d.y_plane)
    h[:] = hh
    c[:] = cc


def emu_qsample(plan, d: nv.QsampleDesc):
    N = d.B * d.n
    x0 = _flat(plan, d.x0, torch.float32)[:N].reshape(d.B, d.n)
    x1 = _flat(plan, d.x1, torch.float32)[:N].reshape(d.B, d.n)
    z = _flat(plan, d.z_unit, torch.float32)[:N].reshape(d.B, d.n) * torch.tensor(d.d, dtype=torch.float32)
    t = torch.clip(_flat(plan, d.step, torch.float32)[: d.B], 0.001, 1.0 - 0.001)
    tb = t[:, None]
    gamma = 1.4142 * tb * (1 - tb)
    xt = (1 - tb) * x0 + tb * x1 + gamma * z
    _flat(plan, d.xt, torch.float32)[:N] = xt.reshape(-1)
    _flat(plan, d.tclip, torch.float32)[: d.B] = t
    if d.xpad:
        rows = torch.arange(N // d.A)
        _store(plan, d.xpad, d.xpad_dtype, rows[:, None] * d.xpad_ld + torch.arange(d.A)[None, :], xt.reshape(-1, d.A), d.xpad_plane)


def emu_siloss(plan, d: nv.SilossDesc):
    N = d.B * d.n
    bvs = _flat(plan, d.bvs, torch.float32)[: 3 * N].reshape(3, d.B, d.n)
    x0 = _flat(plan, d.x0, torch.float32)[:N].reshape(d.B, d.n)
    x1 = _flat(plan, d.x1, torch.float32)[:N].reshape(d.B, d.n)
    z = _flat(plan, d.z_unit, torch.float32)[:N].reshape(d.B, d.n) * torch.tensor(d.d, dtype=torch.float32)
    t = _flat(plan, d.tclip, torch.float32)[: d.B]
    gd = (1.4142 * (1 - 2 * t))[:, None]
    pt = x1 - x0
    b, v, s = bvs[0], bvs[1], bvs[2]
    lv = (0.5 * (v * v).sum(-1) - (pt * v).sum(-1)).mean()
    ls = (0.5 * (s * s).sum(-1) + (z * s).sum(-1)).mean()
    lb = (0.5 * (b * b).sum(-1) - ((pt + gd * z) * b).sum(-1)).mean()
    out = _flat(plan, d.out, torch.float32)
    out[0], out[1], out[2], out[3] = lv + ls + lb, lv, ls, lb


def emu_tcol(plan, d: nv.TcolDesc):
    src = _flat(plan, d.src, TORCH_DT[d.src_dtype])
    out = _flat(plan, d.out, torch.bfloat16)
    K = d.B * d.t_out
    k = torch.arange(K)
    b, t = k // d.t_out, k % d.t_out
    c = torch.arange(d.C)
    for g in range(d.G):
        for tap in range(d.taps):
            pos = t * d.stride + d.tap_off[tap]
            ok = (pos >= 0) & (pos < d.T_src)
            idx = g * d.sG + b * d.sB + pos.clamp(0, d.T_src - 1) * d.ld                       # [K]
            vals = src[(idx[None, :] + c[:, None]).reshape(-1)].reshape(d.C, K).float() * ok[None, :]
            oidx = g * d.out_g + (tap * d.c_pad + c[:, None]) * d.k_ld + k[None, :]
            out[oidx.reshape(-1)] = vals.reshape(-1).to(torch.bfloat16)


def _mish_grad(y):
    tsp = torch.tanh(F.softplus(y))
    return tsp + y * (1 - tsp * tsp) * torch.sigmoid(y)


def emu_gnbwd(plan, d: nv.GnbwdDesc):
    G, B, T, C = d.G, d.B, d.T, d.C
    raw = _flat(plan, d.raw, torch.float32)[: G * B * T * C].reshape(G, B, T, C)
    doutf = _flat(plan, d.dout, torch.float32)
    draw = _flat(plan, d.draw, torch.bfloat16)
    cc = torch.arange(C)
    rows = torch.arange(B * T)
    for g in range(G):
        go = doutf[(g * d.dout_g + rows[:, None] * d.dout_ld + cc[None, :]).reshape(-1)].reshape(B, T, C)
        gamma = _flat(plan, d.gamma, torch.float32)[g * d.p_ld: g * d.p_ld + C]
        beta = _flat(plan, d.beta, torch.float32)[g * d.p_ld: g * d.p_ld + C]
        r = raw[g].permute(0, 2, 1).reshape(B, d.groups, -1)                                 # [B, groups, Cg*T]
        mean = r.mean(dim=-1, keepdim=True)
        rstd = torch.rsqrt(r.var(dim=-1, unbiased=False, keepdim=True) + d.eps)
        xh = ((r - mean) * rstd).reshape(B, C, T).permute(0, 2, 1)                           # [B, T, C]
        y = xh * gamma + beta
        m = F.mish(y)
        dm = go
        if d.film:
            film = _flat(plan, d.film, torch.float32)
            fidx = g * d.film_g + torch.arange(B)[:, None] * d.film_ld + d.film_off + cc[None, :]
            scale = film[fidx.reshape(-1)].reshape(B, 1, C)
            dm = go * scale
            if d.dfilm:
                df = _flat(plan, d.dfilm, torch.float32)
                df[fidx.reshape(-1)] = (go * m).sum(dim=1).reshape(-1)
                df[(fidx + C).reshape(-1)] = go.sum(dim=1).reshape(-1)
        da = dm * _mish_grad(y)
        dxh = (da * gamma).permute(0, 2, 1).reshape(B, d.groups, -1)
        xg = xh.permute(0, 2, 1).reshape(B, d.groups, -1)
        dr = rstd * (dxh - dxh.mean(dim=-1, keepdim=True) - xg * (dxh * xg).mean(dim=-1, keepdim=True))
        dr = dr.reshape(B, C, T).permute(0, 2, 1)                                            # [B, T, C]
        draw[g * B * T * C: (g + 1) * B * T * C] = dr.reshape(-1).to(torch.bfloat16)
        for ptr_, val in ((d.dgamma, (da * xh).sum(dim=(0, 1))), (d.dbeta, da.sum(dim=(0, 1))), (d.dbias, dr.sum(dim=(0, 1)))):
            if ptr_:
                _flat(plan, ptr_, torch.float32)[g * d.p_ld: g * d.p_ld + C] = val


def emu_colsum(plan, d: nv.ColsumDesc):
    x = _flat(plan, d.x, torch.float32)
    out = _flat(plan, d.out, torch.float32)
    rows, cc = torch.arange(d.rows), torch.arange(d.C)
    for g in range(d.G):
        v = x[(g * d.x_g + rows[:, None] * d.ld + cc[None, :]).reshape(-1)].reshape(d.rows, d.C)
        out[g * d.out_ld: g * d.out_ld + d.C] = v.sum(dim=0)


def emu_ewise(plan, d: nv.EwiseDesc):
    a, b, out = (_flat(plan, x, torch.float32) for x in (d.a, d.b, d.out))
    r, c = torch.arange(d.rows)[:, None], torch.arange(d.cols)[None, :]
    x, y = a[(r * d.a_ld + c).reshape(-1)], b[(r * d.b_ld + c).reshape(-1)]
    gelu_grad = lambda v: 0.5 * (1 + torch.erf(v / 2 ** 0.5)) + v * torch.exp(-0.5 * v * v) / (2 * math.pi) ** 0.5
    res = {nv.EW_ADD: lambda: x + y, nv.EW_MISH_BWD: lambda: x * _mish_grad(y), nv.EW_GELU_BWD: lambda: x * gelu_grad(y),
           nv.EW_MUL: lambda: x * y, nv.EW_SCALED_DIFF: lambda: torch.tensor(d.alpha, dtype=torch.float32) * (x - y)}[d.op]()
    out[(r * d.out_ld + c).reshape(-1)] = res


def emu_silossbwd(plan, d: nv.SilossBwdDesc):
    N = d.B * d.n
    bvs = _flat(plan, d.bvs, torch.float32)[: 3 * N].reshape(3, d.B, d.n)
    x0 = _flat(plan, d.x0, torch.float32)[:N].reshape(d.B, d.n)
    x1 = _flat(plan, d.x1, torch.float32)[:N].reshape(d.B, d.n)
    z = _flat(plan, d.z_unit, torch.float32)[:N].reshape(d.B, d.n) * torch.tensor(d.d, dtype=torch.float32)
    gd = (1.4142 * (1 - 2 * _flat(plan, d.tclip, torch.float32)[: d.B]))[:, None]
    pt = x1 - x0
    tgt = torch.stack([pt + gd * z, pt, -z])
    _flat(plan, d.dvs, torch.float32)[: 3 * N] = ((bvs - tgt) / d.B).reshape(-1)


def emu_lstm_train(plan, d: nv.LstmTrainDesc):
    H, B, T = d.H, d.B, d.T
    xw = _flat(plan, d.xw, torch.float32)[: B * T * 4 * H].reshape(B, T, 4 * H)
    w = _flat(plan, d.w_hh, torch.float32)[: 4 * H * H].reshape(H, 4 * H)     # transposed: [H][4H]
    gates = _flat(plan, d.gates, torch.float32)[: B * T * 4 * H].reshape(B, T, 4 * H)
    call = _flat(plan, d.c, torch.float32)[: B * T * H].reshape(B, T, H)
    hh, cc = torch.zeros(B, H), torch.zeros(B, H)
    rows = torch.arange(B)
    for t in range(T):
        i_, f_, g_, o_ = (xw[:, t] + hh @ w).chunk(4, dim=-1)
        i_, f_, g_, o_ = torch.sigmoid(i_), torch.sigmoid(f_), torch.tanh(g_), torch.sigmoid(o_)
        cc = f_ * cc + i_ * g_
        hh = o_ * torch.tanh(cc)
        gates[:, t] = torch.cat((i_, f_, g_, o_), dim=-1)
        call[:, t] = cc
        _store(plan, d.y, d.y_dtype, (rows[:, None] * T + t) * d.y_ld + torch.arange(H)[None, :], hh, 0)


def emu_lstm_bwd(plan, d: nv.LstmBwdDesc):
    H, B, T = d.H, d.B, d.T
    gates = _flat(plan, d.gates, torch.float32)[: B * T * 4 * H].reshape(B, T, 4 * H)
    call = _flat(plan, d.c, torch.float32)[: B * T * H].reshape(B, T, H)
    dyf = _flat(plan, d.dy, torch.float32)
    w = _flat(plan, d.w_hh, torch.float32)[: 4 * H * H].reshape(4 * H, H)
    dg_all = _flat(plan, d.dgates, torch.float32)[: B * T * 4 * H].reshape(B, T, 4 * H)
    rows = torch.arange(B)
    dh_next, dc_next = torch.zeros(B, H), torch.zeros(B, H)
    for t in reversed(range(T)):
        i_, f_, g_, o_ = gates[:, t].chunk(4, dim=-1)
        tc = torch.tanh(call[:, t])
        cp = call[:, t - 1] if t > 0 else torch.zeros(B, H)
        dy = dyf[((rows[:, None] * T + t) * d.dy_ld + torch.arange(H)[None, :]).reshape(-1)].reshape(B, H)
        dh = dy + dh_next
        dc = dc_next + dh * o_ * (1 - tc * tc)
        dg = torch.cat((dc * g_ * i_ * (1 - i_), dc * cp * f_ * (1 - f_), dc * i_ * (1 - g_ * g_), dh * tc * o_ * (1 - o_)), dim=-1)
        dg_all[:, t] = dg
        dh_next = dg @ w
        dc_next = dc * f_


def emu_lngelubwd(plan, d: nv.LnGeluBwdDesc):
    n, D = d.rows * d.D, d.D
    z0 = _flat(plan, d.z0, torch.float32)[:n].reshape(d.rows, D)
    dzn = _flat(plan, d.dzn, torch.float32)[:n].reshape(d.rows, D)
    g, b = _flat(plan, d.gamma, torch.float32)[:D], _flat(plan, d.beta, torch.float32)[:D]
    mu = z0.mean(dim=-1, keepdim=True)
    rstd = torch.rsqrt(z0.var(dim=-1, unbiased=False, keepdim=True) + d.eps)
    zh = (z0 - mu) * rstd
    z1 = zh * g + b
    d1 = dzn * (0.5 * (1 + torch.erf(z1 / 2 ** 0.5)) + z1 * torch.exp(-0.5 * z1 * z1) / (2 * math.pi) ** 0.5)
    dzh = d1 * g
    _flat(plan, d.dz0, torch.float32)[:n] = (rstd * (dzh - dzh.mean(dim=-1, keepdim=True) - zh * (dzh * zh).mean(dim=-1, keepdim=True))).reshape(-1)
    _flat(plan, d.d1, torch.float32)[:n] = d1.reshape(-1)
    _flat(plan, d.d1zh, torch.float32)[:n] = (d1 * zh).reshape(-1)


def emu_dropmask(plan, d: nv.DropmaskDesc):
    assert d.inject, "the emulator needs injected uniforms"
    u = _flat(plan, d.inject, torch.float32)[: d.n]
    p = torch.tensor(d.p, dtype=torch.float32)
    _flat(plan, d.mask, torch.float32)[: d.n] = torch.where(u >= p, 1.0 / (1.0 - p), torch.zeros(()))


def emu_wgrad(plan, d: nv.WgradDesc):
    """out[g][r][tap * c_pad + c] = sum_{b,t} rows[g][b][t + rows_t (phase rows_p)][r] * cols[g][b][t + tap_t (phase tap_p)][c]"""
    bf = torch.bfloat16
    rows, cols = _flat(plan, d.rows, bf), _flat(plan, d.cols, bf)
    out = _flat(plan, d.out, torch.float32)
    b = torch.arange(d.B)[:, None]
    t = torch.arange(d.t_out)[None, :]

    def gather(flat, g, sG, sB, ld, P, T, p, dt, c0, C, C_vis):
        tq = t + dt
        ok = (tq >= 0) & (tq < T)
        base = g * sG + b * sB + (tq.clamp(0, T - 1) * P + p) * ld                     # [B, t_out]
        ch = c0 + torch.arange(C)
        vis = ch < C_vis
        idx = base[:, :, None] + ch.clamp(max=C_vis - 1)[None, None, :]
        v = flat[idx.reshape(-1)].reshape(d.B, d.t_out, C).float()
        return v * (ok[:, :, None] & vis[None, None, :])

    for g in range(d.G):
        X = gather(rows, g, d.rows_sG, d.rows_sB, d.rows_ld, d.rows_P, d.rows_T, d.rows_p, d.rows_t, 0, d.R, d.rows_C).reshape(-1, d.R)
        for tap in range(d.taps):
            Y = gather(cols, g, d.cols_sG, d.cols_sB, d.cols_ld, d.cols_P, d.cols_T, d.tap_p[tap], d.tap_t[tap], 0, d.c_pad, d.cols_C).reshape(-1, d.c_pad)
            blk = X.t() @ Y                                                            # [R, c_pad]
            rr = torch.arange(d.R)[:, None]
            cc = tap * d.c_pad + torch.arange(d.c_pad)[None, :]
            out[(g * d.out_g + rr * d.ldc + cc).reshape(-1)] = blk.reshape(-1)


def emu_persist(plan, d: nv.PersistDesc):
    """The persistent multi-layer launch = its layers in order, step by step (film_t / noise rows and the Euler-Maruyama scalars of
    the step substituted exactly as the kernel does)."""
    clone = lambda st: type(st).from_buffer_copy(st)
    for step in range(d.n_steps):
        for g in d.layers:
            if g.film_t and d.film_t_step:
                g = clone(g)
                g.film_t = g.film_t + 4 * step * d.film_t_step
            emu_gemm(plan, g)
        if d.py_sde is not None:
            s = clone(d.py_sde)
            s.ginv, s.dgg, s.eps, s.dt, s.nscale = (float(x) for x in d.coef[step])
            s.step = step
            if s.noise:
                s.noise = s.noise + 4 * step * d.noise_step
            emu_sde(plan, s)


_EMU = {nv.WgradDesc: emu_wgrad, nv.PersistDesc: emu_persist, nv.QsampleDesc: emu_qsample, nv.SilossDesc: emu_siloss, nv.GemmDesc: emu_gemm, nv.LnDesc: emu_layernorm, nv.AttnDesc: emu_attention, nv.ImgStatsDesc: emu_imgstats,
        nv.PatchifyDesc: emu_patchify, nv.ClsDesc: emu_cls, nv.PackDesc: emu_pack, nv.AffineDesc: emu_affine,
        nv.TcolDesc: emu_tcol, nv.GnbwdDesc: emu_gnbwd, nv.ColsumDesc: emu_colsum, nv.EwiseDesc: emu_ewise, nv.SilossBwdDesc: emu_silossbwd, nv.LstmTrainDesc: emu_lstm_train, nv.LstmBwdDesc: emu_lstm_bwd, nv.LnGeluBwdDesc: emu_lngelubwd, nv.DropmaskDesc: emu_dropmask, nv.TembedDesc: emu_tembed, nv.SdeDesc: emu_sde, nv.LstmDesc: emu_lstm, nv.MlpDesc: emu_mlp, nv.RowprojDesc: emu_rowproj}


@torch.no_grad()
def run(plan: Plan, first: int = 0, count: int = -1):
    n = len(plan.descs)
    last = n if count < 0 else first + count
    for i in range(first, last):
        _EMU[type(plan.descs[i])](plan, plan.descs[i])
