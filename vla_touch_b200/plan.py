"""Plans: host-side descriptions of the kernel sequences that replace the reference's nn.Module calls.

A `Plan` is a list of op descriptors (the ctypes structs of vla_touch_b200.native, i.e. exactly what crosses
the C ABI) plus the device buffers they point into.  `Plan.compile()` hands the descriptors to libvt_b200.so,
which encodes the TMA tensor maps and returns a replayable native program.  Building a plan touches no GPU API
besides tensor allocation, so the descriptor logic is unit-tested on CPU tensors by interpreting the very same
descriptors (tests/plan_emu.py); the kernels are then checked against that interpretation on the B200.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, List, Optional, Sequence, Tuple

import torch

from . import native as nv

TORCH_DT = {nv.VT_BF16: torch.bfloat16, nv.VT_F32: torch.float32, nv.VT_U8: torch.uint8}
VT_DT = {torch.bfloat16: nv.VT_BF16, torch.float32: nv.VT_F32, torch.uint8: nv.VT_U8}


def round_up(x: int, m: int) -> int:
    return (x + m - 1) // m * m


def ptr(t: Optional[torch.Tensor], offset_elems: int = 0) -> Optional[int]:
    if t is None:
        return None
    return t.data_ptr() + offset_elems * t.element_size()


class Plan:
    def __init__(self, device) -> None:
        self.device = torch.device(device)
        self.descs: List[C.Structure] = []
        self.tags: List[str] = []
        self.bufs: Dict[str, torch.Tensor] = {}
        self._reg: List[torch.Tensor] = []
        self._program: Optional[nv.Program] = None
        self._arenas: Dict[str, list] = {}      # name -> [flat tensor, elements used, [(offset, numel, op index at allocation)]]

    # ---- memory ----
    def reg(self, t: torch.Tensor) -> torch.Tensor:
        """Register an externally owned tensor (weights, inputs) so that it stays alive and resolvable."""
        assert t.is_contiguous(), "plan tensors must be contiguous"
        assert t.device.type == self.device.type, f"tensor on {t.device}, plan on {self.device}"
        self._reg.append(t)
        return t

    def set_arena(self, arena: str, numel: int, dtype: torch.dtype = torch.float32) -> torch.Tensor:
        """One contiguous allocation that `buf(..., arena=...)` carves its tensors from, in allocation order (the training
        programs put every parameter gradient in one arena: the data-parallel all-reduce then runs in place on slices of it)."""
        t = torch.zeros(numel, dtype=dtype, device=self.device)
        self._arenas[arena] = [t, 0, []]
        self._reg.append(t)
        return t

    def arena(self, arena: str):
        """(flat tensor, elements used, [(offset, numel, ops in the plan when it was allocated)]) or None."""
        a = self._arenas.get(arena)
        return None if a is None else (a[0], a[1], a[2])

    def buf(self, name: str, shape: Sequence[int], dtype: torch.dtype, zero: bool = True, arena: Optional[str] = None) -> torch.Tensor:
        assert name not in self.bufs, name
        a = self._arenas.get(arena) if arena else None
        if a is not None and a[0].dtype == dtype:
            n = 1
            for d in shape:
                n *= int(d)
            off = a[1]
            if off + n > a[0].numel():
                raise RuntimeError(f"plan arena '{arena}' is too small for '{name}' ({off} + {n} > {a[0].numel()} elements)")
            t = a[0][off: off + n].view(tuple(shape))
            a[2].append((off, n, len(self.descs)))
            a[1] = off + (n + 63) // 64 * 64           # 256-byte granules: every tensor stays 16-byte (TMA / float4) aligned
            self.bufs[name] = t
            return t
        t = (torch.zeros if zero else torch.empty)(tuple(shape), dtype=dtype, device=self.device)
        self.bufs[name] = t
        self._reg.append(t)
        return t

    def resolve(self, address: int, dtype: torch.dtype) -> torch.Tensor:
        """Flat view of the registered tensor containing `address`, starting at that address (emulator use)."""
        for t in self._reg:
            base = t.data_ptr()
            if base <= address < base + t.numel() * t.element_size():
                assert t.dtype == dtype, f"buffer dtype {t.dtype} read as {dtype}"
                off = (address - base) // t.element_size()
                return t.view(-1)[off:]
        raise KeyError(f"address {address:#x} is not inside a registered plan tensor")

    # ---- ops ----
    def add(self, desc: C.Structure, tag: str = "") -> int:
        self.descs.append(desc)
        self.tags.append(tag)
        if self._program is not None:
            raise RuntimeError("plan already compiled")
        return len(self.descs) - 1

    def __len__(self) -> int:
        return len(self.descs)

    def compile(self) -> nv.Program:
        if self._program is None:
            if self.device.type != "cuda":
                raise nv.NativeError("plans execute on a CUDA device only (no CPU fallback)")
            prog = nv.Program()
            for d, tag in zip(self.descs, self.tags):
                try:
                    prog.add(d)
                except nv.NativeError as e:
                    raise nv.NativeError(f"op '{tag}': {e}") from None
            prog.keep = self._reg
            self._program = prog
        return self._program


# ------------------------------------------------------------------------------------------------
# descriptor helpers
# ------------------------------------------------------------------------------------------------
def gemm_desc(*, a, in_dtype, a_C, a_T, a_B=1, a_P=1, a_G=1, a_ld, a_sB=0, a_sG=0, a_c0=0, kc, taps=((0, 0),),
              t_box, b_box, w, n_pad, w_ld, G=1, M, N, bn, out, out_dtype, ldc, out_g=0, row_div=1, out_q=1, out_r=0,
              out_off=0, out_plane=0, epi=nv.EPI_LINEAR, act=nv.ACT_NONE, bias=None, colscale=None, res=None, ldres=0,
              res_g=0, res_q=0, res_r=0, res_off=0, res_plane=0, gn_gamma=None, gn_beta=None, gn_group_ch=0, gn_eps=1e-5,
              film_c=None, film_t=None, film_g=0, film_tg=0, film_ld=0, film_C=0, film_off=0, passes=1, a_plane=0,
              w_plane=0, raw_out=None, raw_g=0, raw_ld=0) -> nv.GemmDesc:
    d = nv.GemmDesc()
    d.a, d.in_dtype, d.a_C, d.a_P, d.a_T, d.a_B, d.a_G = a, in_dtype, a_C, a_P, a_T, a_B, a_G
    d.a_ld, d.a_sB, d.a_sG, d.a_c0, d.kc = a_ld, a_sB if a_sB else a_ld * a_T * a_P, a_sG, a_c0, kc
    d.taps = len(taps)
    for i, (p, t) in enumerate(taps):
        d.tap_p[i], d.tap_t[i] = p, t
    d.t_box, d.b_box, d.passes, d.a_plane, d.w_plane = t_box, b_box, passes, a_plane, w_plane
    d.w, d.n_pad, d.w_ld = w, n_pad, w_ld
    d.G, d.M, d.N, d.bn = G, M, N, bn
    d.out, d.out_dtype, d.ldc, d.out_g, d.row_div = out, out_dtype, ldc, out_g, row_div
    d.out_q, d.out_r, d.out_off, d.out_plane = out_q, out_r, out_off, out_plane
    d.epi, d.act, d.bias, d.colscale, d.res, d.ldres = epi, act, bias, colscale, res, ldres
    d.res_g, d.res_q, d.res_r, d.res_off, d.res_plane = res_g, res_q, res_r, res_off, res_plane
    d.gn_gamma, d.gn_beta, d.gn_group_ch, d.gn_eps = gn_gamma, gn_beta, gn_group_ch, gn_eps
    d.film_c, d.film_t, d.film_g, d.film_tg = film_c, film_t, film_g, film_tg
    d.film_ld, d.film_C, d.film_off = film_ld, film_C, film_off
    d.raw_out, d.raw_g, d.raw_ld = raw_out, raw_g, raw_ld
    return d


def pick_bn(n: int, n_pad: int, in_dtype: int, G: int = 1) -> int:
    """Output-tile width.  Wider tiles re-use the A tile for more columns (the kernel is bound by L2 -> SM operand
    traffic at 128 x 128); 192 divides the ViT widths 384 / 1152 that 256 does not."""
    if n <= 32:
        return 32
    if in_dtype != nv.VT_BF16:
        return 128
    for bn in (256, 192):
        if n % bn == 0 and (G == 1 or n_pad % bn == 0):
            return bn
    return 128


def linear_desc(*, a: torch.Tensor, a_off: int = 0, rows: int, k: int, a_ld: int, w: torch.Tensor, w_off: int = 0,
                n: int, n_pad: int, w_ld: int, out: torch.Tensor, out_off: int = 0, ldc: int, bias=None, bias_off: int = 0,
                act=nv.ACT_NONE, colscale=None, res=None, res_off_elems: int = 0, ldres: int = 0, G: int = 1, a_G: int = 1,
                a_sG: int = 0, out_g: int = 0, res_g: int = 0, bn: Optional[int] = None, out_plane: int = 0, passes: int = 1,
                a_plane: int = 0, w_plane: int = 0, a_C: Optional[int] = None) -> nv.GemmDesc:
    """Plain row-major GEMM out[rows, n] = act(a[rows, k] @ w[n, k]^T + bias) (* colscale) (+ res).

    `k` is the padded reduction length (multiple of 64 bf16 / 32 f32 elements); a's visible width `a_C`
    (default a_ld) bounds what TMA may read, anything beyond reads as zero."""
    in_dtype = VT_DT[a.dtype]
    if bn is None:
        bn = pick_bn(n, n_pad, in_dtype, G)
    return gemm_desc(
        a=ptr(a, a_off), in_dtype=in_dtype, a_C=a_C if a_C is not None else a_ld, a_T=rows, a_B=1, a_ld=a_ld,
        a_sB=a_ld * rows, a_G=a_G, a_sG=a_sG, kc=k, t_box=min(128, rows), b_box=1, w=ptr(w, w_off), n_pad=n_pad, w_ld=w_ld, G=G,
        M=rows, N=n, bn=bn, out=ptr(out, out_off), out_dtype=VT_DT[out.dtype], ldc=ldc, out_g=out_g, row_div=1, out_q=1,
        out_r=0, out_off=0, out_plane=out_plane, act=act, bias=ptr(bias, bias_off), colscale=ptr(colscale),
        res=ptr(res, res_off_elems), ldres=ldres, res_g=res_g, res_q=1, res_r=0, res_off=0, passes=passes,
        a_plane=a_plane, w_plane=w_plane)


def pack_linear_weight(w: torch.Tensor, dtype: torch.dtype, bn: int = 128, k_mult: Optional[int] = None,
                       split: bool = False) -> Tuple[torch.Tensor, int, int]:
    """[N, K] -> zero-padded [n_pad, k_pad] (k_pad doubled with the tf32 lo plane when split).  Returns (w, n_pad, k_pad)."""
    N, K = w.shape
    km = k_mult or (64 if dtype == torch.bfloat16 else 32)
    n_pad, k_pad = round_up(N, bn), round_up(K, km)
    out = torch.zeros(n_pad, k_pad * (2 if split else 1), dtype=torch.float32, device=w.device)
    out[:N, :K] = w.float()
    if split:
        hi = tf32_round(out[:, :k_pad])
        out[:, k_pad:] = out[:, :k_pad] - hi
        out[:, :k_pad] = hi
    return out.to(dtype).contiguous(), n_pad, k_pad


def pad_vec(v: Optional[torch.Tensor], n_pad: int, fill: float = 0.0) -> torch.Tensor:
    out = torch.full((n_pad,), fill, dtype=torch.float32, device=v.device)
    out[: v.numel()] = v.float().reshape(-1)
    return out


def tf32_round(x: torch.Tensor) -> torch.Tensor:
    """Round-to-nearest-even onto tf32 (matches vt::tf32_hi)."""
    i = x.contiguous().view(torch.int32)
    lsb = (i >> 13) & 1
    i = (i + 0x0FFF + lsb) & ~0x1FFF
    return i.view(torch.float32)
