"""ctypes binding of libvt_b200.so (C ABI: include/vt_b200.h).

The library is built in-tree by `vla_touch_b200.build.build()` (nvcc, sm_100a).  There is NO fallback:
if the shared object is missing, or a call is made without a CUDA device, this module raises.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import List, Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
# VT_LIB: developer override (tools/ only), e.g. the instrumented build `python -m vla_touch_b200.build --debug-knobs` writes
LIB_PATH = os.environ.get("VT_LIB") or os.path.join(_HERE, "lib", "libvt_b200.so")

VT_BF16, VT_F32, VT_U8 = 0, 1, 2
ACT_NONE, ACT_GELU, ACT_MISH = 0, 1, 2
EPI_LINEAR, EPI_GN = 0, 1
LAYOUT_BHWC, LAYOUT_BCHW = 0, 1
MAX_TAPS = 8
ABI_VERSION = 6

i32, i64, f32, u64, vp = C.c_int32, C.c_int64, C.c_float, C.c_uint64, C.c_void_p


class GemmDesc(C.Structure):
    _fields_ = [
        ("a", vp), ("in_dtype", i32), ("a_C", i32), ("a_P", i32), ("a_T", i32), ("a_B", i32), ("a_G", i32),
        ("a_ld", i64), ("a_sB", i64), ("a_sG", i64), ("a_c0", i32), ("kc", i32), ("taps", i32),
        ("tap_p", i32 * MAX_TAPS), ("tap_t", i32 * MAX_TAPS), ("t_box", i32), ("b_box", i32), ("passes", i32),
        ("a_plane", i32), ("w_plane", i32),
        ("w", vp), ("n_pad", i32), ("w_ld", i32),
        ("G", i32), ("M", i32), ("N", i32), ("bn", i32),
        ("out", vp), ("out_dtype", i32), ("ldc", i32), ("out_g", i64), ("row_div", i32),
        ("out_q", i64), ("out_r", i64), ("out_off", i64), ("out_plane", i64),
        ("epi", i32), ("act", i32), ("bias", vp), ("colscale", vp), ("res", vp), ("ldres", i32),
        ("res_g", i64), ("res_q", i64), ("res_r", i64), ("res_off", i64), ("res_plane", i64),
        ("gn_gamma", vp), ("gn_beta", vp), ("gn_group_ch", i32), ("gn_eps", f32),
        ("film_c", vp), ("film_t", vp), ("film_g", i64), ("film_tg", i64), ("film_ld", i32), ("film_C", i32), ("film_off", i32),
        ("raw_out", vp), ("raw_g", i64), ("raw_ld", i32),
    ]


class LnDesc(C.Structure):
    _fields_ = [("x", vp), ("in_ld", i64), ("in_row_stride", i64), ("rows", i32), ("D", i32), ("gamma", vp),
                ("beta", vp), ("eps", f32), ("out", vp), ("out_dtype", i32), ("out_ld", i64), ("out_plane", i64),
                ("act", i32)]


class AttnDesc(C.Structure):
    _fields_ = [("qkv", vp), ("ctx", vp), ("in_dtype", i32), ("images", i32), ("tokens", i32), ("heads", i32),
                ("ctx_ld", i64), ("ctx_plane", i64)]


class MlpDesc(C.Structure):
    _fields_ = [("xn", vp), ("ld_x", i64), ("w1", vp), ("w1_ld", i64), ("b1", vp), ("w2", vp), ("w2_ld", i64), ("b2", vp),
                ("ls2", vp), ("h", vp), ("ld_h", i64), ("rows", i32), ("D", i32),
                ("ln_gamma", vp), ("ln_beta", vp), ("ln_out", vp), ("ln_ld", i64), ("ln_eps", C.c_float)]


class RowprojDesc(C.Structure):
    _fields_ = [("x", vp), ("ld_x", i64), ("w", vp), ("w_ld", i64), ("bias", vp), ("colscale", vp), ("h", vp), ("ld_h", i64),
                ("rows", i32), ("D", i32), ("ln_gamma", vp), ("ln_beta", vp), ("ln_out", vp), ("ln_ld", i64), ("ln_eps", C.c_float)]


class ImgStatsDesc(C.Structure):
    _fields_ = [("img", vp), ("dtype", i32), ("count", i64), ("partial", vp), ("flags", vp)]


class PatchifyDesc(C.Structure):
    _fields_ = [("img", vp), ("dtype", i32), ("layout", i32), ("images", i32), ("H", i32), ("W", i32), ("patch", i32),
                ("flags", vp), ("out", vp), ("out_dtype", i32), ("out_cols", i32), ("out_ld", i32), ("out_plane", i64)]


class ClsDesc(C.Structure):
    _fields_ = [("cls", vp), ("pos", vp), ("h", vp), ("images", i32), ("tokens", i32), ("D", i32)]


class PackDesc(C.Structure):
    _fields_ = [("src", vp), ("src_ld", i64), ("rows", i32), ("cols", i32), ("act", i32), ("out", vp),
                ("out_dtype", i32), ("out_ld", i64), ("dst_c0", i32), ("out_plane", i64), ("zero_to", i32),
                ("src_row_div", i32)]


class AffineDesc(C.Structure):
    _fields_ = [("x", vp), ("out", vp), ("mins", vp), ("maxs", vp), ("rows", i32), ("A", i32), ("denorm", i32),
                ("pad", f32), ("xpad", vp), ("xpad_dtype", i32), ("xpad_ld", i32), ("xpad_plane", i64), ("add", vp)]


class TembedDesc(C.Structure):
    _fields_ = [("t", vp), ("rows", i32), ("dim", i32), ("out", vp), ("out_dtype", i32), ("out_ld", i64),
                ("out_plane", i64)]


class SdeDesc(C.Structure):
    _fields_ = [("x", vp), ("v", vp), ("s", vp), ("noise", vp), ("rows", i32), ("A", i32), ("ginv", f32),
                ("dgg", f32), ("eps", f32), ("dt", f32), ("nscale", f32), ("d", f32), ("seed", u64), ("seed_dev", vp), ("step", i32),
                ("xpad", vp), ("xpad_dtype", i32), ("xpad_ld", i32), ("xpad_plane", i64)]


class LstmDesc(C.Structure):
    _fields_ = [("xw", vp), ("w_hh", vp), ("h", vp), ("c", vp), ("y", vp), ("y_dtype", i32), ("y_ld", i64),
                ("y_plane", i64), ("B", i32), ("T", i32), ("H", i32), ("w_hh_tc", vp), ("h_tc", vp), ("zero_init", i32)]


class QsampleDesc(C.Structure):
    _fields_ = [("x0", vp), ("x1", vp), ("step", vp), ("z_unit", vp), ("d", f32), ("B", i32), ("n", i32), ("A", i32),
                ("xt", vp), ("tclip", vp), ("xpad", vp), ("xpad_dtype", i32), ("xpad_ld", i32), ("xpad_plane", i64)]


class SilossDesc(C.Structure):
    _fields_ = [("bvs", vp), ("x0", vp), ("x1", vp), ("z_unit", vp), ("tclip", vp), ("d", f32), ("B", i32), ("n", i32),
                ("per_sample", vp), ("out", vp)]


class TcolDesc(C.Structure):
    _fields_ = [("src", vp), ("src_dtype", i32), ("ld", i64), ("sB", i64), ("sG", i64), ("G", i32), ("B", i32), ("T_src", i32),
                ("C", i32), ("taps", i32), ("tap_off", i32 * 8), ("stride", i32), ("t_out", i32), ("out", vp), ("c_pad", i32),
                ("k_ld", i64), ("out_g", i64)]


class WgradDesc(C.Structure):
    _fields_ = [("rows", vp), ("rows_C", i32), ("rows_P", i32), ("rows_T", i32), ("rows_ld", i64), ("rows_sB", i64), ("rows_sG", i64),
                ("rows_p", i32), ("rows_t", i32), ("cols", vp), ("cols_C", i32), ("cols_P", i32), ("cols_T", i32), ("cols_ld", i64),
                ("cols_sB", i64), ("cols_sG", i64), ("taps", i32), ("tap_p", i32 * MAX_TAPS), ("tap_t", i32 * MAX_TAPS), ("c_pad", i32),
                ("G", i32), ("B", i32), ("t_out", i32), ("R", i32), ("out", vp), ("ldc", i32), ("out_g", i64)]


class GnbwdDesc(C.Structure):
    _fields_ = [("raw", vp), ("dout", vp), ("dout_ld", i64), ("dout_g", i64), ("gamma", vp), ("beta", vp), ("p_ld", i32),
                ("film", vp), ("dfilm", vp), ("film_g", i64), ("film_ld", i32), ("film_off", i32), ("draw", vp), ("part", vp),
                ("dgamma", vp), ("dbeta", vp), ("dbias", vp), ("G", i32), ("B", i32), ("T", i32), ("C", i32), ("groups", i32),
                ("eps", f32)]


class ColsumDesc(C.Structure):
    _fields_ = [("x", vp), ("ld", i64), ("x_g", i64), ("G", i32), ("rows", i32), ("C", i32), ("out", vp), ("out_ld", i32)]


class EwiseDesc(C.Structure):
    _fields_ = [("a", vp), ("a_ld", i64), ("b", vp), ("b_ld", i64), ("out", vp), ("out_ld", i64), ("rows", i64), ("cols", i32),
                ("op", i32), ("alpha", f32)]


EW_ADD, EW_MISH_BWD, EW_GELU_BWD, EW_MUL, EW_SCALED_DIFF = 0, 1, 2, 3, 4


class LnGeluBwdDesc(C.Structure):
    _fields_ = [("z0", vp), ("dzn", vp), ("gamma", vp), ("beta", vp), ("eps", f32), ("dz0", vp), ("d1", vp), ("d1zh", vp),
                ("rows", i32), ("D", i32)]


class SilossBwdDesc(C.Structure):
    _fields_ = [("bvs", vp), ("x0", vp), ("x1", vp), ("z_unit", vp), ("tclip", vp), ("d", f32), ("B", i32), ("n", i32),
                ("dvs", vp)]


class LstmTrainDesc(C.Structure):
    _fields_ = [("xw", vp), ("w_hh", vp), ("y", vp), ("y_dtype", i32), ("y_ld", i64), ("gates", vp), ("c", vp), ("B", i32),
                ("T", i32), ("H", i32), ("w_hh_tc", vp), ("h_tc", vp)]


class LstmBwdDesc(C.Structure):
    _fields_ = [("gates", vp), ("c", vp), ("dy", vp), ("dy_ld", i64), ("w_hh", vp), ("dgates", vp), ("B", i32), ("T", i32),
                ("H", i32), ("w_hh_t_tc", vp), ("dg_tc", vp)]


class DropmaskDesc(C.Structure):
    _fields_ = [("inject", vp), ("p", f32), ("seed", u64), ("seed_dev", vp), ("stream", i32), ("mask", vp), ("n", i64)]


class OptTensor(C.Structure):
    _fields_ = [("p", vp), ("g", vp), ("m", vp), ("v", vp), ("ema", vp), ("numel", i64),
                ("taps", i32), ("c", i32), ("c_pad", i32), ("reserved", i32), ("w_op", vp)]


PERSIST_MAX_DEPS, PERSIST_MAX_LAYERS = 4, 37


class PersistDesc(C.Structure):
    """vt_persist_desc.  The host arrays it points to are kept alive as Python attributes of the instance (`_keep`); `layers`,
    `sde`, `coef` mirror them for inspection and for the CPU plan interpreter."""
    _fields_ = [("gemms", vp), ("n_gemms", i32), ("sde", vp), ("deps", vp), ("dep_lag", vp), ("n_steps", i32),
                ("film_t_step", i64), ("sde_coef", vp), ("noise_step", i64), ("sde_T", i32)]


class AdamwDesc(C.Structure):
    _fields_ = [("tensors", vp), ("chunks", vp), ("n_chunks", i32), ("chunk_elems", i32), ("lr", f32), ("beta1", f32),
                ("beta2", f32), ("eps", f32), ("weight_decay", f32), ("bias_corr1", f32), ("bias_corr2", f32),
                ("ema_decay", f32), ("grad_scale", f32)]


class BatchGatherDesc(C.Structure):
    _fields_ = [("qpos", vp), ("grip_scaled", vp), ("vla", vp), ("vla_last_scaled", vp), ("forces", vp), ("disps", vp), ("feats", vp),
                ("frame_mean", vp), ("start", vp), ("B", i32), ("A", i32), ("vla_T", i32), ("Fd", i32), ("Dd", i32), ("D", i32),
                ("context_frames", i32), ("horizon", i32), ("states", vp), ("expert_actions", vp), ("vla_actions", vp),
                ("forces_out", vp), ("disps_out", vp), ("feat_cam1", vp), ("feat_cam2", vp), ("branch", vp),
                ("action_mins", vp), ("action_maxs", vp), ("vla_mins", vp), ("vla_maxs", vp), ("pad", f32), ("expert_n", vp), ("vla_n", vp)]


EXPORTS = [
    "vt_last_error", "vt_abi_version", "vt_device_info", "vt_program_create", "vt_program_destroy",
    "vt_program_num_ops", "vt_program_num_launches", "vt_program_add_gemm", "vt_program_add_layernorm",
    "vt_program_add_attention", "vt_program_add_mlp", "vt_program_add_rowproj", "vt_debug_timestamps", "vt_debug_persist_trace", "vt_pad_resize_area", "vt_gather_repack", "vt_program_add_imgstats", "vt_program_add_patchify", "vt_program_add_cls",
    "vt_program_add_pack", "vt_program_add_affine", "vt_program_add_tembed", "vt_program_add_sde",
    "vt_program_add_lstm", "vt_program_add_qsample", "vt_program_add_siloss", "vt_program_add_tcol", "vt_program_add_gnbwd", "vt_program_add_colsum", "vt_program_add_ewise", "vt_program_add_silossbwd", "vt_program_add_lstm_train", "vt_program_add_lstm_bwd", "vt_program_add_lngelubwd", "vt_program_add_dropmask", "vt_program_add_persist", "vt_program_add_wgrad", "vt_program_run", "vt_program_graph_build", "vt_program_graph_launch",
    "vt_pos_embed_resize", "vt_adamw_ema_step", "vt_batch_gather", "vt_chunk_handoff",
]

_ADD = {
    GemmDesc: "vt_program_add_gemm", LnDesc: "vt_program_add_layernorm", AttnDesc: "vt_program_add_attention", MlpDesc: "vt_program_add_mlp", RowprojDesc: "vt_program_add_rowproj",
    ImgStatsDesc: "vt_program_add_imgstats", PatchifyDesc: "vt_program_add_patchify", ClsDesc: "vt_program_add_cls",
    PackDesc: "vt_program_add_pack", AffineDesc: "vt_program_add_affine", TembedDesc: "vt_program_add_tembed",
    SdeDesc: "vt_program_add_sde", LstmDesc: "vt_program_add_lstm", QsampleDesc: "vt_program_add_qsample",
    SilossDesc: "vt_program_add_siloss", TcolDesc: "vt_program_add_tcol", GnbwdDesc: "vt_program_add_gnbwd",
    ColsumDesc: "vt_program_add_colsum", EwiseDesc: "vt_program_add_ewise",
    SilossBwdDesc: "vt_program_add_silossbwd", LstmTrainDesc: "vt_program_add_lstm_train", LstmBwdDesc: "vt_program_add_lstm_bwd",
    LnGeluBwdDesc: "vt_program_add_lngelubwd", DropmaskDesc: "vt_program_add_dropmask", PersistDesc: "vt_program_add_persist",
    WgradDesc: "vt_program_add_wgrad",
}

_lib: Optional[C.CDLL] = None


class NativeError(RuntimeError):
    pass


def lib() -> C.CDLL:
    """Load libvt_b200.so (once).  Raises if it has not been built: there is no CPU/PyTorch fallback."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NativeError(f"{LIB_PATH} not found: build it with `python -m vla_touch_b200.build` "
                              "(nvcc, sm_100a). vla_touch_b200 has no fallback path.")
        L = C.CDLL(LIB_PATH)
        L.vt_last_error.restype = C.c_char_p
        L.vt_program_create.argtypes = [C.POINTER(vp)]
        L.vt_program_destroy.argtypes = [vp]
        L.vt_program_num_ops.argtypes = [vp]
        L.vt_program_num_launches.argtypes = [vp, C.c_int, C.c_int]
        for desc, name in _ADD.items():
            getattr(L, name).argtypes = [vp, C.POINTER(desc)]
        L.vt_program_run.argtypes = [vp, C.c_int, C.c_int, vp]
        L.vt_program_graph_build.argtypes = [vp, C.c_int, C.c_int]
        L.vt_program_graph_launch.argtypes = [vp, vp]
        L.vt_device_info.argtypes = [C.POINTER(i32), C.POINTER(i32), C.POINTER(i32)]
        L.vt_pos_embed_resize.argtypes = [vp, i32, vp, i32, i32, i32, vp]
        L.vt_adamw_ema_step.argtypes = [C.POINTER(AdamwDesc), vp]
        L.vt_pad_resize_area.argtypes = [vp, i32, i32, i32, i32, vp, i32, vp]
        L.vt_gather_repack.argtypes = [vp, vp, i32, vp, vp]
        L.vt_batch_gather.argtypes = [C.POINTER(BatchGatherDesc), vp]
        L.vt_chunk_handoff.argtypes = [vp, i32, i32, i32, i32, vp, vp, i32, f32, vp, vp, i32, vp]
        if L.vt_abi_version() != ABI_VERSION:
            raise NativeError("libvt_b200.so ABI version mismatch; rebuild it")
        _lib = L
    return _lib


def check(rc: int) -> None:
    if rc != 0:
        raise NativeError(f"libvt_b200 error {rc}: {lib().vt_last_error().decode()}")


def device_info():
    sm, ma, mi = i32(), i32(), i32()
    check(lib().vt_device_info(C.byref(sm), C.byref(ma), C.byref(mi)))
    return sm.value, ma.value, mi.value


def require_b200() -> None:
    """The product path runs on sm_100 only; anything else is an error, never a fallback."""
    sm, ma, mi = device_info()
    if ma != 10:
        raise NativeError(f"vla_touch_b200 needs an sm_100a (B200) device, found compute capability {ma}.{mi}")


class nvtx_range:
    """`with nvtx_range("vt.predict"):` -- an NVTX range around a host-side phase when VT_NVTX=1 (ncu --nvtx / any NVTX-aware profiler
    can then filter the launches of one phase); a no-op otherwise (the reference has no tracing of its own)."""
    _on = None

    def __init__(self, name: str):
        self.name = name

    def __enter__(self):
        if nvtx_range._on is None:
            nvtx_range._on = os.environ.get("VT_NVTX") == "1"
        if nvtx_range._on:
            import torch
            torch.cuda.nvtx.range_push(self.name)
        return self

    def __exit__(self, *exc):
        if nvtx_range._on:
            import torch
            torch.cuda.nvtx.range_pop()
        return False


def current_stream_ptr() -> int:
    import torch
    return torch.cuda.current_stream().cuda_stream


class Program:
    """An ordered list of pre-encoded kernel launches living in the native library."""

    def __init__(self) -> None:
        h = vp()
        check(lib().vt_program_create(C.byref(h)))
        self._h = h
        self.descs: List[C.Structure] = []   # python-side mirror (kept for inspection and the CPU plan emulator)
        self.keep: list = []                 # tensors that must outlive the program

    def add(self, desc: C.Structure) -> int:
        check(getattr(lib(), _ADD[type(desc)])(self._h, C.byref(desc)))
        self.descs.append(desc)
        return len(self.descs) - 1

    def __len__(self) -> int:
        return len(self.descs)

    def num_launches(self, first: int = 0, count: int = -1) -> int:
        return lib().vt_program_num_launches(self._h, first, count)

    def run(self, first: int = 0, count: int = -1, stream: Optional[int] = None) -> None:
        check(lib().vt_program_run(self._h, first, count, vp(current_stream_ptr() if stream is None else stream)))

    def graph_build(self, first: int = 0, count: int = -1) -> None:
        check(lib().vt_program_graph_build(self._h, first, count))

    def graph_launch(self, stream: Optional[int] = None) -> None:
        check(lib().vt_program_graph_launch(self._h, vp(current_stream_ptr() if stream is None else stream)))

    def __del__(self) -> None:
        try:
            if self._h:
                lib().vt_program_destroy(self._h)
                self._h = None
        except Exception:
            pass
