// Host side of libvt_b200.so: validates op descriptors, encodes TMA tensor maps once, keeps programs (ordered
// lists of ready-to-launch kernels) and replays them on a stream or as a CUDA graph.  C ABI in include/vt_b200.h.
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <algorithm>
#include <cmath>
#include <memory>
#include <string>
#include <vector>

#include "../../include/vt_b200.h"
#include "vt_attn.cuh"
#include "vt_mlp.cuh"
#include "vt_rowproj.cuh"
#include "vt_elem.cuh"
#include "vt_gemm.cuh"
#include "vt_persist.cuh"
#include "vt_attn_pp.cuh"
#include "vt_resize.cuh"
#include "vt_dataset.cuh"
#include "vt_wgrad.cuh"
#include "vt_lstm.cuh"
#include "vt_lstm_tc.cuh"
#include "vt_bwd.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof(buf), fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define VT_CUDA(expr)                                                                         \
  do {                                                                                        \
    cudaError_t e__ = (expr);                                                                 \
    if (e__ != cudaSuccess) return fail(VT_E_CUDA, "%s: %s", #expr, cudaGetErrorString(e__)); \
  } while (0)

#define VT_REQUIRE(cond, ...) \
  do {                        \
    if (!(cond)) return fail(VT_E_INVALID, __VA_ARGS__); \
  } while (0)

// --------------------------------------------------------------------------------------------
// TMA tensor maps (driver entry point fetched at run time: no link-time dependency on libcuda)
// --------------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

int make_tmap(CUtensorMap* out, int dtype, int rank, const void* base, const uint64_t* dims, const uint64_t* strides_bytes,
              const uint32_t* box) {
  EncodeTiledFn fn = encode_fn();
  if (!fn) return fail(VT_E_CUDA, "cuTensorMapEncodeTiled is unavailable (no CUDA driver?)");
  VT_REQUIRE((reinterpret_cast<uintptr_t>(base) & 15) == 0, "TMA base pointer %p is not 16-byte aligned", base);
  cuuint64_t gd[5], gs[4];
  cuuint32_t bx[5], es[5];
  for (int i = 0; i < rank; ++i) {
    gd[i] = dims[i];
    bx[i] = box[i];
    es[i] = 1;
    VT_REQUIRE(dims[i] >= 1 && dims[i] <= (1ull << 32), "TMA dim %d extent %llu out of range", i, (unsigned long long)dims[i]);
    VT_REQUIRE(box[i] >= 1 && box[i] <= 256, "TMA box %d extent %u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gs[i] = strides_bytes[i];
    VT_REQUIRE((gs[i] & 15) == 0 && gs[i] < (1ull << 40), "TMA stride %d = %llu bytes must be a multiple of 16", i,
               (unsigned long long)gs[i]);
  }
  const CUtensorMapDataType dt = dtype == VT_BF16 ? CU_TENSOR_MAP_DATA_TYPE_BFLOAT16 : CU_TENSOR_MAP_DATA_TYPE_FLOAT32;
  CUresult r = fn(out, dt, (cuuint32_t)rank, const_cast<void*>(base), gd, gs, bx, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(VT_E_CUDA, "cuTensorMapEncodeTiled failed with CUresult %d", (int)r);
  return VT_OK;
}

int grid_for(long long total, int per_block) {
  long long b = (total + per_block - 1) / per_block;
  if (b < 1) b = 1;
  if (b > 148 * 16) b = 148 * 16;
  return (int)b;
}

// --------------------------------------------------------------------------------------------
// ops
// --------------------------------------------------------------------------------------------
struct Op {
  virtual ~Op() {}
  virtual int launch(cudaStream_t s) = 0;
  virtual int launches() const { return 1; }
};

#define VT_LAUNCH_CHECK(name)                                                                \
  do {                                                                                       \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) return fail(VT_E_CUDA, "%s launch: %s", name, cudaGetErrorString(e__)); \
  } while (0)

// Launch with optional cluster-of-two and programmatic-dependent-launch attributes.  PDL (env VT_PDL=0 disables it): the
// kernel may be scheduled while its stream predecessor drains; every kernel launched this way calls griddepcontrol.wait
// before it touches global memory, so the data dependences are unchanged -- only launch latency and prologue are hidden.
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("VT_PDL");
    v = (e && atoi(e) == 0) ? 0 : 1;
  }
  return v == 1;
}
template <typename... KArgs, typename... Args>
cudaError_t launch_ex(void (*fn)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t s, bool cluster2, bool pdl, Args... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = s;
  cudaLaunchAttribute at[2];
  int n = 0;
  if (cluster2) {
    at[n].id = cudaLaunchAttributeClusterDimension;
    at[n].val.clusterDim.x = 2;
    at[n].val.clusterDim.y = 1;
    at[n].val.clusterDim.z = 1;
    ++n;
  }
  if (pdl && pdl_enabled()) {
    at[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = at;
  cfg.numAttrs = (unsigned)n;
  return cudaLaunchKernelEx(&cfg, fn, KArgs(args)...);
}

// ---- GEMM ----
typedef void (*GemmKernel)(const vt::GemmArgs);
struct GemmVariant {
  GemmKernel fn;
  int smem;
  int ctas;   // 1, or 2 = CTA pairs (cluster of two, tcgen05 cta_group::2, M = 256 per MMA)
  int threads;
  bool attr_set;
};

template <typename TIn, int BN, int MODE, typename TOut, int STAGES, int KA, int CTAS = 1, int EW = 8>
GemmVariant make_variant() {
  GemmVariant v;
  v.fn = vt::gemm_tc_kernel<TIn, BN, MODE, TOut, STAGES, (sizeof(TIn) == 4), KA, CTAS, EW>;
  v.smem = vt::gemm_smem_bytes<BN, STAGES, MODE, KA, CTAS, EW>();
  static_assert(vt::gemm_smem_bytes<BN, STAGES, MODE, KA, CTAS, EW>() <= 227 * 1024, "shared memory budget");
  v.ctas = CTAS;
  v.threads = vt::GEMM_THREADS(EW);
  v.attr_set = false;
  return v;
}

// (stages x K atoms per stage) per tile width: ~190 KB of operand ring, two atoms per barrier round.
// `pair` selects the CTA-pair kernels (bf16, bn 192 / 256): each CTA stages its 128 rows of A and HALF of the B tile, so
// the operand bytes per MMA cycle that cross L2 -> shared memory halve against the single-CTA 128 x bn tile.
GemmVariant* gemm_variant(int in_dtype, int bn, int epi, int out_dtype, bool pair) {
  using bf = __nv_bfloat16;
  static GemmVariant v_b_256_l_b = make_variant<bf, 256, vt::EPI_LINEAR, bf, 3, 1>();
  static GemmVariant v_b_256_l_f = make_variant<bf, 256, vt::EPI_LINEAR, float, 3, 1>();
  static GemmVariant v_b_192_l_b = make_variant<bf, 192, vt::EPI_LINEAR, bf, 2, 2>();
  static GemmVariant v_b_192_l_f = make_variant<bf, 192, vt::EPI_LINEAR, float, 2, 2>();
  static GemmVariant v_b_128_l_b = make_variant<bf, 128, vt::EPI_LINEAR, bf, 5, 1>();
  static GemmVariant v_b_128_l_f = make_variant<bf, 128, vt::EPI_LINEAR, float, 5, 1>();
  static GemmVariant v_b_32_l_b = make_variant<bf, 32, vt::EPI_LINEAR, bf, 4, 2>();
  static GemmVariant v_b_32_l_f = make_variant<bf, 32, vt::EPI_LINEAR, float, 4, 2>();
  static GemmVariant v_b_128_g_b = make_variant<bf, 128, vt::EPI_GN, bf, 3, 2>();
  static GemmVariant v_f_128_l_f = make_variant<float, 128, vt::EPI_LINEAR, float, 5, 1>();
  static GemmVariant v_f_32_l_f = make_variant<float, 32, vt::EPI_LINEAR, float, 4, 2>();
  static GemmVariant v_f_128_g_f = make_variant<float, 128, vt::EPI_GN, float, 3, 2>();
  static GemmVariant p_b_256_l_b = make_variant<bf, 256, vt::EPI_LINEAR, bf, 5, 1, 2>();
  static GemmVariant p_b_256_l_f = make_variant<bf, 256, vt::EPI_LINEAR, float, 5, 1, 2>();
  static GemmVariant p_b_192_l_b = make_variant<bf, 192, vt::EPI_LINEAR, bf, 3, 2, 2>();
  static GemmVariant p_b_192_l_f = make_variant<bf, 192, vt::EPI_LINEAR, float, 3, 2, 2>();
  static GemmVariant p_b_256_g_b = make_variant<bf, 256, vt::EPI_GN, bf, 3, 2, 2>();
  static GemmVariant p_b_128_g_b = make_variant<bf, 128, vt::EPI_GN, bf, 4, 2, 2>();
  if (in_dtype == VT_BF16) {
    if (epi == VT_EPI_GN) {
      if (out_dtype != VT_BF16) return nullptr;
      if (bn == 256) return pair ? &p_b_256_g_b : nullptr;
      if (bn == 128) return pair ? &p_b_128_g_b : &v_b_128_g_b;
      return nullptr;
    }
    if (pair && bn == 256) return out_dtype == VT_BF16 ? &p_b_256_l_b : &p_b_256_l_f;
    if (pair && bn == 192) return out_dtype == VT_BF16 ? &p_b_192_l_b : &p_b_192_l_f;
    if (bn == 256) return out_dtype == VT_BF16 ? &v_b_256_l_b : &v_b_256_l_f;
    if (bn == 192) return out_dtype == VT_BF16 ? &v_b_192_l_b : &v_b_192_l_f;
    if (bn == 128) return out_dtype == VT_BF16 ? &v_b_128_l_b : &v_b_128_l_f;
    if (bn == 32) return out_dtype == VT_BF16 ? &v_b_32_l_b : &v_b_32_l_f;
    return nullptr;
  }
  if (out_dtype != VT_F32) return nullptr;
  if (epi == VT_EPI_GN) return bn == 128 ? &v_f_128_g_f : nullptr;
  if (bn == 128) return &v_f_128_l_f;
  if (bn == 32) return &v_f_32_l_f;
  return nullptr;
}

int sm_count() {
  static int n = 0;
  if (!n) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0)
      n = 148;
  }
  return n;
}

struct GemmOp : Op {
  vt::GemmArgs args;
  GemmVariant* var;
  dim3 grid;
  int launch(cudaStream_t s) override {
    if (!var->attr_set) {
      VT_CUDA(cudaFuncSetAttribute(var->fn, cudaFuncAttributeMaxDynamicSharedMemorySize, var->smem));
      var->attr_set = true;
    }
    {
      cudaError_t e = launch_ex(var->fn, grid, dim3((unsigned)var->threads, 1, 1), (size_t)var->smem, s, var->ctas == 2, true, args);
      if (e != cudaSuccess) return fail(VT_E_CUDA, "gemm_tc_kernel launch: %s", cudaGetErrorString(e));
    }
    VT_LAUNCH_CHECK("gemm_tc_kernel");
    return VT_OK;
  }
};

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int fill_gemm_args(const vt_gemm_desc& d, int ctas, vt::GemmArgs* out);

int validate_gemm(const vt_gemm_desc& d) {
  VT_REQUIRE(d.in_dtype == VT_BF16 || d.in_dtype == VT_F32, "gemm: in_dtype %d", d.in_dtype);
  VT_REQUIRE(d.out_dtype == VT_BF16 || d.out_dtype == VT_F32, "gemm: out_dtype %d", d.out_dtype);
  const int es = d.in_dtype == VT_BF16 ? 2 : 4;
  const int KE = 128 / es;
  VT_REQUIRE(d.a && d.w && d.out, "gemm: null operand pointer");
  VT_REQUIRE(d.kc > 0 && d.kc % KE == 0, "gemm: kc=%d must be a positive multiple of %d", d.kc, KE);
  VT_REQUIRE(d.taps >= 1 && d.taps <= VT_MAX_TAPS, "gemm: taps=%d", d.taps);
  VT_REQUIRE(d.passes == 1 || (d.passes == 3 && d.in_dtype == VT_F32), "gemm: passes=%d", d.passes);
  VT_REQUIRE(d.t_box >= 1 && d.b_box >= 1 && d.t_box * d.b_box <= 128, "gemm: tile box %dx%d", d.t_box, d.b_box);
  VT_REQUIRE(d.G >= 1 && (d.a_G == 1 || d.a_G == d.G), "gemm: G=%d a_G=%d", d.G, d.a_G);
  VT_REQUIRE(d.M >= 1 && d.N >= 1 && d.N <= d.n_pad, "gemm: M=%d N=%d n_pad=%d", d.M, d.N, d.n_pad);
  VT_REQUIRE(d.bn == 32 || d.bn == 128 || d.bn == 192 || d.bn == 256, "gemm: bn=%d", d.bn);
  // tiles must not straddle groups; with one group the last tile may be ragged (TMA zero-fills, columns >= N are masked)
  VT_REQUIRE(d.G == 1 || d.n_pad % d.bn == 0, "gemm: n_pad=%d must be a multiple of bn=%d when G > 1", d.n_pad, d.bn);
  VT_REQUIRE(d.w_ld >= d.taps * d.kc, "gemm: w_ld=%d < taps*kc=%d", d.w_ld, d.taps * d.kc);
  VT_REQUIRE(d.a_P >= 1 && d.a_T >= 1 && d.a_B >= 1 && d.a_C >= 1, "gemm: bad A extents");
  VT_REQUIRE(d.row_div >= 1, "gemm: row_div=%d", d.row_div);
  return VT_OK;
}

int build_gemm(const vt_gemm_desc& d, GemmOp* op) {
  int rc0 = validate_gemm(d);
  if (rc0) return rc0;
  // CTA pairs whenever the shape has at least two row tiles (env VT_GEMM_PAIR=0 keeps the single-CTA kernels: A/B runs)
  const int m_tiles_1 = (d.M + d.t_box * d.b_box - 1) / (d.t_box * d.b_box);
  const bool pair_enabled = !(getenv("VT_GEMM_PAIR") && atoi(getenv("VT_GEMM_PAIR")) == 0);
  bool pair = pair_enabled && d.in_dtype == VT_BF16 && (d.bn == 256 || d.bn == 192 || (d.bn == 128 && d.epi == VT_EPI_GN)) &&
              m_tiles_1 >= 2 && d.passes == 1;
  if (d.epi == VT_EPI_GN && d.bn == 256) pair = true;   // the 256-wide GroupNorm epilogue exists as a pair kernel only
  GemmVariant* var = gemm_variant(d.in_dtype, d.bn, d.epi, d.out_dtype, pair);
  if (!var) return fail(VT_E_UNSUPPORTED, "gemm: no kernel for in=%d bn=%d epi=%d out=%d", d.in_dtype, d.bn, d.epi, d.out_dtype);
  op->var = var;
  int rc = fill_gemm_args(d, var->ctas, &op->args);
  if (rc) return rc;
  const int workers = sm_count() / var->ctas;
  const long long total = op->args.total_tiles;
  op->grid = dim3((unsigned)(total < workers ? total : workers) * var->ctas, 1u, 1u);   // persistent: one CTA per SM
  return VT_OK;
}

// Kernel arguments of one implicit GEMM for tiles of `ctas` x 128 rows (2 = CTA pairs: the B tensor map's box holds half the
// tile's columns).  Shared by gemm_tc_kernel and the persistent multi-layer kernel.
int fill_gemm_args(const vt_gemm_desc& d, int ctas, vt::GemmArgs* out) {
  const int es = d.in_dtype == VT_BF16 ? 2 : 4;
  const int KE = 128 / es;
  vt::GemmArgs& a = *out;
  memset(&a, 0, sizeof(a));
  const int rows_valid = d.t_box * d.b_box;
  const int t_out = d.M / d.a_B;
  if (d.b_box == 1 && d.a_B == 1) {  // one sample: tiles step along its positions
    a.m_t_step = d.t_box;
    a.m_b_step = 0;
  } else {
    VT_REQUIRE(d.t_box == t_out && d.M == t_out * d.a_B, "gemm: t_box=%d must equal positions per sample %d", d.t_box, t_out);
    a.m_t_step = 0;
    a.m_b_step = d.b_box;
  }
  {
    const uint64_t dims[5] = {(uint64_t)d.a_C, (uint64_t)d.a_P, (uint64_t)d.a_T, (uint64_t)d.a_B, (uint64_t)d.a_G};
    const uint64_t sG = d.a_G > 1 ? (uint64_t)d.a_sG : (uint64_t)d.a_sB * d.a_B;
    const uint64_t st[4] = {(uint64_t)d.a_ld * es, (uint64_t)d.a_ld * d.a_P * es, (uint64_t)d.a_sB * es, sG * es};
    const uint32_t box[5] = {(uint32_t)KE, 1u, (uint32_t)d.t_box, (uint32_t)d.b_box, 1u};
    int rc = make_tmap(&a.tmA, d.in_dtype, 5, d.a, dims, st, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d.w_ld, (uint64_t)d.G * d.n_pad};
    const uint64_t st[1] = {(uint64_t)d.w_ld * es};
    const uint32_t box[2] = {(uint32_t)KE, (uint32_t)(d.bn / ctas)};   // a pair splits the tile's N rows of B
    int rc = make_tmap(&a.tmB, d.in_dtype, 2, d.w, dims, st, box);
    if (rc) return rc;
  }
  {
    const char* dbg = VT_DEBUG_KNOBS ? getenv("VT_GEMM_DEBUG") : nullptr;
    a.debug = dbg ? atoi(dbg) : 0;
  }
  a.passes = d.passes;
  a.taps = d.taps;
  a.cblocks = d.kc / KE;
  a.a_c0 = d.a_c0;
  a.a_plane = d.a_plane;
  a.b_plane = d.w_plane;
  a.a_g_mul = d.a_G > 1 ? 1 : 0;
  for (int i = 0; i < d.taps; ++i) {
    VT_REQUIRE(d.tap_p[i] >= 0 && d.tap_p[i] < d.a_P, "gemm: tap %d phase %d", i, d.tap_p[i]);
    a.tap_p[i] = d.tap_p[i];
    a.tap_t[i] = d.tap_t[i];
  }
  a.rows_valid = rows_valid;
  a.a_box_bytes = rows_valid * 128;
  a.n_pad = d.n_pad;
  a.M_total = d.M;
  a.N = d.N;
  a.row_div = d.row_div;
  a.out_q = d.out_q;
  a.out_r = d.out_r;
  a.out_off = d.out_off;
  a.out_g = d.out_g;
  a.ldc = d.ldc;
  a.out_plane = d.out_plane;
  a.out = d.out;
  a.bias = d.bias;
  a.act = d.act;
  a.colscale = d.colscale;
  a.res = d.res;
  a.res_q = d.res_q;
  a.res_r = d.res_r;
  a.res_off = d.res_off;
  a.res_g = d.res_g;
  a.res_plane = d.res_plane;
  a.ldres = d.ldres;
  const int oes = d.out_dtype == VT_BF16 ? 2 : 4;
  const int ov = 16 / oes;
  bool vec = aligned16(d.out) && d.ldc % ov == 0 && d.out_g % ov == 0 && d.out_plane % ov == 0;
  if (d.res) {
    const int res_es = (d.epi == VT_EPI_GN) ? oes : 4;
    const int rv = 16 / res_es;
    vec = vec && aligned16(d.res) && d.ldres % rv == 0 && d.res_g % rv == 0;
  }
  a.vec = vec ? 1 : 0;
  {  // coalescing epilogue: full-width tiles, plain (no hi|lo plane) output, row offsets that fit 32 bits
    auto span = [&](long long q, long long r, long long off, long long ld) {
      return (((long long)d.M / d.row_div + 1) * (q < 0 ? -q : q) + (long long)d.row_div * (r < 0 ? -r : r) + off + 1) * ld;
    };
    const bool combo = d.act == VT_ACT_NONE || (d.act == VT_ACT_GELU && !d.res);   // instantiated (activation, residual) pairs
    bool fast = vec && combo && d.epi == VT_EPI_LINEAR && d.bn >= 128 && d.N % d.bn == 0 && d.out_plane == 0 && d.res_plane == 0 &&
                d.out_q >= 0 && d.out_r >= 0 && d.out_off >= 0 && span(d.out_q, d.out_r, d.out_off, d.ldc) < (1ll << 31);
    if (d.res) fast = fast && d.res_q >= 0 && d.res_r >= 0 && d.res_off >= 0 && span(d.res_q, d.res_r, d.res_off, d.ldres) < (1ll << 31);
    const char* nf = getenv("VT_GEMM_FAST");
    if (nf && atoi(nf) == 0) fast = false;
    a.fast = fast ? 1 : 0;
  }
  if (d.epi == VT_EPI_GN) {
    VT_REQUIRE(d.gn_gamma && d.gn_beta, "gemm: GroupNorm epilogue needs gamma/beta");
    VT_REQUIRE(d.gn_group_ch == 32 || d.gn_group_ch == 64, "gemm: gn_group_ch=%d", d.gn_group_ch);
    VT_REQUIRE(d.N % d.bn == 0, "gemm: GroupNorm epilogue needs N %% bn == 0 (N=%d, bn=%d)", d.N, d.bn);
    VT_REQUIRE(d.t_box == t_out, "gemm: GroupNorm epilogue needs whole samples per tile (t_box=%d, positions=%d)", d.t_box, t_out);
    VT_REQUIRE(d.row_div == d.t_box, "gemm: GroupNorm epilogue needs row_div == t_box");
    VT_REQUIRE(vec, "gemm: GroupNorm epilogue needs 16-byte aligned rows");
    VT_REQUIRE(d.b_box <= 32, "gemm: GroupNorm epilogue supports at most 32 samples per tile");
    if (d.film_c) VT_REQUIRE(d.film_ld % 4 == 0 && d.film_off % 4 == 0 && d.film_C % 4 == 0 && d.film_g % 4 == 0 && aligned16(d.film_c) &&
                                 (!d.film_t || (aligned16(d.film_t) && d.film_tg % 4 == 0)),
                             "gemm: FiLM tables must be 16-byte aligned");
    a.gn_gamma = d.gn_gamma;
    a.gn_beta = d.gn_beta;
    a.gn_gs_log2 = d.gn_group_ch == 32 ? 5 : 6;
    a.gn_rows = d.t_box;
    a.gn_eps = d.gn_eps;
    a.film_c = d.film_c;
    a.film_t = d.film_t;
    a.film_g = d.film_g;
    a.film_tg = d.film_tg;
    a.film_ld = d.film_ld;
    a.film_C = d.film_C;
    a.film_off = d.film_off;
    if (d.raw_out) {
      VT_REQUIRE(d.in_dtype == VT_BF16 && d.out_dtype == VT_BF16 && d.bn == 256 && d.out_plane == 0 && d.res_plane == 0,
                 "gemm: raw_out is written by the bf16 GroupNorm epilogue at bn = 256 only");
      VT_REQUIRE(d.raw_ld >= d.N && d.raw_ld % 8 == 0 && d.raw_g % 8 == 0 && ((uintptr_t)d.raw_out & 31) == 0, "gemm: raw_out rows must be 32-byte aligned");
      a.raw = d.raw_out;
      a.raw_g = d.raw_g;
      a.raw_ld = d.raw_ld;
    }
  } else {
    VT_REQUIRE(!d.raw_out, "gemm: raw_out needs the GroupNorm epilogue");
  }
  const int m_tiles = (d.M + rows_valid - 1) / rows_valid;
  const int n_tiles = (d.N + d.bn - 1) / d.bn;
  // work units: one 128-row tile per CTA, or two vertically adjacent tiles per CTA pair (a missing second tile is all
  // out-of-bounds: zero-filled by TMA, masked in the epilogue)
  const int m_units = (m_tiles + ctas - 1) / ctas;
  const long long total = (long long)m_units * n_tiles * d.G;
  VT_REQUIRE(total < (1ll << 31), "gemm: too many tiles");
  a.n_tiles = n_tiles;
  a.m_tiles = m_units;
  a.total_tiles = (int)total;
  return VT_OK;
}

// ---- attention ----
struct AttnOp : Op {
  vt::AttnArgs args;
  vt::AttnRowArgs rargs;
  bool row_kernel = false;   // whole key range resident in TMEM (256..272 tokens): attn_row_kernel
  bool pp_kernel = false;    // 257 tokens: attn_pp_kernel (both query tiles of a unit in flight, P in tensor memory)
  vt_attn_desc d;
  dim3 grid;
  int tail_first = -1;   // first query row handled by attn_tail_kernel, or -1
  static bool attr_set;
  int launches() const override { return (d.in_dtype == VT_BF16 && !row_kernel && tail_first >= 0 && grid.x > 0) ? 2 : 1; }
  int launch(cudaStream_t s) override {
    if (d.in_dtype == VT_BF16 && row_kernel && pp_kernel) {
      static bool pp_attr_set = false;
      if (!pp_attr_set) {
        VT_CUDA(cudaFuncSetAttribute(vt::attn_pp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vt::APP_SMEM_BYTES));
        pp_attr_set = true;
      }
      {
        cudaError_t e = launch_ex(vt::attn_pp_kernel, grid, dim3(vt::APP_THREADS, 1, 1), (size_t)vt::APP_SMEM_BYTES, s, false, true, rargs);
        if (e != cudaSuccess) return fail(VT_E_CUDA, "attn_pp_kernel launch: %s", cudaGetErrorString(e));
      }
      VT_LAUNCH_CHECK("attn_pp_kernel");
    } else if (d.in_dtype == VT_BF16 && row_kernel) {
      static bool row_attr_set = false;
      if (!row_attr_set) {
        VT_CUDA(cudaFuncSetAttribute(vt::attn_row_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vt::ATR_SMEM_BYTES));
        row_attr_set = true;
      }
      {
        cudaError_t e = launch_ex(vt::attn_row_kernel, grid, dim3(vt::ATR_THREADS, 1, 1), (size_t)vt::ATR_SMEM_BYTES, s, false, true, rargs);
        if (e != cudaSuccess) return fail(VT_E_CUDA, "attn_row_kernel launch: %s", cudaGetErrorString(e));
      }
      VT_LAUNCH_CHECK("attn_row_kernel");
    } else if (d.in_dtype == VT_BF16) {
      if (!attr_set) {
        VT_CUDA(cudaFuncSetAttribute(vt::attn_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vt::ATT_SMEM_BYTES));
        attr_set = true;
      }
      if (grid.x > 0) {
        vt::attn_tc_kernel<<<grid, vt::ATT_THREADS, vt::ATT_SMEM_BYTES, s>>>(args);
        VT_LAUNCH_CHECK("attn_tc_kernel");
      }
      if (tail_first >= 0) {
        const long long blocks = (long long)d.images * d.heads * (d.tokens - tail_first);
        const size_t smem = (size_t)(d.tokens + 8 + 256) * sizeof(float);
        vt::attn_tail_kernel<<<(unsigned)blocks, vt::ATTT_THREADS, smem, s>>>(reinterpret_cast<const __nv_bfloat16*>(d.qkv),
                                                                            args.ctx, d.images, d.tokens, d.heads, args.D,
                                                                            args.ctx_ld, tail_first, args.scale_log2);
        VT_LAUNCH_CHECK("attn_tail_kernel");
      }
    } else {
      const long long warps = (long long)d.images * d.heads * d.tokens;
      const int blocks = (int)((warps + vt::ATTF_WARPS - 1) / vt::ATTF_WARPS);
      const size_t smem = (size_t)vt::ATTF_WARPS * d.tokens * sizeof(float);
      vt::attn_f32_kernel<<<blocks, vt::ATTF_WARPS * 32, smem, s>>>(reinterpret_cast<const float*>(d.qkv),
                                                                    reinterpret_cast<float*>(d.ctx), d.images, d.tokens,
                                                                    d.heads, d.heads * 64, d.ctx_ld, d.ctx_plane);
      VT_LAUNCH_CHECK("attn_f32_kernel");
    }
    return VT_OK;
  }
};
bool AttnOp::attr_set = false;

// ---- fused ViT MLP ----
struct MlpOp : Op {
  vt::MlpArgs args;
  dim3 grid;
  int launch(cudaStream_t s) override {
    static bool attr_set = false;
    if (!attr_set) {
      VT_CUDA(cudaFuncSetAttribute(vt::mlp_fused_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vt::MLP_SMEM_BYTES));
      attr_set = true;
    }
    cudaError_t e = launch_ex(vt::mlp_fused_kernel, grid, dim3(vt::MLP_THREADS, 1, 1), (size_t)vt::MLP_SMEM_BYTES, s, true, true, args);
    if (e != cudaSuccess) return fail(VT_E_CUDA, "mlp_fused_kernel launch: %s", cudaGetErrorString(e));
    VT_LAUNCH_CHECK("mlp_fused_kernel");
    return VT_OK;
  }
};

// ---- whole-row output projection + LayerNorm ----
struct RowprojOp : Op {
  vt::RowprojArgs args;
  dim3 grid;
  int launch(cudaStream_t s) override {
    static bool attr_set = false;
    if (!attr_set) {
      VT_CUDA(cudaFuncSetAttribute(vt::rowproj_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vt::RP_SMEM_BYTES));
      attr_set = true;
    }
    cudaError_t e = launch_ex(vt::rowproj_kernel, grid, dim3(vt::RP_THREADS, 1, 1), (size_t)vt::RP_SMEM_BYTES, s, true, true, args);
    if (e != cudaSuccess) return fail(VT_E_CUDA, "rowproj_kernel launch: %s", cudaGetErrorString(e));
    VT_LAUNCH_CHECK("rowproj_kernel");
    return VT_OK;
  }
};

// ---- simple ops ----
struct LnOp : Op {
  vt_ln_desc d;
  int launch(cudaStream_t s) override {
    const int blocks = (d.rows + 7) / 8;
#define VT_LN(V)                                                                                                   \
  launch_ex(vt::layernorm_kernel<V>, dim3(blocks), dim3(256), 0, s, false, true, d.x, d.in_ld, d.in_row_stride, d.rows, d.gamma, \
            d.beta, d.eps, d.out, d.out_dtype, d.out_ld, d.out_plane, d.act)
    switch (d.D) {
      case 256: VT_LN(2); break;
      case 384: VT_LN(3); break;
      case 768: VT_LN(6); break;
      case 1024: VT_LN(8); break;
      default: return fail(VT_E_UNSUPPORTED, "layernorm: D=%d", d.D);
    }
#undef VT_LN
    VT_LAUNCH_CHECK("layernorm_kernel");
    return VT_OK;
  }
};

struct ImgStatsOp : Op {
  vt_imgstats_desc d;
  int launches() const override { return 2; }
  int launch(cudaStream_t s) override {
    const int es = d.dtype == VT_U8 ? 1 : 4;
    const long long nvec = d.count / (16 / es);
    int blocks = (int)((nvec + 256 * 8 - 1) / (256 * 8));
    blocks = blocks < 1 ? 1 : (blocks > vt::STATS_MAX_BLOCKS ? vt::STATS_MAX_BLOCKS : blocks);
    float* pmax = d.partial;
    double* psum = reinterpret_cast<double*>(d.partial + vt::STATS_MAX_BLOCKS);
    if (d.dtype == VT_U8)
      vt::imgstats_partial_kernel<uint8_t><<<blocks, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(d.img), d.count, pmax, psum);
    else
      vt::imgstats_partial_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<const float*>(d.img), d.count, pmax, psum);
    VT_LAUNCH_CHECK("imgstats_partial_kernel");
    vt::imgstats_final_kernel<<<1, 256, 0, s>>>(pmax, psum, blocks, d.count, d.flags);
    VT_LAUNCH_CHECK("imgstats_final_kernel");
    return VT_OK;
  }
};

struct PatchifyOp : Op {
  vt_patchify_desc d;
  int launch(cudaStream_t s) override {
    const long long total = (long long)d.images * (d.H / d.patch) * (d.W / d.patch) * (d.out_cols / 8);
    const int blocks = grid_for(total, 256);
    if (d.dtype == VT_U8)
      vt::patchify_kernel<uint8_t><<<blocks, 256, 0, s>>>(reinterpret_cast<const uint8_t*>(d.img), d.layout, d.images, d.H,
                                                          d.W, d.patch, d.flags, d.out, d.out_dtype, d.out_cols, d.out_ld, d.out_plane);
    else
      vt::patchify_kernel<float><<<blocks, 256, 0, s>>>(reinterpret_cast<const float*>(d.img), d.layout, d.images, d.H, d.W,
                                                        d.patch, d.flags, d.out, d.out_dtype, d.out_cols, d.out_ld, d.out_plane);
    VT_LAUNCH_CHECK("patchify_kernel");
    return VT_OK;
  }
};

struct ClsOp : Op {
  vt_cls_desc d;
  int launch(cudaStream_t s) override {
    vt::cls_kernel<<<(d.images * d.D + 255) / 256, 256, 0, s>>>(d.cls, d.pos, d.h, d.images, d.tokens, d.D);
    VT_LAUNCH_CHECK("cls_kernel");
    return VT_OK;
  }
};

struct PackOp : Op {
  vt_pack_desc d;
  int launch(cudaStream_t s) override {
    const int width = d.zero_to > d.cols ? d.zero_to : d.cols;
    __nv_bfloat16* out16 = reinterpret_cast<__nv_bfloat16*>(d.out) + d.dst_c0;
    if (d.act == VT_ACT_NONE && d.out_dtype == VT_BF16 && d.src_row_div <= 1 && width == d.cols && d.cols % 8 == 0 && d.src_ld % 4 == 0 &&
        d.out_ld % 8 == 0 && aligned16(d.src) && aligned16(out16)) {
      vt::pack_cast8_kernel<<<grid_for((long long)d.rows * (d.cols / 8), 256), 256, 0, s>>>(d.src, d.src_ld, d.rows, d.cols / 8, out16, d.out_ld);
      VT_LAUNCH_CHECK("pack_cast8_kernel");
      return VT_OK;
    }
    vt::pack_kernel<<<grid_for((long long)d.rows * width, 256), 256, 0, s>>>(d.src, d.src_ld, d.rows, d.cols, d.act, d.out,
                                                                             d.out_dtype, d.out_ld, d.dst_c0, d.out_plane,
                                                                             d.zero_to, d.src_row_div);
    VT_LAUNCH_CHECK("pack_kernel");
    return VT_OK;
  }
};

struct AffineOp : Op {
  vt_affine_desc d;
  int launch(cudaStream_t s) override {
    vt::affine_kernel<<<grid_for((long long)d.rows * d.A, 256), 256, 0, s>>>(d.x, d.out, d.mins, d.maxs, d.rows, d.A, d.denorm,
                                                                             d.pad, d.xpad, d.xpad_dtype, d.xpad_ld, d.xpad_plane,
                                                                             d.add);
    VT_LAUNCH_CHECK("affine_kernel");
    return VT_OK;
  }
};

struct TembedOp : Op {
  vt_tembed_desc d;
  int launch(cudaStream_t s) override {
    vt::tembed_kernel<<<grid_for((long long)d.rows * d.dim / 2, 256), 256, 0, s>>>(d.t, d.rows, d.dim, d.out, d.out_dtype,
                                                                                   d.out_ld, d.out_plane);
    VT_LAUNCH_CHECK("tembed_kernel");
    return VT_OK;
  }
};

struct SdeOp : Op {
  vt_sde_desc d;
  int launch(cudaStream_t s) override {
    launch_ex(vt::sde_step_kernel, dim3(grid_for((long long)d.rows * d.A, 256)), dim3(256), 0, s, false, true,
        d.x, d.v, d.s, d.noise, d.rows, d.A, d.ginv, d.dgg, d.eps, d.dt, d.nscale, d.d, (unsigned long long)d.seed,
        reinterpret_cast<const unsigned long long*>(d.seed_dev), d.step,
        d.xpad, d.xpad_dtype, d.xpad_ld, d.xpad_plane);
    VT_LAUNCH_CHECK("sde_step_kernel");
    return VT_OK;
  }
};

struct QsampleOp : Op {
  vt_qsample_desc d;
  int launch(cudaStream_t s) override {
    vt::qsample_kernel<<<grid_for((long long)d.B * d.n, 256), 256, 0, s>>>(d.x0, d.x1, d.step, d.z_unit, d.d, d.B, d.n, d.A, d.xt,
                                                                           d.tclip, d.xpad, d.xpad_dtype, d.xpad_ld, d.xpad_plane);
    VT_LAUNCH_CHECK("qsample_kernel");
    return VT_OK;
  }
};

struct SilossOp : Op {
  vt_siloss_desc d;
  int launches() const override { return 2; }
  int launch(cudaStream_t s) override {
    vt::siloss_sample_kernel<<<d.B, 128, 0, s>>>(d.bvs, d.x0, d.x1, d.z_unit, d.tclip, d.d, d.B, d.n, d.per_sample);
    VT_LAUNCH_CHECK("siloss_sample_kernel");
    vt::siloss_mean_kernel<<<1, 96, 0, s>>>(d.per_sample, d.B, d.out);
    VT_LAUNCH_CHECK("siloss_mean_kernel");
    return VT_OK;
  }
};

struct TcolOp : Op {
  vt_tcol_desc d;
  int launch(cudaStream_t s) override {
    vt::TcolArgs a;
    a.src = d.src; a.src_dtype = d.src_dtype == VT_BF16 ? 0 : 1;
    a.ld = d.ld; a.sB = d.sB; a.sG = d.sG;
    a.G = d.G; a.B = d.B; a.T_src = d.T_src; a.C = d.C; a.taps = d.taps;
    for (int i = 0; i < 8; ++i) a.tap_off[i] = i < d.taps ? d.tap_off[i] : 0;
    a.stride = d.stride; a.t_out = d.t_out;
    a.out = reinterpret_cast<__nv_bfloat16*>(d.out);
    a.c_pad = d.c_pad; a.k_ld = d.k_ld; a.out_g = d.out_g;
    const dim3 grid((unsigned)((d.B * d.t_out + 31) / 32), (unsigned)((d.C + 31) / 32), (unsigned)(d.G * d.taps));
    vt::tcol_kernel<<<grid, dim3(32, 8, 1), 0, s>>>(a);
    VT_LAUNCH_CHECK("tcol_kernel");
    return VT_OK;
  }
};

struct GnbwdOp : Op {
  vt_gnbwd_desc d;
  int launches() const override { return 2; }
  int launch(cudaStream_t s) override {
    vt::GnBwdArgs a;
    a.raw = d.raw; a.dout = d.dout; a.dout_ld = d.dout_ld; a.dout_g = d.dout_g;
    a.gamma = d.gamma; a.beta = d.beta; a.p_ld = d.p_ld;
    a.film = d.film; a.film_g = d.film_g; a.film_ld = d.film_ld; a.film_off = d.film_off; a.dfilm = d.dfilm;
    a.draw = reinterpret_cast<__nv_bfloat16*>(d.draw); a.part = d.part;
    a.G = d.G; a.B = d.B; a.T = d.T; a.C = d.C; a.groups = d.groups; a.eps = d.eps;
    const size_t smem = vt::gnbs_smem_bytes(d.T, d.C);
    const char* env = getenv("VT_GNBWD_SMEM");
    const int parts = d.C >= 2 ? vt::GNBS_THREADS / (d.C / 2) : 1;
    const bool fits = (d.C == 128 || d.C == 256 || d.C == 512) && d.T % parts == 0 && smem <= 200 * 1024 && d.groups <= vt::GNBS_THREADS / 32 &&
                      (d.C / d.groups) % 2 == 0 && d.dout_ld % 4 == 0 && d.p_ld % 2 == 0 && aligned16(d.raw) && aligned16(d.dout) && d.dout_g % 4 == 0 &&
                      (!d.film || (d.film_g % 2 == 0 && d.film_ld % 2 == 0 && d.film_off % 2 == 0 && ((uintptr_t)d.film & 7) == 0)) &&
                      ((uintptr_t)d.gamma & 7) == 0 && ((uintptr_t)d.beta & 7) == 0 && !(env && atoi(env) == 0);
    if (fits) {   // shared-memory-resident kernel: the sample's tiles are read from HBM once
      static bool attr_set = false;
      if (!attr_set) {
        VT_CUDA(cudaFuncSetAttribute(vt::gn_mish_bwd_smem_kernel<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VT_CUDA(cudaFuncSetAttribute(vt::gn_mish_bwd_smem_kernel<256>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        VT_CUDA(cudaFuncSetAttribute(vt::gn_mish_bwd_smem_kernel<512>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
        attr_set = true;
      }
      const char* e2 = getenv("VT_GNBWD_2CTA");
      const bool two = smem <= 100 * 1024 && !(e2 && atoi(e2) == 0);   // two CTAs per SM when the tiles allow it
      if (two) {
        static bool attr2_set = false;
        if (!attr2_set) {
          VT_CUDA(cudaFuncSetAttribute(vt::gn_mish_bwd_smem_kernel<128, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          VT_CUDA(cudaFuncSetAttribute(vt::gn_mish_bwd_smem_kernel<256, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          VT_CUDA(cudaFuncSetAttribute(vt::gn_mish_bwd_smem_kernel<512, 2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 100 * 1024));
          attr2_set = true;
        }
        if (d.C == 128) vt::gn_mish_bwd_smem_kernel<128, 2><<<d.G * d.B, vt::GNBS_THREADS, smem, s>>>(a);
        else if (d.C == 256) vt::gn_mish_bwd_smem_kernel<256, 2><<<d.G * d.B, vt::GNBS_THREADS, smem, s>>>(a);
        else vt::gn_mish_bwd_smem_kernel<512, 2><<<d.G * d.B, vt::GNBS_THREADS, smem, s>>>(a);
      } else if (const char* e3 = getenv("VT_GNBWD_1024"); d.C >= 256 && d.T % (1024 / (d.C / 2)) == 0 && vt::gnbs_smem_bytes(d.T, d.C, 1024) <= 200 * 1024 &&
                 !(e3 && atoi(e3) == 0)) {
        // the shapes that own an SM alone: 1024 threads (twice the warps per sample) at 64 registers
        static bool attr3_set = false;
        if (!attr3_set) {
          VT_CUDA(cudaFuncSetAttribute(vt::gn_mish_bwd_smem_kernel<256, 1, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          VT_CUDA(cudaFuncSetAttribute(vt::gn_mish_bwd_smem_kernel<512, 1, 1024>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
          attr3_set = true;
        }
        const size_t smem3 = vt::gnbs_smem_bytes(d.T, d.C, 1024);
        if (d.C == 256) vt::gn_mish_bwd_smem_kernel<256, 1, 1024><<<d.G * d.B, 1024, smem3, s>>>(a);
        else vt::gn_mish_bwd_smem_kernel<512, 1, 1024><<<d.G * d.B, 1024, smem3, s>>>(a);
      } else if (d.C == 128) vt::gn_mish_bwd_smem_kernel<128><<<d.G * d.B, vt::GNBS_THREADS, smem, s>>>(a);
      else if (d.C == 256) vt::gn_mish_bwd_smem_kernel<256><<<d.G * d.B, vt::GNBS_THREADS, smem, s>>>(a);
      else vt::gn_mish_bwd_smem_kernel<512><<<d.G * d.B, vt::GNBS_THREADS, smem, s>>>(a);
      VT_LAUNCH_CHECK("gn_mish_bwd_smem_kernel");
    } else {
      vt::gn_mish_bwd_kernel<<<d.G * d.B, 256, 0, s>>>(a);
      VT_LAUNCH_CHECK("gn_mish_bwd_kernel");
    }
    vt::gn_colsum_kernel<<<d.G * 3 * ((d.C + 31) / 32), 256, 0, s>>>(d.part, d.G, d.B, d.C, d.dgamma, d.dbeta, d.dbias, d.p_ld);
    VT_LAUNCH_CHECK("gn_colsum_kernel");
    return VT_OK;
  }
};

struct ColsumOp : Op {
  vt_colsum_desc d;
  int launch(cudaStream_t s) override {
    if (d.rows <= 32 && d.C >= 4096) {   // few rows, many columns (split-K partial sums): elementwise over the columns
      vt::colsum_fewrows_kernel<<<dim3((unsigned)grid_for((long long)d.C / 4, 256), (unsigned)d.G, 1), 256, 0, s>>>(d.x, d.ld, d.x_g, d.rows, d.C,
                                                                                                                 d.out, d.out_ld);
      VT_LAUNCH_CHECK("colsum_fewrows_kernel");
      return VT_OK;
    }
    vt::colsum_kernel<<<dim3((unsigned)((d.C + 31) / 32), (unsigned)d.G, vt::COLSUM_SPLIT), dim3(256, 1, 1), 0, s>>>(
        d.x, d.ld, d.x_g, d.rows, d.C, d.out, d.out_ld);
    VT_LAUNCH_CHECK("colsum_kernel");
    return VT_OK;
  }
};

struct EwiseOp : Op {
  vt_ewise_desc d;
  int launch(cudaStream_t s) override {
    vt::ewise_kernel<<<grid_for(d.rows * d.cols, 256), 256, 0, s>>>(d.a, d.a_ld, d.b, d.b_ld, d.out, d.out_ld, d.rows, d.cols, d.op, d.alpha);
    VT_LAUNCH_CHECK("ewise_kernel");
    return VT_OK;
  }
};

struct SilossBwdOp : Op {
  vt_silossbwd_desc d;
  int launch(cudaStream_t s) override {
    vt::siloss_bwd_kernel<<<grid_for(3ll * d.B * d.n, 256), 256, 0, s>>>(d.bvs, d.x0, d.x1, d.z_unit, d.tclip, d.d, d.B, d.n, d.dvs);
    VT_LAUNCH_CHECK("siloss_bwd_kernel");
    return VT_OK;
  }
};

// tensor-core recurrence (vt_lstm_tc.cuh): shared by the inference and the training forward
constexpr int LSTM_TC_MIN_BATCH = 16;
int build_lstm_tc(vt::LstmTcArgs* a, const void* w_hh_tc, void* h_tc, const float* xw, void* y, int y_dtype, long long y_ld,
                  float* gates, float* c_all, float* h_out, float* c_out, int B, int T) {
  const int H = vt::LTC_H;
  memset(a, 0, sizeof(*a));
  {
    const uint64_t dims[2] = {(uint64_t)H, (uint64_t)4 * H};
    const uint64_t st[1] = {(uint64_t)H * 2};
    const uint32_t box[2] = {64u, 128u};
    int rc = make_tmap(&a->tmW, VT_BF16, 2, w_hh_tc, dims, st, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)H, (uint64_t)T, (uint64_t)B};
    const uint64_t st[2] = {(uint64_t)H * 2, (uint64_t)T * H * 2};
    const uint32_t box[3] = {64u, 1u, 128u};
    int rc = make_tmap(&a->tmH, VT_BF16, 3, h_tc, dims, st, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[3] = {(uint64_t)4 * H, (uint64_t)T, (uint64_t)B};
    const uint64_t st[2] = {(uint64_t)4 * H * 4, (uint64_t)T * 4 * H * 4};
    const uint32_t box[3] = {32u, 1u, 128u};
    int rc = make_tmap(&a->tmX, VT_F32, 3, xw, dims, st, box);
    if (rc) return rc;
  }
  VT_REQUIRE(!y || (y_ld % 8 == 0 && aligned16(y)), "lstm (tensor-core path): y rows must be 16-byte aligned");
  a->hbuf = reinterpret_cast<__nv_bfloat16*>(h_tc);
  a->y = y; a->y_dtype = y_dtype == VT_BF16 ? 0 : 1; a->y_ld = y_ld;
  a->gates = gates; a->c_all = c_all; a->h_out = h_out; a->c_out = c_out;
  a->B = B; a->T = T;
  return VT_OK;
}
int launch_lstm_tc(const vt::LstmTcArgs& a, cudaStream_t s) {
  static bool attr_set = false;
  if (!attr_set) {
    VT_CUDA(cudaFuncSetAttribute(vt::lstm_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vt::LTC_FWD_SMEM));
    attr_set = true;
  }
  const unsigned grid = (unsigned)((a.B + vt::LTC_ROWS - 1) / vt::LTC_ROWS) * vt::LTC_CLUSTER;
  vt::lstm_tc_kernel<<<grid, vt::LTC_THREADS, vt::LTC_FWD_SMEM, s>>>(a);
  VT_LAUNCH_CHECK("lstm_tc_kernel");
  return VT_OK;
}

struct LstmOp : Op {
  vt_lstm_desc d;
  vt::LstmTcArgs tc;
  bool use_tc = false;
  int launch(cudaStream_t s) override {
    if (use_tc) return launch_lstm_tc(tc, s);
    const int blocks = (d.B + vt::LSTM_ROWS - 1) / vt::LSTM_ROWS;
    if (d.H == 256)
      vt::lstm_seq_kernel<256><<<blocks, 256, 0, s>>>(d.xw, d.w_hh, d.h, d.c, d.y, d.y_dtype, d.y_ld, d.y_plane, d.B, d.T);
    else
      return fail(VT_E_UNSUPPORTED, "lstm: H=%d", d.H);
    VT_LAUNCH_CHECK("lstm_seq_kernel");
    return VT_OK;
  }
};

struct DropmaskOp : Op {
  vt_dropmask_desc d;
  int launch(cudaStream_t s) override {
    vt::dropmask_kernel<<<grid_for(d.n, 256), 256, 0, s>>>(d.inject, d.p, (unsigned long long)d.seed,
                                                          reinterpret_cast<const unsigned long long*>(d.seed_dev), d.stream, d.mask, d.n);
    VT_LAUNCH_CHECK("dropmask_kernel");
    return VT_OK;
  }
};

struct LnGeluBwdOp : Op {
  vt_lngelubwd_desc d;
  int launch(cudaStream_t s) override {
    vt::ln_gelu_bwd_kernel<<<(d.rows + 7) / 8, 256, 0, s>>>(d.z0, d.dzn, d.gamma, d.beta, d.eps, d.dz0, d.d1, d.d1zh, d.rows);
    VT_LAUNCH_CHECK("ln_gelu_bwd_kernel");
    return VT_OK;
  }
};

struct LstmTrainOp : Op {
  vt_lstm_train_desc d;
  vt::LstmTcArgs tc;
  bool use_tc = false;
  int launch(cudaStream_t s) override {
    if (use_tc) return launch_lstm_tc(tc, s);
    const int blocks = (d.B + vt::LSTM_ROWS - 1) / vt::LSTM_ROWS;
    vt::lstm_seq_train_kernel<256><<<blocks, 256, 0, s>>>(d.xw, d.w_hh, d.y, d.y_dtype, d.y_ld, d.gates, d.c, d.B, d.T);
    VT_LAUNCH_CHECK("lstm_seq_train_kernel");
    return VT_OK;
  }
};

struct LstmBwdOp : Op {
  vt_lstm_bwd_desc d;
  vt::LstmBwdTcArgs tc;
  bool use_tc = false;
  int launch(cudaStream_t s) override {
    if (use_tc) {
      static bool attr_set = false;
      if (!attr_set) {
        VT_CUDA(cudaFuncSetAttribute(vt::lstm_bwd_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vt::LTC_BWD_SMEM));
        attr_set = true;
      }
      const unsigned grid = (unsigned)((d.B + vt::LTC_ROWS - 1) / vt::LTC_ROWS) * vt::LTC_CLUSTER;
      vt::lstm_bwd_tc_kernel<<<grid, vt::LTC_THREADS, vt::LTC_BWD_SMEM, s>>>(tc);
      VT_LAUNCH_CHECK("lstm_bwd_tc_kernel");
      return VT_OK;
    }
    const int blocks = (d.B + vt::LSTM_ROWS - 1) / vt::LSTM_ROWS;
    vt::lstm_bwd_kernel<256><<<blocks, 256, 0, s>>>(d.gates, d.c, d.dy, d.dy_ld, d.w_hh, d.dgates, d.B, d.T);
    VT_LAUNCH_CHECK("lstm_bwd_kernel");
    return VT_OK;
  }
};

// ---- weight gradients with MN-major operands ----
struct WgradOp : Op {
  vt::WgradArgs args;
  dim3 grid;
  int launch(cudaStream_t s) override {
    static bool attr_set = false;
    if (!attr_set) {
      VT_CUDA(cudaFuncSetAttribute(vt::wgrad_tc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vt::WG_SMEM_BYTES));
      attr_set = true;
    }
    cudaError_t e = launch_ex(vt::wgrad_tc_kernel, grid, dim3(vt::WG_THREADS, 1, 1), (size_t)vt::WG_SMEM_BYTES, s, true, true, args);
    if (e != cudaSuccess) return fail(VT_E_CUDA, "wgrad_tc_kernel launch: %s", cudaGetErrorString(e));
    VT_LAUNCH_CHECK("wgrad_tc_kernel");
    return VT_OK;
  }
};

int build_wgrad(const vt_wgrad_desc& d, WgradOp* op) {
  VT_REQUIRE(d.rows && d.cols && d.out && d.G >= 1 && d.B >= 1 && d.R >= 1, "wgrad: bad descriptor");
  VT_REQUIRE(d.t_out >= 1 && d.t_out <= 64 && 64 % d.t_out == 0, "wgrad: t_out=%d must divide 64", d.t_out);
  VT_REQUIRE(d.taps >= 1 && d.taps <= VT_MAX_TAPS && d.c_pad >= 64 && d.c_pad % 64 == 0, "wgrad: taps=%d c_pad=%d", d.taps, d.c_pad);
  VT_REQUIRE(d.rows_P >= 1 && d.cols_P >= 1 && d.rows_T >= 1 && d.cols_T >= 1 && d.rows_C >= 1 && d.cols_C >= 1, "wgrad: bad operand extents");
  VT_REQUIRE(d.ldc >= d.taps * d.c_pad || d.ldc >= 1, "wgrad: ldc");
  vt::WgradArgs& a = op->args;
  memset(&a, 0, sizeof(a));
  const uint32_t b_box = (uint32_t)(64 / d.t_out);
  {
    const uint64_t dims[5] = {(uint64_t)d.rows_C, (uint64_t)d.rows_P, (uint64_t)d.rows_T, (uint64_t)d.B, (uint64_t)d.G};
    const uint64_t st[4] = {(uint64_t)d.rows_ld * 2, (uint64_t)d.rows_ld * d.rows_P * 2, (uint64_t)d.rows_sB * 2,
                            (uint64_t)(d.G > 1 ? d.rows_sG : d.rows_sB * d.B) * 2};
    const uint32_t box[5] = {64u, 1u, (uint32_t)d.t_out, b_box, 1u};
    int rc = make_tmap(&a.tmR, VT_BF16, 5, d.rows, dims, st, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[5] = {(uint64_t)d.cols_C, (uint64_t)d.cols_P, (uint64_t)d.cols_T, (uint64_t)d.B, (uint64_t)d.G};
    const uint64_t st[4] = {(uint64_t)d.cols_ld * 2, (uint64_t)d.cols_ld * d.cols_P * 2, (uint64_t)d.cols_sB * 2,
                            (uint64_t)(d.G > 1 ? d.cols_sG : d.cols_sB * d.B) * 2};
    const uint32_t box[5] = {64u, 1u, (uint32_t)d.t_out, b_box, 1u};
    int rc = make_tmap(&a.tmC, VT_BF16, 5, d.cols, dims, st, box);
    if (rc) return rc;
  }
  a.rows_p = d.rows_p;
  a.rows_t = d.rows_t;
  a.taps = d.taps;
  a.c_pad = d.c_pad;
  for (int i = 0; i < d.taps; ++i) {
    VT_REQUIRE(d.tap_p[i] >= 0 && d.tap_p[i] < d.cols_P, "wgrad: tap %d phase %d", i, d.tap_p[i]);
    a.tap_p[i] = d.tap_p[i];
    a.tap_t[i] = d.tap_t[i];
  }
  VT_REQUIRE(d.rows_p >= 0 && d.rows_p < d.rows_P, "wgrad: rows phase %d", d.rows_p);
  a.k_tiles = (d.B + (int)b_box - 1) / (int)b_box;
  a.k_t_step = 0;
  a.k_b_step = (int)b_box;
  a.box_bytes = 64 * 128;
  const int N = d.taps * d.c_pad;
  a.m_units = (d.R + 255) / 256;
  a.n_tiles = (N + 255) / 256;
  const long long total = (long long)a.m_units * a.n_tiles * d.G;
  VT_REQUIRE(total < (1ll << 31), "wgrad: too many tiles");
  a.total_tiles = (int)total;
  vt::GemmArgs& e = a.epi;
  e.rows_valid = 128;
  e.n_pad = a.n_tiles * 256;
  e.M_total = d.R;
  e.N = N;
  e.row_div = 1;
  e.out_q = 1;
  e.out_g = d.out_g;
  e.ldc = d.ldc;
  e.out = d.out;
  e.act = VT_ACT_NONE;
  e.vec = (aligned16(d.out) && d.ldc % 4 == 0 && d.out_g % 4 == 0) ? 1 : 0;
  e.fast = (e.vec && N % 256 == 0 && ((long long)d.R + 1) * d.ldc < (1ll << 31)) ? 1 : 0;
  const int workers = sm_count() / 2;
  op->grid = dim3((unsigned)(total < workers ? total : workers) * 2u, 1u, 1u);
  return VT_OK;
}

// ---- persistent multi-layer launch ----
struct PersistOp : Op {
  vt::PersistArgs args;
  dim3 grid;
  void* dev_mem = nullptr;      // counters | expected | coefficients, one allocation
  size_t counter_bytes = 0;
  ~PersistOp() override {
    if (dev_mem) cudaFree(dev_mem);
  }
  int launches() const override { return 1; }
  int launch(cudaStream_t s) override {
    static bool attr_set = false;
    if (!attr_set) {
      VT_CUDA(cudaFuncSetAttribute(vt::unet_persist_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, vt::PERSIST_SMEM_BYTES));
      attr_set = true;
    }
    VT_CUDA(cudaMemsetAsync(args.done, 0, counter_bytes, s));
    cudaError_t e = launch_ex(vt::unet_persist_kernel, grid, dim3(vt::PERSIST_THREADS, 1, 1), (size_t)vt::PERSIST_SMEM_BYTES, s, true, false, args);
    if (e != cudaSuccess) return fail(VT_E_CUDA, "unet_persist_kernel launch: %s", cudaGetErrorString(e));
    VT_LAUNCH_CHECK("unet_persist_kernel");
    return VT_OK;
  }
};

int build_persist(const vt_persist_desc& d, PersistOp* op) {
  static_assert(VT_PERSIST_MAX_DEPS == vt::PERSIST_MAX_DEPS && VT_PERSIST_MAX_LAYERS + 1 == vt::PERSIST_MAX_LAYERS, "header constants");
  static_assert(sizeof(vt::PersistArgs) <= 32764, "kernel parameter space");
  VT_REQUIRE(d.gemms && d.deps && d.dep_lag && d.n_gemms >= 1 && d.n_gemms <= VT_PERSIST_MAX_LAYERS && d.n_steps >= 1,
             "persist: bad descriptor (n_gemms=%d, n_steps=%d)", d.n_gemms, d.n_steps);
  VT_REQUIRE(!d.sde || (d.sde_coef && d.sde_T >= 1), "persist: the Euler-Maruyama update needs sde_coef and sde_T");
  vt::PersistArgs& P = op->args;
  memset(&P, 0, sizeof(P));
  const int n_layers = d.n_gemms + (d.sde ? 1 : 0);
  const int B = d.gemms[0].a_B;
  int sbs = 1;
  for (int l = 0; l < d.n_gemms; ++l) {
    const vt_gemm_desc& g = d.gemms[l];
    int rc = validate_gemm(g);
    if (rc) return rc;
    VT_REQUIRE(g.in_dtype == VT_BF16 && g.passes == 1, "persist: layer %d: bf16 operands only", l);
    VT_REQUIRE(g.a_B == B && g.M % B == 0 && g.t_box == g.M / B, "persist: layer %d: tiles must hold whole samples of the same batch", l);
    if (g.epi == VT_EPI_GN) VT_REQUIRE(g.bn == 256 && g.out_dtype == VT_BF16, "persist: layer %d: GroupNorm layers need bn = 256, bf16 out", l);
    else VT_REQUIRE(g.bn == 256 || g.bn == 32, "persist: layer %d: bn = %d (256 or 32)", l, g.bn);
    if (2 * g.b_box > sbs) sbs = 2 * g.b_box;
  }
  const int n_sb = (B + sbs - 1) / sbs;
  std::vector<unsigned> expected;
  int tile_base = 0, cnt_off = 0, max_tiles = 1;
  for (int l = 0; l < n_layers; ++l) {
    vt::PersistLayer& L = P.layers[l];
    L.exp_off = (int)expected.size();
    L.cnt_off = cnt_off;
    L.tile_base = tile_base;
    if (l < d.n_gemms) {
      const vt_gemm_desc& g = d.gemms[l];
      int rc = fill_gemm_args(g, 2, &L.g);
      if (rc) return rc;
      L.kind = 0;
      L.mode = g.epi;
      L.bn = g.bn;
      L.out_f32 = g.out_dtype == VT_F32;
      L.idesc = vt::umma_idesc(vt::UMMA_FMT_BF16, (uint32_t)g.bn, 0, 0, 256);
      L.G = g.G;
      L.spu = 2 * g.b_box;
      L.n_tiles_total = L.g.total_tiles;
      L.film_t_step = g.film_t ? d.film_t_step : 0;
      VT_REQUIRE(L.film_t_step % 4 == 0, "persist: film_t_step must be a multiple of 4 elements");
      for (int sb = 0; sb < n_sb; ++sb) {
        int units = 0;
        for (int u = 0; u < L.g.m_tiles; ++u) {
          const int s0 = u * L.spu, s1 = (s0 + L.spu < B ? s0 + L.spu : B) - 1;
          if (s0 < B && s0 / sbs <= sb && sb <= s1 / sbs) ++units;
        }
        expected.push_back(16u * (unsigned)units * (unsigned)L.g.n_tiles);
      }
    } else {
      L.kind = 1;
      L.G = 1;
      L.spu = sbs;
      L.n_tiles_total = n_sb;
      for (int sb = 0; sb < n_sb; ++sb) expected.push_back(16u);
    }
    cnt_off += L.G * n_sb;
    tile_base += L.n_tiles_total;
    if (L.n_tiles_total > max_tiles) max_tiles = L.n_tiles_total;
    int nd = 0;
    for (int k = 0; k < VT_PERSIST_MAX_DEPS; ++k) {
      const int dep = d.deps[l * VT_PERSIST_MAX_DEPS + k];
      if (dep < 0) continue;
      const int lag = d.dep_lag[l * VT_PERSIST_MAX_DEPS + k];
      VT_REQUIRE(dep < n_layers && (lag == 0 || lag == 1), "persist: layer %d depends on layer %d (lag %d)", l, dep, lag);
      VT_REQUIRE(lag == 1 || dep < l, "persist: layer %d depends on the later layer %d of the same step", l, dep);
      L.dep[nd] = dep;
      L.dep_lag[nd] = lag;
      ++nd;
    }
    L.n_dep = nd;
  }
  P.n_layers = n_layers;
  P.n_steps = d.n_steps;
  P.tiles_per_step = tile_base;
  P.n_sb = n_sb;
  P.sbs = sbs;
  P.B = B;
  VT_REQUIRE((long long)tile_base * d.n_steps < (1ll << 31), "persist: too many tiles");
  if (d.sde) {
    const vt_sde_desc& s = *d.sde;
    VT_REQUIRE(s.x && s.v && s.s && s.A >= 1 && s.rows == B * d.sde_T, "persist: bad Euler-Maruyama descriptor");
    VT_REQUIRE(s.s == s.v + (long long)s.rows * s.A, "persist: v and s must be consecutive groups of the nets' output");
    vt::PersistSde& S = P.sde;
    S.x = s.x; S.v = s.v; S.s = s.s; S.noise = s.noise; S.noise_step = d.noise_step;
    S.A = s.A; S.T = d.sde_T; S.d = s.d; S.seed = s.seed;
    S.seed_dev = reinterpret_cast<const unsigned long long*>(s.seed_dev);
    S.xpad = s.xpad; S.xpad_dtype = s.xpad_dtype; S.xpad_ld = s.xpad_ld; S.xpad_plane = s.xpad_plane;
  }
  // device tables: counters | expected | coefficients
  const size_t cnt_bytes = ((size_t)cnt_off * 4 + 255) / 256 * 256;
  const size_t exp_bytes = (expected.size() * 4 + 255) / 256 * 256;
  const size_t coef_bytes = (size_t)d.n_steps * sizeof(vt::PersistCoef);
  VT_CUDA(cudaMalloc(&op->dev_mem, cnt_bytes + exp_bytes + coef_bytes));
  char* base = reinterpret_cast<char*>(op->dev_mem);
  P.done = reinterpret_cast<unsigned*>(base);
  P.expected = reinterpret_cast<const unsigned*>(base + cnt_bytes);
  P.coef = reinterpret_cast<const vt::PersistCoef*>(base + cnt_bytes + exp_bytes);
  op->counter_bytes = cnt_bytes;
  VT_CUDA(cudaMemcpy(base + cnt_bytes, expected.data(), expected.size() * 4, cudaMemcpyHostToDevice));
  std::vector<vt::PersistCoef> coef((size_t)d.n_steps);
  for (int k = 0; k < d.n_steps; ++k) {
    vt::PersistCoef c;
    memset(&c, 0, sizeof(c));
    if (d.sde_coef) {
      c.ginv = d.sde_coef[5 * k]; c.dgg = d.sde_coef[5 * k + 1]; c.eps = d.sde_coef[5 * k + 2];
      c.dt = d.sde_coef[5 * k + 3]; c.nscale = d.sde_coef[5 * k + 4];
    }
    coef[(size_t)k] = c;
  }
  VT_CUDA(cudaMemcpy(base + cnt_bytes + exp_bytes, coef.data(), coef_bytes, cudaMemcpyHostToDevice));
  const int workers_max = sm_count() / 2;
  const int workers = max_tiles < workers_max ? max_tiles : workers_max;
  op->grid = dim3((unsigned)workers * 2u, 1u, 1u);
  return VT_OK;
}

}  // namespace

struct vt_program {
  std::vector<std::unique_ptr<Op>> ops;
  cudaGraphExec_t graph_exec = nullptr;
  cudaStream_t capture_stream = nullptr;
};

extern "C" {

const char* vt_last_error(void) { return g_err.c_str(); }
int vt_abi_version(void) { return VT_ABI_VERSION; }

int vt_device_info(int32_t* sm_count, int32_t* cc_major, int32_t* cc_minor) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess || n == 0) {
    cudaGetLastError();
    return fail(VT_E_NODEVICE, "no CUDA device");
  }
  int dev = 0;
  VT_CUDA(cudaGetDevice(&dev));
  cudaDeviceProp p;
  VT_CUDA(cudaGetDeviceProperties(&p, dev));
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return VT_OK;
}

int vt_program_create(vt_program** out) {
  if (!out) return fail(VT_E_INVALID, "null out");
  *out = new vt_program();
  return VT_OK;
}

int vt_program_destroy(vt_program* p) {
  if (!p) return VT_OK;
  if (p->graph_exec) cudaGraphExecDestroy(p->graph_exec);
  if (p->capture_stream) cudaStreamDestroy(p->capture_stream);
  delete p;
  return VT_OK;
}

int vt_program_num_ops(const vt_program* p) { return p ? (int)p->ops.size() : 0; }

static int clamp_range(const vt_program* p, int first, int* count) {
  if (!p) return fail(VT_E_INVALID, "null program");
  const int n = (int)p->ops.size();
  if (first < 0 || first > n) return fail(VT_E_INVALID, "first=%d out of range (%d ops)", first, n);
  if (*count < 0 || first + *count > n) *count = n - first;
  return VT_OK;
}

int vt_program_num_launches(const vt_program* p, int first, int count) {
  if (clamp_range(p, first, &count)) return VT_E_INVALID;
  int k = 0;
  for (int i = first; i < first + count; ++i) k += p->ops[i]->launches();
  return k;
}

int vt_program_add_gemm(vt_program* p, const vt_gemm_desc* d) {
  if (!p || !d) return fail(VT_E_INVALID, "null argument");
  std::unique_ptr<GemmOp> op(new GemmOp());
  int rc = build_gemm(*d, op.get());
  if (rc) return rc;
  p->ops.push_back(std::move(op));
  return VT_OK;
}

int vt_program_add_persist(vt_program* p, const vt_persist_desc* d) {
  if (!p || !d) return fail(VT_E_INVALID, "null argument");
  std::unique_ptr<PersistOp> op(new PersistOp());
  int rc = build_persist(*d, op.get());
  if (rc) return rc;
  p->ops.push_back(std::move(op));
  return VT_OK;
}

int vt_program_add_wgrad(vt_program* p, const vt_wgrad_desc* d) {
  if (!p || !d) return fail(VT_E_INVALID, "null argument");
  std::unique_ptr<WgradOp> op(new WgradOp());
  int rc = build_wgrad(*d, op.get());
  if (rc) return rc;
  p->ops.push_back(std::move(op));
  return VT_OK;
}

// developer instrumentation: copy out (and reset) the epilogue timestamps recorded under VT_GEMM_DEBUG bit 128
int vt_debug_timestamps(long long* out, int max_entries) {
  int n = 0;
  if (cudaMemcpyFromSymbol(&n, vt::vt_dbg_n, sizeof(int)) != cudaSuccess) return -1;
  if (n > max_entries) n = max_entries;
  if (n > 0 && cudaMemcpyFromSymbol(out, vt::vt_dbg_ts, sizeof(long long) * n) != cudaSuccess) return -1;
  const int zero = 0;
  cudaMemcpyToSymbol(vt::vt_dbg_n, &zero, sizeof(int));
  return n;
}

int vt_debug_persist_trace(long long* out, long long* cal, int32_t* dims) {
  if (dims) {
    dims[0] = vt::PTRACE_WORKERS;
    dims[1] = vt::PTRACE_TILES;
    dims[2] = vt::PTRACE_SLOTS;
  }
#if VT_DEBUG_KNOBS
  if (out && cudaMemcpyFromSymbol(out, vt::vt_ptrace, sizeof(vt::vt_ptrace)) != cudaSuccess) return -1;
  if (cal && cudaMemcpyFromSymbol(cal, vt::vt_ptrace_cal, sizeof(vt::vt_ptrace_cal)) != cudaSuccess) return -1;
  static std::vector<long long> zeros(sizeof(vt::vt_ptrace) / sizeof(long long), 0);
  cudaMemcpyToSymbol(vt::vt_ptrace, zeros.data(), sizeof(vt::vt_ptrace));
  return 0;
#else
  return -1;
#endif
}

int vt_program_add_mlp(vt_program* p, const vt_mlp_desc* d) {
  if (!p || !d) return fail(VT_E_INVALID, "null argument");
  VT_REQUIRE(d->xn && d->w1 && d->b1 && d->w2 && d->b2 && d->h && d->rows >= 1, "mlp: bad descriptor");
  VT_REQUIRE(d->D == vt::MLP_D, "mlp: the fused kernel is built for D = %d (got %d)", vt::MLP_D, d->D);
  VT_REQUIRE(d->ld_x >= d->D && d->ld_x % 8 == 0 && d->w1_ld >= d->D && d->w1_ld % 8 == 0 && d->w2_ld >= 4 * d->D && d->w2_ld % 8 == 0,
             "mlp: leading dimensions");
  VT_REQUIRE(d->ld_h >= d->D && d->ld_h % 4 == 0 && aligned16(d->h) && aligned16(d->b1) && aligned16(d->b2) && (!d->ls2 || aligned16(d->ls2)),
             "mlp: fp32 operands must be 16-byte aligned");
  VT_REQUIRE((long long)d->rows * d->ld_h < (1ll << 31), "mlp: residual stream too large for 32-bit offsets");
  std::unique_ptr<MlpOp> op(new MlpOp());
  vt::MlpArgs& a = op->args;
  memset(&a, 0, sizeof(a));
  {
    const uint64_t dims[2] = {(uint64_t)d->D, (uint64_t)d->rows};
    const uint64_t st[1] = {(uint64_t)d->ld_x * 2};
    const uint32_t box[2] = {64u, 128u};
    int rc = make_tmap(&a.tmX, VT_BF16, 2, d->xn, dims, st, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d->D, (uint64_t)4 * d->D};
    const uint64_t st[1] = {(uint64_t)d->w1_ld * 2};
    const uint32_t box[2] = {64u, 64u};
    int rc = make_tmap(&a.tmW1, VT_BF16, 2, d->w1, dims, st, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)4 * d->D, (uint64_t)d->D};
    const uint64_t st[1] = {(uint64_t)d->w2_ld * 2};
    const uint32_t box[2] = {64u, 96u};
    int rc = make_tmap(&a.tmW2, VT_BF16, 2, d->w2, dims, st, box);
    if (rc) return rc;
  }
  a.b1 = d->b1;
  vt::GemmArgs& e = a.epi;
  e.M_total = d->rows;
  e.N = d->D;
  e.n_pad = d->D;
  e.rows_valid = 128;
  e.row_div = 1;
  e.out_q = 1;
  e.res_q = 1;
  e.ldc = (int)d->ld_h;
  e.ldres = (int)d->ld_h;
  e.out = d->h;
  e.res = d->h;
  e.bias = d->b2;
  e.colscale = d->ls2;
  e.act = VT_ACT_NONE;
  e.vec = 1;
  e.fast = 1;
  {
    const char* dbg = VT_DEBUG_KNOBS ? getenv("VT_GEMM_DEBUG") : nullptr;
    e.debug = dbg ? (atoi(dbg) & 128) : 0;
  }
  a.m_tiles = (d->rows + 127) / 128;
  a.n_pairs = (a.m_tiles + 1) / 2;
  if (d->ln_out) {
    VT_REQUIRE(d->ln_gamma && d->ln_beta && d->ln_ld >= d->D && d->ln_ld % 4 == 0 && aligned16(d->ln_out) && aligned16(d->ln_gamma) &&
                   aligned16(d->ln_beta) && d->ld_h % 4 == 0,
               "mlp: fused LayerNorm output needs 16-byte aligned rows / vectors");
    a.ln_gamma = d->ln_gamma;
    a.ln_beta = d->ln_beta;
    a.ln_out = reinterpret_cast<__nv_bfloat16*>(d->ln_out);
    a.ln_ld = d->ln_ld;
    a.ln_eps = d->ln_eps;
    const char* ldbg = VT_DEBUG_KNOBS ? getenv("VT_MLP_LN_DEBUG") : nullptr;
    a.ln_debug = ldbg ? atoi(ldbg) : 0;
  }
  const int workers = sm_count() / 2;
  op->grid = dim3((unsigned)(a.n_pairs < workers ? a.n_pairs : workers) * 2u, 1u, 1u);
  p->ops.push_back(std::move(op));
  return VT_OK;
}

int vt_program_add_rowproj(vt_program* p, const vt_rowproj_desc* d) {
  if (!p || !d) return fail(VT_E_INVALID, "null argument");
  VT_REQUIRE(d->x && d->w && d->bias && d->h && d->rows >= 1, "rowproj: bad descriptor");
  VT_REQUIRE(d->D == vt::RP_D, "rowproj: the kernel is built for D = %d (got %d)", vt::RP_D, d->D);
  VT_REQUIRE(d->ld_x >= d->D && d->ld_x % 8 == 0 && d->w_ld >= d->D && d->w_ld % 8 == 0, "rowproj: leading dimensions");
  VT_REQUIRE(d->ld_h >= d->D && d->ld_h % 4 == 0 && aligned16(d->h) && aligned16(d->bias) && (!d->colscale || aligned16(d->colscale)),
             "rowproj: fp32 operands must be 16-byte aligned");
  VT_REQUIRE((long long)d->rows * d->ld_h < (1ll << 31), "rowproj: residual stream too large for 32-bit offsets");
  std::unique_ptr<RowprojOp> op(new RowprojOp());
  vt::RowprojArgs& a = op->args;
  memset(&a, 0, sizeof(a));
  {
    const uint64_t dims[2] = {(uint64_t)d->D, (uint64_t)d->rows};
    const uint64_t st[1] = {(uint64_t)d->ld_x * 2};
    const uint32_t box[2] = {64u, 128u};
    int rc = make_tmap(&a.tmX, VT_BF16, 2, d->x, dims, st, box);
    if (rc) return rc;
  }
  {
    const uint64_t dims[2] = {(uint64_t)d->D, (uint64_t)d->D};
    const uint64_t st[1] = {(uint64_t)d->w_ld * 2};
    const uint32_t box[2] = {64u, 32u};
    int rc = make_tmap(&a.tmW, VT_BF16, 2, d->w, dims, st, box);
    if (rc) return rc;
  }
  vt::GemmArgs& e = a.epi;
  e.M_total = d->rows;
  e.N = d->D;
  e.n_pad = d->D;
  e.rows_valid = 128;
  e.row_div = 1;
  e.out_q = 1;
  e.res_q = 1;
  e.ldc = (int)d->ld_h;
  e.ldres = (int)d->ld_h;
  e.out = d->h;
  e.res = d->h;
  e.bias = d->bias;
  e.colscale = d->colscale;
  e.act = VT_ACT_NONE;
  e.vec = 1;
  e.fast = 1;
  a.m_tiles = (d->rows + 127) / 128;
  a.n_pairs = (a.m_tiles + 1) / 2;
  if (d->ln_out) {
    VT_REQUIRE(d->ln_gamma && d->ln_beta && d->ln_ld >= d->D && d->ln_ld % 4 == 0 && aligned16(d->ln_out) && aligned16(d->ln_gamma) &&
                   aligned16(d->ln_beta),
               "rowproj: fused LayerNorm output needs 16-byte aligned rows / vectors");
    const char* ldbg = VT_DEBUG_KNOBS ? getenv("VT_MLP_LN_DEBUG") : nullptr;
    a.ln_debug = ldbg ? atoi(ldbg) : 0;
    a.ln_gamma = d->ln_gamma;
    a.ln_beta = d->ln_beta;
    a.ln_out = reinterpret_cast<__nv_bfloat16*>(d->ln_out);
    a.ln_ld = d->ln_ld;
    a.ln_eps = d->ln_eps;
  }
  const int workers = sm_count() / 2;
  op->grid = dim3((unsigned)(a.n_pairs < workers ? a.n_pairs : workers) * 2u, 1u, 1u);
  p->ops.push_back(std::move(op));
  return VT_OK;
}

int vt_program_add_attention(vt_program* p, const vt_attn_desc* d) {
  if (!p || !d) return fail(VT_E_INVALID, "null argument");
  VT_REQUIRE(d->qkv && d->ctx && d->images >= 1 && d->tokens >= 1 && d->heads >= 1, "attention: bad descriptor");
  std::unique_ptr<AttnOp> op(new AttnOp());
  op->d = *d;
  if (d->in_dtype == VT_BF16) {
    const int D = d->heads * 64;
    const uint64_t dims[2] = {(uint64_t)3 * D, (uint64_t)d->images * d->tokens};
    const uint64_t st[1] = {(uint64_t)3 * D * 2};
    const uint32_t box[2] = {64u, 128u};
    int rc = make_tmap(&op->args.tm, VT_BF16, 2, d->qkv, dims, st, box);
    if (rc) return rc;
    VT_REQUIRE(aligned16(d->ctx), "attention: ctx must be 16-byte aligned");
    op->args.ctx = reinterpret_cast<__nv_bfloat16*>(d->ctx);
    op->args.ctx_ld = d->ctx_ld;
    VT_REQUIRE(d->ctx_ld >= D && d->ctx_ld % 8 == 0, "attention: ctx_ld=%lld", (long long)d->ctx_ld);
    op->args.tokens = d->tokens;
    op->args.heads = d->heads;
    op->args.D = D;
    op->args.scale_log2 = 0.125f * 1.4426950408889634f;
    int q_tiles = (d->tokens + 127) / 128;
    const int rem = d->tokens % 128;
    if (rem > 0 && rem <= vt::ATT_TAIL_MAX && (size_t)(d->tokens + 264) * 4 <= 48 * 1024) {
      op->tail_first = d->tokens - rem;   // the short last tile runs on the CUDA cores
      q_tiles -= 1;
    }
    op->grid = dim3((unsigned)q_tiles, (unsigned)d->heads, (unsigned)d->images);
    VT_REQUIRE(d->images <= 65535, "attention: images=%d exceeds the grid z limit", d->images);
    const char* rk = getenv("VT_ATTN_ROW");
    if (d->tokens >= 256 && d->tokens <= 272 && !(rk && atoi(rk) == 0) && (long long)d->images * d->tokens < (1ll << 31)) {
      vt::AttnRowArgs& r = op->rargs;
      r.tm = op->args.tm;
      const uint32_t box16[2] = {64u, 16u};
      rc = make_tmap(&r.tm16, VT_BF16, 2, d->qkv, dims, st, box16);
      if (rc) return rc;
      const uint64_t odims[2] = {(uint64_t)D, (uint64_t)d->images * d->tokens};
      const uint64_t ost[1] = {(uint64_t)d->ctx_ld * 2};
      rc = make_tmap(&r.tmO, VT_BF16, 2, d->ctx, odims, ost, box);
      if (rc) return rc;
      r.qkv = reinterpret_cast<const __nv_bfloat16*>(d->qkv);
      r.ctx = op->args.ctx;
      r.ctx_ld = d->ctx_ld;
      r.tokens = d->tokens;
      r.heads = d->heads;
      r.D = D;
      r.units = d->images * d->heads;
      r.scale_log2 = op->args.scale_log2;
      op->row_kernel = true;
      const char* pp = getenv("VT_ATTN_PP");
      op->pp_kernel = d->tokens == 257 && pp && atoi(pp) != 0;   // opt-in until it beats attn_row_kernel (profiles/r02_experiments.txt)
      const int sms = sm_count();
      op->grid = dim3((unsigned)(r.units < sms ? r.units : sms), 1u, 1u);
    }
  } else {
    VT_REQUIRE(d->in_dtype == VT_F32, "attention: in_dtype %d", d->in_dtype);
    VT_REQUIRE((size_t)vt::ATTF_WARPS * d->tokens * 4 <= 48 * 1024, "attention(f32): tokens=%d too many", d->tokens);
  }
  p->ops.push_back(std::move(op));
  return VT_OK;
}

#define VT_SIMPLE_ADD(fname, OpT, DescT, check)                  \
  int fname(vt_program* p, const DescT* d) {                     \
    if (!p || !d) return fail(VT_E_INVALID, "null argument");    \
    check;                                                       \
    std::unique_ptr<OpT> op(new OpT());                          \
    op->d = *d;                                                  \
    p->ops.push_back(std::move(op));                             \
    return VT_OK;                                                \
  }

VT_SIMPLE_ADD(vt_program_add_layernorm, LnOp, vt_ln_desc,
              VT_REQUIRE(d->x && d->out && d->gamma && d->beta && d->rows >= 1 &&
                             (d->D == 256 || d->D == 384 || d->D == 768 || d->D == 1024) && d->in_ld % 4 == 0 &&
                             d->out_ld % 4 == 0 && aligned16(d->x) && aligned16(d->out) && d->out_plane % 4 == 0,
                         "layernorm: bad descriptor (D=%d)", d->D))
VT_SIMPLE_ADD(vt_program_add_imgstats, ImgStatsOp, vt_imgstats_desc,
              VT_REQUIRE(d->img && d->partial && d->flags && d->count >= 1 && (d->dtype == VT_U8 || d->dtype == VT_F32) &&
                             aligned16(d->img) && (reinterpret_cast<uintptr_t>(d->partial) & 7) == 0,
                         "imgstats: bad descriptor"))
VT_SIMPLE_ADD(vt_program_add_patchify, PatchifyOp, vt_patchify_desc,
              VT_REQUIRE(d->img && d->out && d->flags && d->images >= 1 && d->patch >= 1 && d->H >= d->patch &&
                             d->W >= d->patch && d->out_cols >= 3 * d->patch * d->patch && d->out_ld >= d->out_cols &&
                             d->out_cols % 8 == 0 && d->out_ld % 8 == 0 && aligned16(d->out) &&
                             (d->dtype == VT_U8 || d->dtype == VT_F32),
                         "patchify: bad descriptor"))
VT_SIMPLE_ADD(vt_program_add_cls, ClsOp, vt_cls_desc, VT_REQUIRE(d->cls && d->pos && d->h, "cls: bad descriptor"))
VT_SIMPLE_ADD(vt_program_add_pack, PackOp, vt_pack_desc,
              VT_REQUIRE(d->src && d->out && d->rows >= 1 && d->cols >= 1, "pack: bad descriptor"))
VT_SIMPLE_ADD(vt_program_add_affine, AffineOp, vt_affine_desc,
              VT_REQUIRE(d->x && (d->out || d->xpad) && d->mins && d->maxs && d->rows >= 1 && d->A >= 1, "affine: bad descriptor"))
VT_SIMPLE_ADD(vt_program_add_tembed, TembedOp, vt_tembed_desc,
              VT_REQUIRE(d->t && d->out && d->rows >= 1 && d->dim >= 4 && d->dim % 2 == 0, "tembed: bad descriptor"))
VT_SIMPLE_ADD(vt_program_add_sde, SdeOp, vt_sde_desc,
              VT_REQUIRE(d->x && d->v && d->s && d->rows >= 1 && d->A >= 1, "sde: bad descriptor"))
int vt_program_add_lstm(vt_program* p, const vt_lstm_desc* d) {
  if (!p || !d) return fail(VT_E_INVALID, "null argument");
  VT_REQUIRE(d->xw && d->w_hh && d->h && d->c && d->y && d->B >= 1 && d->T >= 1 && d->H == 256, "lstm: bad descriptor");
  std::unique_ptr<LstmOp> op(new LstmOp());
  op->d = *d;
  if (d->w_hh_tc && d->h_tc && d->zero_init && d->B >= LSTM_TC_MIN_BATCH && d->y_plane == 0) {
    int rc = build_lstm_tc(&op->tc, d->w_hh_tc, d->h_tc, d->xw, d->y, d->y_dtype, d->y_ld, nullptr, nullptr, d->h, d->c, d->B, d->T);
    if (rc) return rc;
    op->use_tc = true;
  }
  p->ops.push_back(std::move(op));
  return VT_OK;
}

VT_SIMPLE_ADD(vt_program_add_tcol, TcolOp, vt_tcol_desc,
              VT_REQUIRE(d->src && d->out && (d->src_dtype == VT_BF16 || d->src_dtype == VT_F32) && d->G >= 1 && d->B >= 1 &&
                             d->T_src >= 1 && d->C >= 1 && d->taps >= 1 && d->taps <= VT_MAX_TAPS && d->stride >= 1 &&
                             d->t_out >= 1 && d->c_pad >= d->C && d->k_ld >= (int64_t)d->B * d->t_out && d->k_ld % 64 == 0 &&
                             (int64_t)d->G * d->taps <= 65535,
                         "tcol: bad descriptor"))
VT_SIMPLE_ADD(vt_program_add_gnbwd, GnbwdOp, vt_gnbwd_desc,
              VT_REQUIRE(d->raw && d->dout && d->gamma && d->beta && d->draw && d->part && d->G >= 1 && d->B >= 1 && d->T >= 1 &&
                             d->C >= 1 && d->C <= 256 * vt::GNB_MAX_CPT && d->groups >= 1 && d->groups <= 64 &&
                             d->C % d->groups == 0 && d->dout_ld >= d->C && (!d->dfilm || d->film),
                         "gnbwd: bad descriptor (C=%d groups=%d)", d->C, d->groups))
VT_SIMPLE_ADD(vt_program_add_colsum, ColsumOp, vt_colsum_desc,
              VT_REQUIRE(d->x && d->out && d->G >= 1 && d->G <= 65535 && d->rows >= 1 && d->C >= 1 && d->ld >= d->C,
                         "colsum: bad descriptor"))

VT_SIMPLE_ADD(vt_program_add_ewise, EwiseOp, vt_ewise_desc,
              VT_REQUIRE(d->a && d->b && d->out && d->rows >= 1 && d->cols >= 1 && d->a_ld >= d->cols && d->b_ld >= d->cols &&
                             d->out_ld >= d->cols && d->op >= VT_EW_ADD && d->op <= VT_EW_SCALED_DIFF,
                         "ewise: bad descriptor"))

VT_SIMPLE_ADD(vt_program_add_silossbwd, SilossBwdOp, vt_silossbwd_desc,
              VT_REQUIRE(d->bvs && d->x0 && d->x1 && d->z_unit && d->tclip && d->dvs && d->B >= 1 && d->n >= 1,
                         "silossbwd: bad descriptor"))

VT_SIMPLE_ADD(vt_program_add_dropmask, DropmaskOp, vt_dropmask_desc,
              VT_REQUIRE(d->mask && d->n >= 1 && d->p >= 0.f && d->p < 1.f, "dropmask: bad descriptor (p=%f)", (double)d->p))
VT_SIMPLE_ADD(vt_program_add_lngelubwd, LnGeluBwdOp, vt_lngelubwd_desc,
              VT_REQUIRE(d->z0 && d->dzn && d->gamma && d->beta && d->dz0 && d->d1 && d->d1zh && d->rows >= 1 && d->D == 256,
                         "lngelubwd: bad descriptor (D=%d)", d->D))
int vt_program_add_lstm_train(vt_program* p, const vt_lstm_train_desc* d) {
  if (!p || !d) return fail(VT_E_INVALID, "null argument");
  VT_REQUIRE(d->xw && d->w_hh && d->y && d->gates && d->c && d->B >= 1 && d->T >= 1 && d->H == 256 &&
                 (d->y_dtype == VT_BF16 || d->y_dtype == VT_F32) && d->y_ld >= d->H, "lstm_train: bad descriptor");
  std::unique_ptr<LstmTrainOp> op(new LstmTrainOp());
  op->d = *d;
  if (d->w_hh_tc && d->h_tc && d->B >= LSTM_TC_MIN_BATCH) {
    int rc = build_lstm_tc(&op->tc, d->w_hh_tc, d->h_tc, d->xw, d->y, d->y_dtype, d->y_ld, d->gates, d->c, nullptr, nullptr, d->B, d->T);
    if (rc) return rc;
    op->use_tc = true;
  }
  p->ops.push_back(std::move(op));
  return VT_OK;
}
int vt_program_add_lstm_bwd(vt_program* p, const vt_lstm_bwd_desc* d) {
  if (!p || !d) return fail(VT_E_INVALID, "null argument");
  VT_REQUIRE(d->gates && d->c && d->dy && d->w_hh && d->dgates && d->B >= 1 && d->T >= 1 && d->H == 256 && d->dy_ld >= d->H,
             "lstm_bwd: bad descriptor");
  std::unique_ptr<LstmBwdOp> op(new LstmBwdOp());
  op->d = *d;
  if (d->w_hh_t_tc && d->dg_tc && d->B >= LSTM_TC_MIN_BATCH && d->dy_ld % 4 == 0 && aligned16(d->dy)) {
    const int H = vt::LTC_H;
    vt::LstmBwdTcArgs& a = op->tc;
    memset(&a, 0, sizeof(a));
    {
      const uint64_t dims[2] = {(uint64_t)4 * H, (uint64_t)H};
      const uint64_t st[1] = {(uint64_t)4 * H * 2};
      const uint32_t box[2] = {64u, (uint32_t)vt::LTC_U};
      int rc = make_tmap(&a.tmW, VT_BF16, 2, d->w_hh_t_tc, dims, st, box);
      if (rc) return rc;
    }
    {
      const uint64_t dims[3] = {(uint64_t)4 * H, (uint64_t)d->T, (uint64_t)d->B};
      const uint64_t st[2] = {(uint64_t)4 * H * 2, (uint64_t)d->T * 4 * H * 2};
      const uint32_t box[3] = {64u, 1u, 128u};
      int rc = make_tmap(&a.tmD, VT_BF16, 3, d->dg_tc, dims, st, box);
      if (rc) return rc;
    }
    a.gates = d->gates; a.c_all = d->c; a.dy = d->dy; a.dy_ld = d->dy_ld; a.dgates = d->dgates;
    a.dgb = reinterpret_cast<__nv_bfloat16*>(d->dg_tc);
    a.B = d->B; a.T = d->T;
    op->use_tc = true;
  }
  p->ops.push_back(std::move(op));
  return VT_OK;
}

VT_SIMPLE_ADD(vt_program_add_qsample, QsampleOp, vt_qsample_desc,
              VT_REQUIRE(d->x0 && d->x1 && d->step && d->z_unit && d->xt && d->tclip && d->B >= 1 && d->n >= 1 && d->A >= 1 &&
                             d->n % d->A == 0, "qsample: bad descriptor"))
VT_SIMPLE_ADD(vt_program_add_siloss, SilossOp, vt_siloss_desc,
              VT_REQUIRE(d->bvs && d->x0 && d->x1 && d->z_unit && d->tclip && d->per_sample && d->out && d->B >= 1 && d->n >= 1,
                         "siloss: bad descriptor"))

int vt_program_run(vt_program* p, int first, int count, void* stream) {
  int rc = clamp_range(p, first, &count);
  if (rc) return rc;
  cudaStream_t s = reinterpret_cast<cudaStream_t>(stream);
  for (int i = first; i < first + count; ++i) {
    rc = p->ops[i]->launch(s);
    if (rc) return rc;
  }
  return VT_OK;
}

int vt_program_graph_build(vt_program* p, int first, int count) {
  int rc = clamp_range(p, first, &count);
  if (rc) return rc;
  if (!p->capture_stream) VT_CUDA(cudaStreamCreateWithFlags(&p->capture_stream, cudaStreamNonBlocking));
  if (p->graph_exec) {
    cudaGraphExecDestroy(p->graph_exec);
    p->graph_exec = nullptr;
  }
  // function attributes cannot be set during capture: run the range once eagerly first
  rc = vt_program_run(p, first, count, p->capture_stream);
  if (rc) return rc;
  VT_CUDA(cudaStreamSynchronize(p->capture_stream));
  VT_CUDA(cudaStreamBeginCapture(p->capture_stream, cudaStreamCaptureModeThreadLocal));
  rc = vt_program_run(p, first, count, p->capture_stream);
  cudaGraph_t graph = nullptr;
  cudaError_t e = cudaStreamEndCapture(p->capture_stream, &graph);
  if (rc) {
    if (graph) cudaGraphDestroy(graph);
    return rc;
  }
  if (e != cudaSuccess) return fail(VT_E_CUDA, "cudaStreamEndCapture: %s", cudaGetErrorString(e));
  e = cudaGraphInstantiate(&p->graph_exec, graph, 0);
  cudaGraphDestroy(graph);
  if (e != cudaSuccess) return fail(VT_E_CUDA, "cudaGraphInstantiate: %s", cudaGetErrorString(e));
  return VT_OK;
}

int vt_program_graph_launch(vt_program* p, void* stream) {
  if (!p || !p->graph_exec) return fail(VT_E_INVALID, "no graph built");
  VT_CUDA(cudaGraphLaunch(p->graph_exec, reinterpret_cast<cudaStream_t>(stream)));
  return VT_OK;
}

int vt_adamw_ema_step(const vt_adamw_desc* d, void* stream) {
  VT_REQUIRE(d && d->tensors && d->chunks && d->n_chunks >= 1 && d->chunk_elems >= 1, "adamw: bad descriptor");
  static_assert(sizeof(vt_opt_tensor) == sizeof(vt::OptTensor), "record layout");
  vt::adamw_ema_kernel<<<d->n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const vt::OptTensor*>(d->tensors), reinterpret_cast<const long long*>(d->chunks), d->chunk_elems, d->lr,
      d->beta1, d->beta2, d->eps, d->weight_decay, d->bias_corr1, d->bias_corr2, d->ema_decay, d->grad_scale);
  VT_LAUNCH_CHECK("adamw_ema_kernel");
  return VT_OK;
}

int vt_gather_repack(const void* recs_dev, const int64_t* chunks_dev, int32_t n_chunks, const float* arena_dev, void* stream) {
  VT_REQUIRE(recs_dev && chunks_dev && arena_dev && n_chunks >= 1, "gather_repack: bad arguments");
  vt::gather_repack_kernel<<<n_chunks, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      reinterpret_cast<const vt::GatherRec*>(recs_dev), reinterpret_cast<const long long*>(chunks_dev), arena_dev);
  VT_LAUNCH_CHECK("gather_repack_kernel");
  return VT_OK;
}

// ---- pad_and_resize_for_siglip (scripts/utils_eef.py:44-77) ----
namespace {
struct ResizeTab {
  int side = 0, target = 0;
  int* off = nullptr;
  int* si = nullptr;
  float* alpha = nullptr;
};
std::vector<ResizeTab> g_resize_tabs;   // one per (canvas side, target) seen; a handful in practice

// OpenCV computeResizeAreaTab (imgproc/resize.cpp), double arithmetic, weights stored as float
int resize_area_tab(int side, int target, ResizeTab* out) {
  for (const ResizeTab& t : g_resize_tabs)
    if (t.side == side && t.target == target) {
      *out = t;
      return VT_OK;
    }
  const double scale = (double)side / target;
  std::vector<int> off(target + 1, 0), si;
  std::vector<float> al;
  for (int dx = 0; dx < target; ++dx) {
    off[dx] = (int)si.size();
    const double fsx1 = dx * scale, fsx2 = fsx1 + scale;
    const double cell = std::min(scale, side - fsx1);
    int sx1 = (int)std::ceil(fsx1), sx2 = (int)std::floor(fsx2);
    sx2 = std::min(sx2, side - 1);
    sx1 = std::min(sx1, sx2);
    if (sx1 - fsx1 > 1e-3) {
      si.push_back(sx1 - 1);
      al.push_back((float)((sx1 - fsx1) / cell));
    }
    for (int sx = sx1; sx < sx2; ++sx) {
      si.push_back(sx);
      al.push_back((float)(1.0 / cell));
    }
    if (fsx2 - sx2 > 1e-3) {
      si.push_back(sx2);
      al.push_back((float)(std::min(std::min(fsx2 - sx2, 1.0), cell) / cell));
    }
  }
  off[target] = (int)si.size();
  ResizeTab t;
  t.side = side;
  t.target = target;
  VT_CUDA(cudaMalloc(&t.off, off.size() * sizeof(int)));
  VT_CUDA(cudaMalloc(&t.si, si.size() * sizeof(int)));
  VT_CUDA(cudaMalloc(&t.alpha, al.size() * sizeof(float)));
  VT_CUDA(cudaMemcpy(t.off, off.data(), off.size() * sizeof(int), cudaMemcpyHostToDevice));
  VT_CUDA(cudaMemcpy(t.si, si.data(), si.size() * sizeof(int), cudaMemcpyHostToDevice));
  VT_CUDA(cudaMemcpy(t.alpha, al.data(), al.size() * sizeof(float), cudaMemcpyHostToDevice));
  g_resize_tabs.push_back(t);
  *out = t;
  return VT_OK;
}
}  // namespace

int vt_pad_resize_area(const uint8_t* src_dev, int32_t n, int32_t h, int32_t w, int32_t c, uint8_t* dst_dev, int32_t target, void* stream) {
  VT_REQUIRE(src_dev && dst_dev && n >= 1 && h >= 1 && w >= 1 && c >= 1 && c <= 4 && target >= 1, "pad_resize_area: bad arguments");
  const int side = h > w ? h : w;
  VT_REQUIRE(side >= target, "pad_resize_area: up-scaling (canvas %d < target %d) is a different INTER_AREA code path and is not built", side, target);
  vt::ResizeArgs a;
  memset(&a, 0, sizeof(a));
  a.src = src_dev; a.dst = dst_dev; a.n = n; a.h = h; a.w = w; a.c = c; a.target = target;
  a.side = side; a.pad_y = (side - h) / 2; a.pad_x = (side - w) / 2;
  if (side % target == 0) {
    a.iscale = side / target;
    a.inv_area = 1.f / (float)(a.iscale * a.iscale);
  } else {
    ResizeTab t;
    int rc = resize_area_tab(side, target, &t);
    if (rc) return rc;
    a.tab_off = t.off; a.tab_si = t.si; a.tab_alpha = t.alpha;
  }
  const long long total = (long long)n * target * target * c;
  vt::pad_resize_area_kernel<<<grid_for(total, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a);
  VT_LAUNCH_CHECK("pad_resize_area_kernel");
  return VT_OK;
}

int vt_batch_gather(const vt_batch_gather_desc* d, void* stream) {
  VT_REQUIRE(d && d->qpos && d->grip_scaled && d->vla && d->vla_last_scaled && d->start, "batch_gather: missing store arrays");
  VT_REQUIRE(d->B >= 1 && d->A >= 1 && d->context_frames >= 1 && d->horizon >= 1 && d->horizon <= d->vla_T,
             "batch_gather: bad shape (B %d, A %d, context %d, horizon %d, chunk length %d)", d->B, d->A, d->context_frames, d->horizon, d->vla_T);
  VT_REQUIRE(!d->forces_out || (d->forces && d->Fd >= 1), "batch_gather: forces requested but the store has none");
  VT_REQUIRE(!d->disps_out || (d->disps && d->Dd >= 1), "batch_gather: displacements requested but the store has none");
  VT_REQUIRE(!d->feats || (d->frame_mean && d->feat_cam1 && d->feat_cam2 && d->D >= 1), "batch_gather: feature cache needs frame_mean, feat_cam1, feat_cam2");
  VT_REQUIRE(!(d->expert_n || d->vla_n) || (d->action_mins && d->action_maxs && d->vla_mins && d->vla_maxs), "batch_gather: normalised outputs need the four stats vectors");
  vt::batch_gather_kernel<<<d->B, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(*d);
  VT_LAUNCH_CHECK("batch_gather_kernel");
  return VT_OK;
}

int vt_chunk_handoff(const void* action_dev, int32_t dtype, int32_t B, int32_t N, int32_t S, const int32_t* idx_dev, const float* scale_dev,
                     int32_t A, float last_div, float* raw_dev, float* chunk_dev, int32_t T_exec, void* stream) {
  VT_REQUIRE(action_dev && idx_dev && scale_dev && (raw_dev || chunk_dev), "chunk_handoff: missing pointers");
  VT_REQUIRE(dtype == VT_BF16 || dtype == VT_F32, "chunk_handoff: the action vector must be bf16 or fp32");
  VT_REQUIRE(B >= 1 && N >= 1 && A >= 1 && S >= A && T_exec >= 0 && T_exec <= N, "chunk_handoff: bad shape (B %d, N %d, S %d, A %d, T_exec %d)", B, N, S, A, T_exec);
  VT_REQUIRE(last_div != 0.f, "chunk_handoff: last_div is zero");
  vt::chunk_handoff_kernel<<<grid_for((long long)B * N * A, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      action_dev, dtype, B, N, S, idx_dev, scale_dev, A, last_div, raw_dev, chunk_dev, T_exec);
  VT_LAUNCH_CHECK("chunk_handoff_kernel");
  return VT_OK;
}

int vt_pos_embed_resize(const float* src_dev, int32_t s, float* dst_dev, int32_t nh, int32_t nw, int32_t D, void* stream) {
  VT_REQUIRE(src_dev && dst_dev && s >= 1 && nh >= 1 && nw >= 1 && D >= 1, "pos_embed_resize: bad arguments");
  vt::pos_resize_kernel<<<grid_for((long long)nh * nw * D, 256), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src_dev, s, dst_dev, nh, nw, D);
  VT_LAUNCH_CHECK("pos_resize_kernel");
  return VT_OK;
}

}  // extern "C"
