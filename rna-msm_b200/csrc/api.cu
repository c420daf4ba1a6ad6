// C ABI of librnamsm_b200.so (declared in include/rnamsm_b200.h) and the host-side drivers that
// chain the kernels into one AxialTransformerLayer (modules.py:242-267) / one MSATransformer
// forward (model.py:338-416).  No allocation, no synchronisation: everything is enqueued on the
// caller's stream against caller-owned buffers.
#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdlib.h>

#include <atomic>
#include <cmath>
#include <cstring>
#include <vector>

#include "../../include/rnamsm_b200.h"
#include "common.cuh"
#include "launch.h"

namespace rnamsm {

static thread_local char g_err[1024] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
const char* get_error() { return g_err; }
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }
long long launch_count() { return g_launches.load(std::memory_order_relaxed); }

// ---- optional kernel-class timing ---------------------------------------------------------------
struct ProfEvent { int cls; cudaEvent_t a, b; };
static bool g_prof_on = false;
static std::vector<ProfEvent> g_prof_events;
static std::vector<cudaEvent_t> g_prof_pool;
static const char* kClassNames[KC_COUNT] = {"embed_ln", "layernorm", "row_softmax", "vocab_proj", "linear_qkv",
                                            "linear_fc1_gelu", "linear_out_resid", "linear_fc2_resid",
                                            "row_logits", "row_av", "col_attn", "contact_head", "row_attn_short"};
static cudaEvent_t prof_get_event() {
  if (!g_prof_pool.empty()) { cudaEvent_t e = g_prof_pool.back(); g_prof_pool.pop_back(); return e; }
  cudaEvent_t e = nullptr;
  cudaEventCreate(&e);
  return e;
}
ProfScope::ProfScope(int c, cudaStream_t s) : cls(c), st(s), slot(nullptr) {
  if (!g_prof_on) return;
  ProfEvent ev{c, prof_get_event(), prof_get_event()};
  cudaEventRecord(ev.a, st);
  g_prof_events.push_back(ev);
  slot = reinterpret_cast<void*>(g_prof_events.size());
}
ProfScope::~ProfScope() {
  if (!slot) return;
  cudaEventRecord(g_prof_events[reinterpret_cast<size_t>(slot) - 1].b, st);
}

// RNAMSM_PDL=1 launches the forward's kernels with programmatic dependent launch (each kernel's prologue overlaps its
// predecessor's tail; all parity tests pass with it).  OFF by default: measured no gain on a B200 -- 512 x 36 (132
// launches of ~40 us): 5.358 ms plain vs 5.352 ms; 512 x 256: within the box-to-box clock noise -- back-to-back launches
// on one stream already pipeline, and a persistent 230 KiB CTA cannot become resident before its predecessor's CTA on
// that SM has exited.
bool pdl_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RNAMSM_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

int encode_tmap(CUtensorMap* map, int elem, const void* base, int rank, const uint64_t* dims,
                const uint64_t* strides_bytes, const uint32_t* box) {
  static PFN_cuTensorMapEncodeTiled_v12000 encode = nullptr;
  if (!encode) {
    void* fn = nullptr;
    cudaDriverEntryPointQueryResult qres;
    cudaError_t e = cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &fn, cudaEnableDefault, &qres);
    if (e != cudaSuccess || qres != cudaDriverEntryPointSuccess || !fn) {
      set_error("cuTensorMapEncodeTiled entry point unavailable (%s)", cudaGetErrorString(e));
      return 1;
    }
    encode = reinterpret_cast<PFN_cuTensorMapEncodeTiled_v12000>(fn);
  }
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) {
    set_error("TMA base pointer %p is not 16-byte aligned", base);
    return 1;
  }
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bx[5];
  cuuint32_t estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bx[i] = box[i];
    estr[i] = 1;
    if (i > 0) {
      gstr[i - 1] = strides_bytes[i - 1];
      if (gstr[i - 1] % 16 != 0) {
        set_error("TMA stride %llu (dim %d) is not a multiple of 16 bytes", (unsigned long long)gstr[i - 1], i);
        return 1;
      }
    }
  }
  const CUtensorMapDataType dt = elem == TMAP_F32 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32
                                 : elem == TMAP_F16 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  CUresult r = encode(map, dt, (cuuint32_t)rank, const_cast<void*>(base), gdim, gstr, bx,
                      estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                      CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed with CUresult %d (rank %d dims %llu,%llu,%llu box %u,%u,%u)", (int)r, rank,
              (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
              (unsigned long long)(rank > 2 ? dims[2] : 0), box[0], rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0);
    return 1;
  }
  return 0;
}

// ---------------------------------------------------------------------------------------------
// Workspace plan for one R x C MSA (bytes, every region 256 B aligned):
//   xn      [T, D]      compute dtype   LayerNorm output feeding the next GEMM
//   qkvh    [T, 4D|F]   compute dtype   q|k|v (3D) + attention context (D); aliased by the FFN hidden
//   partial [S, H, C, C] fp32           split-K tied logits
//   probs   [H, C, ldp] compute dtype   softmax probabilities for the AV GEMM (16-bit paths only)
//   map     [H, C, C]   fp32            scratch attention map when the caller does not want it
//   cnt     [2 T / 256] int32           per (m-block, CTA rank) arrival counters of the fused residual + LayerNorm GEMM
//   split   fp32 path only              hi / lo halves of the current GEMM's activation and weight (tf32 x 3 product)
// ---------------------------------------------------------------------------------------------
static size_t tf32_scratch_bytes(long long M, int N, int K);

struct Plan {
  size_t el;  // bytes per element of the compute dtype
  int splits, ldp;
  int short_chunks;   // > 0: the tied row attention runs as ONE launch (row_attn_short.cu) with this many row chunks = splits
  size_t off_xn, off_qkvh, off_partial, off_probs, off_map, off_cnt, off_split, total;
};

static size_t align256(size_t x) { return (x + 255) & ~(size_t)255; }

static inline bool is16(int dtype) { return dtype == RNAMSM_BF16 || dtype == RNAMSM_F16; }

static int pick_splits(int R, int C, int H, int dtype) {
  if (is16(dtype)) return row_logits_splits_16(R, C, H);
  const long long tiles = (long long)H * ceil_div(C, 128) * ceil_div(C, 128);
  int want = (int)std::max<long long>(1, (2 * 148) / std::max<long long>(1, tiles));
  want = std::max(1, std::min(want, std::max(1, R / 4)));
  const int rps = ceil_div(R, want);
  return ceil_div(R, rps);
}

static Plan make_plan(int R, int C, int D, int H, int F, int dtype, bool f32_tensor = false) {
  Plan p{};
  const size_t T = (size_t)R * C;
  p.el = is16(dtype) ? 2 : 4;
  p.short_chunks = is16(dtype) ? row_attn_short_chunks(R, C, H) : 0;
  p.splits = p.short_chunks > 0 ? p.short_chunks : pick_splits(R, C, H, dtype);
  p.ldp = is16(dtype) ? (C + 7) / 8 * 8 : C;
  size_t o = 0;
  p.off_xn = o;      o += align256(T * D * p.el);
  p.off_qkvh = o;    o += align256(T * (size_t)std::max(4 * D, F) * p.el);
  p.off_partial = o; o += align256((size_t)p.splits * H * C * C * 4);
  p.off_probs = o;   o += align256(is16(dtype) ? (size_t)H * C * p.ldp * p.el : 0);
  p.off_map = o;     o += align256((size_t)H * C * C * 4);
  p.off_cnt = o;     o += align256(ln_counter_bytes((long long)T));   // arrival counters of the fused LayerNorm epilogue
  p.off_split = o;   o += (dtype == RNAMSM_F32 && f32_tensor) ? tf32_scratch_bytes((long long)T, std::max(3 * D, F), std::max(D, F)) : 0;
  p.total = o;
  return p;
}

// fp32 path with RNAMSM_F32_TENSOR: the nn.Linear layers run on the tensor cores as a three-term tf32 product
// (umma_gemm.cu, kTf32) on hi / lo halves of both operands staged in the workspace's split region.
static size_t tf32_scratch_bytes(long long M, int N, int K) {
  return 2 * align256((size_t)M * K * 4) + 2 * align256((size_t)N * K * 4);
}
static int linear_tf32(const float* x, const float* W, long long M, int N, int K, const LinearEpilogue& e, float* out,
                       uint8_t* scratch, cudaStream_t st) {
  float* xh = reinterpret_cast<float*>(scratch);
  float* xl = reinterpret_cast<float*>(scratch + align256((size_t)M * K * 4));
  float* wh = reinterpret_cast<float*>(scratch + 2 * align256((size_t)M * K * 4));
  float* wl = reinterpret_cast<float*>(scratch + 2 * align256((size_t)M * K * 4) + align256((size_t)N * K * 4));
  int rc;
  if ((rc = launch_split_tf32(x, xh, xl, M * K, st))) return rc;
  if ((rc = launch_split_tf32(W, wh, wl, (long long)N * K, st))) return rc;
  return launch_linear_tf32(xh, xl, wh, wl, M, N, K, e, out, st);
}

static int linear_any(const void* x, const void* W, long long M, int N, int K, int dtype, const LinearEpilogue& e,
                      void* out, cudaStream_t st, const LnFuse* ln = nullptr, uint8_t* tf32_scratch = nullptr) {
  if (is16(dtype)) return launch_linear_16(x, W, M, N, K, dtype == RNAMSM_F16, e, out, st, ln);
  RNAMSM_REQUIRE(ln == nullptr, "linear: LayerNorm fusion exists in the 16-bit path only");
  if (dtype == RNAMSM_F32 && tf32_scratch != nullptr && N % 32 == 0 && K % 32 == 0)
    return linear_tf32((const float*)x, (const float*)W, M, N, K, e, (float*)out, tf32_scratch, st);
  if (dtype == RNAMSM_F32)
    return launch_linear_f32((const float*)x, (const float*)W, M, N, K, e, (float*)out, st);
  set_error("unknown dtype %d", dtype);
  return 2;
}

static int block_dtype(int requested, int dtype) {
  // per-block operand type: 0 (unset) inherits the layer dtype; the fp32 path is all-fp32
  if (dtype == RNAMSM_F32) return RNAMSM_F32;
  return is16(requested) ? requested : dtype;
}

// RNAMSM_FUSE_LN=1 routes every LayerNorm behind a residual GEMM through that GEMM's fused epilogue (umma_gemm.cu, kLN).
// OFF by default: the fused kernel is exact (bit-identical rows, tests/test_gpu_ops.py) but SLOWER on a B200 -- measured
// at 131072 x 768: residual GEMM 176 us + LayerNorm 94 us = 270 us un-fused vs 804 us fused (K = 768), 598 vs 935 us
// (K = 3072); the handshake itself is free (178 us with the row work skipped).  The 4 LayerNorm warps of a CTA keep only
// 8 rows (24 KiB) in flight while the GEMM's own TMA traffic holds HBM at 85 % of its peak, so each dependent
// load round trip takes ~5 us and the re-read runs at ~0.7 TB/s; hiding that latency needs ~90 KiB of rows in flight per
// SM (Little: 2.7 TB/s x 5 us / 148), which neither the register file nor the shared memory left beside the 6-stage
// operand ring can hold.  The stand-alone pass, with 64 warps per SM in flight, runs at 98 % of the HBM roofline.
static bool fuse_ln_enabled() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("RNAMSM_FUSE_LN");
    v = (e && e[0] == '1') ? 1 : 0;
  }
  return v == 1;
}

// Debug range watch (rnamsm_debug_range_watch): when set, every 16-bit activation the layer writes (LayerNorm outputs,
// q|k|v, attention contexts, the post-GELU hidden) is scanned for values at the edge of its type's range.
static unsigned long long* g_range_counters = nullptr;
static int range_watch(const void* p, long long n, int dtype, cudaStream_t st) {
  if (g_range_counters == nullptr || !is16(dtype)) return 0;
  return launch_range_scan(p, n, dtype, g_range_counters, st);
}

// LayerNorm of the NEXT layer's row block, produced by this layer's fc2 epilogue (16-bit path)
struct NextLn { const float* w; const float* b; int dtype; };

// xn_ready: the workspace's xn region already holds LayerNorm_row(x) (written by the previous layer's fc2 epilogue).
static int layer_forward(const rnamsm_layer_weights* w, int D, int H, int F, float eps, float* x, int R, int C,
                         const uint8_t* pad, int dtype, float* row_probs_out, uint8_t* ws, const Plan& p,
                         cudaStream_t st, bool xn_ready = false, const NextLn* next = nullptr, bool f32_tensor = false) {
  const long long T = (long long)R * C;
  void* xn = ws + p.off_xn;
  uint8_t* qkv = ws + p.off_qkvh;
  void* ctx = qkv + (size_t)T * 3 * D * p.el;
  float* partial = reinterpret_cast<float*>(ws + p.off_partial);
  float* map = row_probs_out ? row_probs_out : reinterpret_cast<float*>(ws + p.off_map);
  const int row_dt = block_dtype(w->row.dtype, dtype);   // tied row attention: fp16 by default in the 16-bit path
  const int col_dt = block_dtype(w->col.dtype, dtype);
  void* probs_lp = is16(row_dt) ? (void*)(ws + p.off_probs) : nullptr;
  int rc;

  // With RNAMSM_FUSE_LN=1 every residual GEMM of the 16-bit path also emits the LayerNorm its successor needs
  // (umma_gemm.cu, kLN) and the stand-alone LayerNorm passes disappear except the very first one (see fuse_ln_enabled()
  // for why this is not the default).
  const bool fuse = is16(dtype) && fuse_ln_enabled();
  uint8_t* split = (dtype == RNAMSM_F32 && f32_tensor) ? ws + p.off_split : nullptr;   // operand halves of the tf32 x 3 product
  int* cnt = reinterpret_cast<int*>(ws + p.off_cnt);
  // the counters are zero between launches (the kernel resets them); a layer that starts a chain zeroes them once
  if (fuse && !xn_ready) RNAMSM_CHECK_CUDA(cudaMemsetAsync(cnt, 0, ln_counter_bytes(T), st));
  RNAMSM_REQUIRE(!xn_ready || fuse, "layer_forward: xn_ready needs the 16-bit path with LayerNorm fusion enabled");
  RNAMSM_REQUIRE(next == nullptr || fuse, "layer_forward: next-layer LayerNorm needs the 16-bit path with fusion enabled");

  // ---- tied row attention: x += out_proj(AV(softmax(sum_r q k^T)))   modules.py:385-401, 802-821
  if (!xn_ready && (rc = launch_layernorm(x, w->row.ln_w, w->row.ln_b, xn, row_dt, T, D, eps, st))) return rc;
  // align_scaling (modules.py:713-715) = 64^-1/2 / sqrt(R).  fp32 path: applied to q as the reference
  // does.  16-bit path: q carries the exact power of two 64^-1/2 and the 1/sqrt(R) factor is applied
  // to the fp32 logit sums in the softmax kernel (keeps q well inside the 16-bit normal range).
  const float q_scale = is16(row_dt) ? 0.125f : 0.125f / sqrtf((float)R);
  const float logit_scale = is16(row_dt) ? 1.0f / sqrtf((float)R) : 1.0f;
  {
    LinearEpilogue e{RNAMSM_EPI_BIAS, w->row.b_qkv, q_scale, D, pad};
    if ((rc = linear_any(xn, w->row.w_qkv, T, 3 * D, D, row_dt, e, qkv, st, nullptr, split))) return rc;
    if ((rc = range_watch(xn, T * D, row_dt, st)) || (rc = range_watch(qkv, T * 3 * D, row_dt, st))) return rc;
  }
  // key mask comes from MSA row 0 (padding_mask[:, 0], modules.py:780-784) = first C entries of pad
  if (is16(row_dt) && p.short_chunks > 0) {
    // short alignment (C <= 128): logits -> softmax -> A V in one cooperative launch (row_attn_short.cu)
    if ((rc = launch_row_attn_short_16(qkv, R, C, H, row_dt == RNAMSM_F16, partial, p.short_chunks, pad, logit_scale, map,
                                       probs_lp, p.ldp, ctx, st)))
      return rc;
  } else {
    if (is16(row_dt)) {
      if ((rc = launch_row_logits_16(qkv, R, C, H, row_dt == RNAMSM_F16, partial, p.splits, st))) return rc;
    } else {
      if ((rc = launch_row_logits_f32((const float*)qkv, R, C, H, partial, p.splits, st))) return rc;
    }
    if ((rc = launch_row_softmax(partial, p.splits, H, C, pad, logit_scale, map, probs_lp, p.ldp, row_dt, st))) return rc;
    if (is16(row_dt)) {
      if ((rc = launch_row_av_16(probs_lp, p.ldp, qkv, R, C, H, row_dt == RNAMSM_F16, ctx, st))) return rc;
    } else {
      if ((rc = launch_row_av_f32(map, C, (const float*)qkv, R, C, H, (float*)ctx, st))) return rc;
    }
  }
  // ---- column attention over the MSA depth                           modules.py:875-945
  // 16-bit path: LayerNorm writes its output in column-major token order (c * R + r), so the QKV GEMM
  // produces q|k|v as [C, R, 3D] and the flash kernel's K/V boxes read rows 3D elements apart instead of
  // C * 3D (one TMA row per 2 MiB page otherwise).  ctx comes back token-major for the out-projection.
  const int col_major = is16(col_dt) && R > 1 ? 1 : 0;
  {
    LinearEpilogue e{RNAMSM_EPI_BIAS_RESIDUAL, w->row.b_out, 1.f, 0, nullptr};
    const LnFuse ln{w->col.ln_w, w->col.ln_b, eps, xn, col_dt == RNAMSM_F16, col_major ? R : 0, col_major ? C : 0, cnt};
    if ((rc = range_watch(ctx, T * D, row_dt, st))) return rc;
    if ((rc = linear_any(ctx, w->row.w_out, T, D, D, row_dt, e, x, st, fuse ? &ln : nullptr, split))) return rc;
  }
  if (!fuse &&
      (rc = launch_layernorm(x, w->col.ln_w, w->col.ln_b, xn, col_dt, T, D, eps, st, col_major ? R : 0, col_major ? C : 0)))
    return rc;
  if (R == 1) {
    // single-row shortcut: out_proj(v_proj(x)), modules.py:882-894.  Project with the v rows only.
    LinearEpilogue ev{RNAMSM_EPI_BIAS, w->col.b_qkv + 2 * D, 1.f, 0, nullptr};
    const void* wv = (const uint8_t*)w->col.w_qkv + (size_t)2 * D * D * p.el;
    if ((rc = linear_any(xn, wv, T, D, D, col_dt, ev, ctx, st, nullptr, split))) return rc;
  } else {
    LinearEpilogue e{RNAMSM_EPI_BIAS, w->col.b_qkv, 1.0f / sqrtf(64.f), D, nullptr};  // q *= scaling, :905
    if ((rc = linear_any(xn, w->col.w_qkv, T, 3 * D, D, col_dt, e, qkv, st, nullptr, split))) return rc;
    if (is16(col_dt)) {
      if ((rc = launch_col_attn_16(qkv, R, C, H, col_dt == RNAMSM_F16, col_major, pad, ctx, st))) return rc;
    } else {
      if ((rc = launch_col_attn_f32((const float*)qkv, R, C, H, pad, (float*)ctx, st))) return rc;
    }
  }
  {
    LinearEpilogue e{RNAMSM_EPI_BIAS_RESIDUAL, w->col.b_out, 1.f, 0, nullptr};
    const LnFuse ln{w->ffn_ln_w, w->ffn_ln_b, eps, xn, dtype == RNAMSM_F16, 0, 0, cnt};
    if ((rc = range_watch(xn, T * D, col_dt, st)) || (rc = range_watch(ctx, T * D, col_dt, st))) return rc;
    if (R > 1 && (rc = range_watch(qkv, T * 3 * D, col_dt, st))) return rc;
    if ((rc = linear_any(ctx, w->col.w_out, T, D, D, col_dt, e, x, st, fuse ? &ln : nullptr, split))) return rc;
  }

  // ---- feed-forward: x += fc2(gelu(fc1(LN(x))))                        modules.py:423-427
  if (!fuse && (rc = launch_layernorm(x, w->ffn_ln_w, w->ffn_ln_b, xn, dtype, T, D, eps, st))) return rc;
  {
    LinearEpilogue e{RNAMSM_EPI_BIAS_GELU, w->fc1_b, 1.f, 0, nullptr};
    if ((rc = linear_any(xn, w->fc1_w, T, F, D, dtype, e, qkv, st, nullptr, split))) return rc;
    if ((rc = range_watch(xn, T * D, dtype, st)) || (rc = range_watch(qkv, T * (long long)F, dtype, st))) return rc;
  }
  {
    LinearEpilogue e{RNAMSM_EPI_BIAS_RESIDUAL, w->fc2_b, 1.f, 0, nullptr};
    LnFuse ln{nullptr, nullptr, eps, xn, 0, 0, 0, cnt};
    if (next) { ln.w = next->w; ln.b = next->b; ln.out_fp16 = next->dtype == RNAMSM_F16; }
    if ((rc = linear_any(qkv, w->fc2_w, T, D, F, dtype, e, x, st, next ? &ln : nullptr, split))) return rc;
  }
  return 0;
}


// ---------------------------------------------------------------------------------------------
// Several short MSAs in one pass.  The tokens of all MSAs sit back to back in one
// [T_total, D] stream: the token-local work (three LayerNorms, six GEMMs per layer) is launched ONCE over
// T_total, so a 256 x 51 alignment no longer pays a partial last wave and a pipeline ramp per GEMM; the
// attention steps, which are per alignment by definition (tied logits sum over that MSA's rows with its
// own 1/sqrt(R), modules.py:713-715; column attention runs over that MSA's depth), are launched per MSA on
// its slice.  Every output element is produced by exactly the instruction sequence of the one-MSA call, so
// the results are bit-identical to n_msa calls of rnamsm_msa_forward -- NOT the reference's padded [B,R,C]
// batch, whose align_scaling would use the padded row count.
// ---------------------------------------------------------------------------------------------
struct BatchPlan {
  size_t el;
  long long T;
  size_t off_xn, off_qkvh, off_partial, off_probs, off_map, off_pad, off_cnt, total;
  size_t partial_bytes, probs_bytes;
  std::vector<Plan> msa;      // every MSA's own plan, computed ONCE here and used by every layer
};

static int make_batch_plan(int n, const int* R, const int* C, int D, int H, int F, int dtype, BatchPlan* bp) {
  BatchPlan b{};
  b.el = 2;
  size_t partial = 0, probs = 0, map = 0;
  for (int i = 0; i < n; ++i) {
    if (R[i] < 2 || C[i] < 1) {
      set_error("msa_forward_batch: MSA %d has R=%d C=%d (need R >= 2, C >= 1; run single-row inputs through rnamsm_msa_forward)",
                i, R[i], C[i]);
      return 2;
    }
    const Plan p = make_plan(R[i], C[i], D, H, F, dtype);
    b.msa.push_back(p);
    partial = std::max(partial, align256((size_t)p.splits * H * C[i] * C[i] * 4));
    probs = std::max(probs, align256((size_t)H * C[i] * p.ldp * p.el));
    map = std::max(map, align256((size_t)H * C[i] * C[i] * 4));
    b.T += (long long)R[i] * C[i];
  }
  size_t o = 0;
  b.off_xn = o;      o += align256((size_t)b.T * D * b.el);
  b.off_qkvh = o;    o += align256((size_t)b.T * (size_t)std::max(4 * D, F) * b.el);
  b.off_partial = o; o += partial;
  b.off_probs = o;   o += probs;
  b.partial_bytes = partial; b.probs_bytes = probs;
  b.off_map = o;     o += map;
  b.off_pad = o;     o += align256((size_t)b.T);
  b.off_cnt = o;     o += align256(ln_counter_bytes(b.T));
  b.total = o;
  *bp = b;
  return 0;
}

static int layer_forward_batch(const rnamsm_layer_weights* w, int D, int H, int F, float eps, float* x, int n,
                               const int* R, const int* C, const uint8_t* has_pad, bool any_pad, int dtype, int layer,
                               float* const* row_attn_out, uint8_t* ws, const BatchPlan& bp,
                               cudaStream_t st, bool xn_ready = false, const NextLn* next = nullptr) {
  const long long T = bp.T;
  const size_t el = bp.el;
  uint8_t* xn = ws + bp.off_xn;
  uint8_t* qkv = ws + bp.off_qkvh;
  uint8_t* ctx = qkv + (size_t)T * 3 * D * el;
  float* partial = reinterpret_cast<float*>(ws + bp.off_partial);
  void* probs_lp = ws + bp.off_probs;
  const uint8_t* pad = ws + bp.off_pad;
  const int row_dt = block_dtype(w->row.dtype, dtype);
  const int col_dt = block_dtype(w->col.dtype, dtype);
  int rc;

  // token-local LayerNorms ride on the residual GEMMs as in layer_forward; the column block's LayerNorm stays a pass
  // per MSA because it transposes each MSA's own token order
  const bool fuse = fuse_ln_enabled();
  int* cnt = reinterpret_cast<int*>(ws + bp.off_cnt);
  if (fuse && !xn_ready) RNAMSM_CHECK_CUDA(cudaMemsetAsync(cnt, 0, ln_counter_bytes(T), st));
  // ---- tied row attention (same constants as layer_forward's 16-bit branch)
  if (!(xn_ready && fuse) && (rc = launch_layernorm(x, w->row.ln_w, w->row.ln_b, xn, row_dt, T, D, eps, st))) return rc;
  {
    LinearEpilogue e{RNAMSM_EPI_BIAS, w->row.b_qkv, 0.125f, D, any_pad ? pad : nullptr};
    if ((rc = linear_any(xn, w->row.w_qkv, T, 3 * D, D, row_dt, e, qkv, st))) return rc;
  }
  long long off = 0;
  for (int i = 0; i < n; ++i) {
    const Plan& p = bp.msa[i];
    RNAMSM_REQUIRE((size_t)p.splits * H * C[i] * C[i] * 4 <= bp.partial_bytes && (size_t)H * C[i] * p.ldp * el <= bp.probs_bytes,
                   "msa_forward_batch: MSA %d (splits %d) does not fit the planned partial / probs slabs", i, p.splits);
    const uint8_t* qkv_i = qkv + (size_t)off * 3 * D * el;
    const uint8_t* pad_i = (has_pad && has_pad[i]) ? pad + off : nullptr;
    float* map = (row_attn_out && row_attn_out[i])
                     ? row_attn_out[i] + (size_t)layer * H * C[i] * C[i]
                     : reinterpret_cast<float*>(ws + bp.off_map);
    if (p.short_chunks > 0) {   // the same one-launch kernel rnamsm_msa_forward uses for this shape (bit-identical results)
      if ((rc = launch_row_attn_short_16(qkv_i, R[i], C[i], H, row_dt == RNAMSM_F16, partial, p.short_chunks, pad_i,
                                         1.0f / sqrtf((float)R[i]), map, probs_lp, p.ldp, ctx + (size_t)off * D * el, st)))
        return rc;
    } else {
      if ((rc = launch_row_logits_16(qkv_i, R[i], C[i], H, row_dt == RNAMSM_F16, partial, p.splits, st))) return rc;
      if ((rc = launch_row_softmax(partial, p.splits, H, C[i], pad_i, 1.0f / sqrtf((float)R[i]), map, probs_lp, p.ldp,
                                   row_dt, st)))
        return rc;
      if ((rc = launch_row_av_16(probs_lp, p.ldp, qkv_i, R[i], C[i], H, row_dt == RNAMSM_F16, ctx + (size_t)off * D * el,
                                 st)))
        return rc;
    }
    off += (long long)R[i] * C[i];
  }
  {
    LinearEpilogue e{RNAMSM_EPI_BIAS_RESIDUAL, w->row.b_out, 1.f, 0, nullptr};
    if ((rc = linear_any(ctx, w->row.w_out, T, D, D, row_dt, e, x, st))) return rc;
  }

  // ---- column attention: LayerNorm per MSA (it transposes that MSA's token order), one QKV GEMM, flash
  // kernel per MSA on its [C, R, 3D] slice, one out-projection.
  off = 0;
  for (int i = 0; i < n; ++i) {
    const long long Ti = (long long)R[i] * C[i];
    if ((rc = launch_layernorm(x + (size_t)off * D, w->col.ln_w, w->col.ln_b, xn + (size_t)off * D * el, col_dt, Ti, D,
                               eps, st, R[i], C[i])))
      return rc;
    off += Ti;
  }
  {
    LinearEpilogue e{RNAMSM_EPI_BIAS, w->col.b_qkv, 1.0f / sqrtf(64.f), D, nullptr};
    if ((rc = linear_any(xn, w->col.w_qkv, T, 3 * D, D, col_dt, e, qkv, st))) return rc;
  }
  off = 0;
  for (int i = 0; i < n; ++i) {
    const uint8_t* pad_i = (has_pad && has_pad[i]) ? pad + off : nullptr;
    if ((rc = launch_col_attn_16(qkv + (size_t)off * 3 * D * el, R[i], C[i], H, col_dt == RNAMSM_F16, 1, pad_i,
                                 ctx + (size_t)off * D * el, st)))
      return rc;
    off += (long long)R[i] * C[i];
  }
  {
    LinearEpilogue e{RNAMSM_EPI_BIAS_RESIDUAL, w->col.b_out, 1.f, 0, nullptr};
    const LnFuse ln{w->ffn_ln_w, w->ffn_ln_b, eps, xn, dtype == RNAMSM_F16, 0, 0, cnt};
    if ((rc = linear_any(ctx, w->col.w_out, T, D, D, col_dt, e, x, st, fuse ? &ln : nullptr))) return rc;
  }

  // ---- feed-forward over all tokens
  if (!fuse && (rc = launch_layernorm(x, w->ffn_ln_w, w->ffn_ln_b, xn, dtype, T, D, eps, st))) return rc;
  {
    LinearEpilogue e{RNAMSM_EPI_BIAS_GELU, w->fc1_b, 1.f, 0, nullptr};
    if ((rc = linear_any(xn, w->fc1_w, T, F, D, dtype, e, qkv, st))) return rc;
  }
  {
    LinearEpilogue e{RNAMSM_EPI_BIAS_RESIDUAL, w->fc2_b, 1.f, 0, nullptr};
    LnFuse ln{nullptr, nullptr, eps, xn, 0, 0, 0, cnt};
    const bool emit = fuse && next != nullptr;
    if (emit) { ln.w = next->w; ln.b = next->b; ln.out_fp16 = next->dtype == RNAMSM_F16; }
    if ((rc = linear_any(qkv, w->fc2_w, T, D, F, dtype, e, x, st, emit ? &ln : nullptr))) return rc;
  }
  return 0;
}

}  // namespace rnamsm

using namespace rnamsm;

extern "C" {

int rnamsm_version(void) { return RNAMSM_ABI_VERSION; }
const char* rnamsm_last_error(void) { return get_error(); }
long long rnamsm_launch_count(void) { return launch_count(); }
int rnamsm_gemm_pairs(void) { return gemm_max_pairs(); }

int rnamsm_profile_enable(int on) {
  for (auto& ev : g_prof_events) { g_prof_pool.push_back(ev.a); g_prof_pool.push_back(ev.b); }
  g_prof_events.clear();
  g_prof_on = on != 0;
  return 0;
}
int rnamsm_profile_num_classes(void) { return KC_COUNT; }
const char* rnamsm_profile_class_name(int i) { return (i >= 0 && i < KC_COUNT) ? kClassNames[i] : ""; }
int rnamsm_profile_collect(double* ms_out, long long* launches_out, int n) {
  RNAMSM_REQUIRE(n >= KC_COUNT, "profile_collect: need %d slots", (int)KC_COUNT);
  for (int i = 0; i < n; ++i) { ms_out[i] = 0.0; launches_out[i] = 0; }
  for (auto& ev : g_prof_events) {
    RNAMSM_CHECK_CUDA(cudaEventSynchronize(ev.b));
    float ms = 0.f;
    RNAMSM_CHECK_CUDA(cudaEventElapsedTime(&ms, ev.a, ev.b));
    ms_out[ev.cls] += ms;
    launches_out[ev.cls] += 1;
    g_prof_pool.push_back(ev.a);
    g_prof_pool.push_back(ev.b);
  }
  g_prof_events.clear();
  return 0;
}

int rnamsm_device_check(void) {
  int dev = 0;
  RNAMSM_CHECK_CUDA(cudaGetDevice(&dev));
  int major = 0, minor = 0;
  RNAMSM_CHECK_CUDA(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
  RNAMSM_CHECK_CUDA(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
  RNAMSM_REQUIRE(major == 10, "rnamsm_b200 is built for sm_100a (B200); device %d is sm_%d%d", dev, major, minor);
  (void)gemm_max_pairs();   // resolve the CTA-pair occupancy now: every later plan and launch sees the same value
  return 0;
}

int rnamsm_embed_layernorm(const int64_t* tokens, int R, int C, const float* tok_emb, int vocab, const float* pos_emb,
                           int n_pos, const float* row_pos, const float* ln_w, const float* ln_b, int D, int pad_idx,
                           float eps, float* x_out, uint8_t* pad_out, void* stream) {
  return launch_embed_ln(tokens, R, C, tok_emb, vocab, pos_emb, n_pos, row_pos, ln_w, ln_b, D, pad_idx, eps, x_out,
                         pad_out, (cudaStream_t)stream);
}

int rnamsm_layernorm(const float* x, const float* w, const float* b, void* y, int y_dtype, long long n_rows, int D,
                     float eps, int tr_R, int tr_C, void* stream) {
  return launch_layernorm(x, w, b, y, y_dtype, n_rows, D, eps, (cudaStream_t)stream, tr_R, tr_C);
}

int rnamsm_linear(const void* x, const void* W, const float* bias, long long M, int N, int K, int dtype, int epilogue,
                  float q_scale, int q_cols, const uint8_t* row_mask, void* out, void* stream) {
  RNAMSM_REQUIRE(epilogue >= 0 && epilogue <= 2, "rnamsm_linear: unknown epilogue %d", epilogue);
  LinearEpilogue e{epilogue, bias, q_scale, q_cols, row_mask};
  return linear_any(x, W, M, N, K, dtype, e, out, (cudaStream_t)stream);
}

size_t rnamsm_linear_tf32_scratch_bytes(long long M, int N, int K) { return tf32_scratch_bytes(M, N, K); }

int rnamsm_linear_tf32(const float* x, const float* W, const float* bias, long long M, int N, int K, int epilogue,
                       float q_scale, int q_cols, const uint8_t* row_mask, float* out, void* scratch, size_t scratch_bytes,
                       void* stream) {
  RNAMSM_REQUIRE(epilogue >= 0 && epilogue <= 2, "linear_tf32: unknown epilogue %d", epilogue);
  RNAMSM_REQUIRE(scratch != nullptr && scratch_bytes >= tf32_scratch_bytes(M, N, K) &&
                     (reinterpret_cast<uintptr_t>(scratch) & 255) == 0,
                 "linear_tf32: scratch of %zu bytes (256 B aligned) required", tf32_scratch_bytes(M, N, K));
  LinearEpilogue e{epilogue, bias, q_scale, q_cols, row_mask};
  return linear_tf32(x, W, M, N, K, e, out, (uint8_t*)scratch, (cudaStream_t)stream);
}

int rnamsm_linear_residual_layernorm(const void* x, const void* W, const float* bias, long long M, int N, int K, int dtype,
                                     float* resid, const float* ln_w, const float* ln_b, float eps, void* y, int y_dtype,
                                     int tr_R, int tr_C, int* counters, void* stream) {
  RNAMSM_REQUIRE(is16(dtype) && is16(y_dtype), "linear_residual_layernorm: 16-bit operands and output only (dtype %d, y %d)",
                 dtype, y_dtype);
  LinearEpilogue e{RNAMSM_EPI_BIAS_RESIDUAL, bias, 1.f, 0, nullptr};
  const LnFuse ln{ln_w, ln_b, eps, y, y_dtype == RNAMSM_F16, tr_R, tr_C, counters};
  return launch_linear_16(x, W, M, N, K, dtype == RNAMSM_F16, e, resid, (cudaStream_t)stream, &ln);
}

int rnamsm_row_attn_splits(int R, int C, int H, int dtype) { return pick_splits(R, C, H, dtype); }

int rnamsm_row_attn_logits(const void* qkv, int R, int C, int H, int dtype, float* partial, int n_splits, void* stream) {
  if (is16(dtype)) return launch_row_logits_16(qkv, R, C, H, dtype == RNAMSM_F16, partial, n_splits, (cudaStream_t)stream);
  return launch_row_logits_f32((const float*)qkv, R, C, H, partial, n_splits, (cudaStream_t)stream);
}

int rnamsm_row_attn_short_chunks(int R, int C, int H) { return row_attn_short_chunks(R, C, H); }

int rnamsm_row_attn_short(const void* qkv, int R, int C, int H, int dtype, const uint8_t* key_pad, float logit_scale,
                          float* partial, int n_chunks, float* probs_out, void* probs_lp, int ld_lp, void* ctx,
                          void* stream) {
  RNAMSM_REQUIRE(is16(dtype), "rnamsm_row_attn_short: 16-bit dtypes only (the fp32 path keeps the three-kernel chain)");
  return launch_row_attn_short_16(qkv, R, C, H, dtype == RNAMSM_F16, partial, n_chunks, key_pad, logit_scale, probs_out,
                                  probs_lp, ld_lp, ctx, (cudaStream_t)stream);
}

int rnamsm_row_softmax(const float* partial, int n_splits, int H, int C, const uint8_t* key_pad, float logit_scale,
                       float* probs_out, void* probs_lp, int ld_lp, int dtype, void* stream) {
  return launch_row_softmax(partial, n_splits, H, C, key_pad, logit_scale, probs_out, probs_lp, ld_lp, dtype,
                            (cudaStream_t)stream);
}

int rnamsm_row_attn_av(const void* probs, int ldp, const void* qkv, int R, int C, int H, int dtype, void* ctx,
                       void* stream) {
  if (is16(dtype)) return launch_row_av_16(probs, ldp, qkv, R, C, H, dtype == RNAMSM_F16, ctx, (cudaStream_t)stream);
  return launch_row_av_f32((const float*)probs, ldp, (const float*)qkv, R, C, H, (float*)ctx, (cudaStream_t)stream);
}

int rnamsm_col_attn(const void* qkv, int R, int C, int H, int dtype, int qkv_col_major, const uint8_t* pad, void* ctx,
                    void* stream) {
  RNAMSM_REQUIRE(R >= 2, "rnamsm_col_attn: R=%d (the R == 1 shortcut is out_proj(v_proj(x)))", R);
  RNAMSM_REQUIRE(!qkv_col_major || is16(dtype), "rnamsm_col_attn: the column-major q|k|v layout exists in the 16-bit path only");
  if (is16(dtype))
    return launch_col_attn_16(qkv, R, C, H, dtype == RNAMSM_F16, qkv_col_major, pad, ctx, (cudaStream_t)stream);
  return launch_col_attn_f32((const float*)qkv, R, C, H, pad, (float*)ctx, (cudaStream_t)stream);
}

int rnamsm_vocab_proj(const float* h, const float* E, const float* bias, long long M, int V, int D, float* out,
                      void* stream) {
  return launch_vocab_proj(h, E, bias, M, V, D, out, (cudaStream_t)stream);
}

size_t rnamsm_workspace_bytes(int R, int C, int D, int H, int F, int dtype) {
  if (R <= 0 || C <= 0) return 0;
  return make_plan(R, C, D, H, F, dtype & 0xff, (dtype & RNAMSM_F32_TENSOR) != 0).total;
}

int rnamsm_range_scan(const void* buf, long long n, int dtype, unsigned long long* counters, void* stream) {
  RNAMSM_REQUIRE(counters != nullptr, "range_scan: counters required");
  return launch_range_scan(buf, n, dtype, counters, (cudaStream_t)stream);
}
int rnamsm_debug_range_watch(unsigned long long* counters) {
  g_range_counters = counters;
  return 0;
}

int rnamsm_fused_layernorm(int dtype) { return (is16(dtype) && fuse_ln_enabled()) ? 1 : 0; }

int rnamsm_layer_forward(const rnamsm_layer_weights* w, int D, int H, int F, float ln_eps, float* x, int R, int C,
                         const uint8_t* pad, int dtype, float* row_probs_out, void* workspace, size_t workspace_bytes,
                         int xn_ready, const float* next_ln_w, const float* next_ln_b, int next_ln_dtype, void* stream) {
  RNAMSM_REQUIRE(D == H * 64, "layer_forward: head_dim must be 64 (D=%d H=%d)", D, H);
  const bool f32_tensor = (dtype & RNAMSM_F32_TENSOR) != 0;
  dtype &= 0xff;
  RNAMSM_REQUIRE(dtype == RNAMSM_F32 || (is16(dtype) && !f32_tensor), "layer_forward: unknown dtype %d", dtype);
  const Plan p = make_plan(R, C, D, H, F, dtype, f32_tensor);
  RNAMSM_REQUIRE(workspace_bytes >= p.total, "layer_forward: workspace %zu < required %zu", workspace_bytes, p.total);
  RNAMSM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "layer_forward: workspace must be 256 B aligned");
  RNAMSM_REQUIRE((next_ln_w == nullptr) == (next_ln_b == nullptr), "layer_forward: next_ln_w and next_ln_b go together");
  const NextLn next{next_ln_w, next_ln_b, block_dtype(next_ln_dtype, dtype)};
  return layer_forward(w, D, H, F, ln_eps, x, R, C, pad, dtype, row_probs_out, (uint8_t*)workspace, p,
                       (cudaStream_t)stream, xn_ready != 0, next_ln_w ? &next : nullptr, f32_tensor);
}

int rnamsm_msa_forward(const rnamsm_model_weights* m, const int64_t* tokens, int R, int C, int has_pad, int dtype,
                       float* x, float* row_attn_out, float* const* rep_out, float* logits_out, void* workspace,
                       size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int D = m->embed_dim, H = m->num_heads, F = m->ffn_dim, N = m->num_layers;
  RNAMSM_REQUIRE(D == H * 64, "msa_forward: head_dim must be 64 (D=%d H=%d)", D, H);
  const bool f32_tensor = (dtype & RNAMSM_F32_TENSOR) != 0;
  dtype &= 0xff;
  RNAMSM_REQUIRE(dtype == RNAMSM_F32 || (is16(dtype) && !f32_tensor), "msa_forward: unknown dtype %d", dtype);
  RNAMSM_REQUIRE(R >= 1 && C >= 1, "msa_forward: empty MSA (R=%d C=%d)", R, C);
  const Plan p = make_plan(R, C, D, H, F, dtype, f32_tensor);
  const size_t T = (size_t)R * C;
  const size_t pad_bytes = (T + 255) & ~(size_t)255;
  RNAMSM_REQUIRE(workspace_bytes >= p.total + pad_bytes, "msa_forward: workspace %zu < required %zu", workspace_bytes,
                 p.total + pad_bytes);
  RNAMSM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "msa_forward: workspace must be 256 B aligned");
  uint8_t* ws = (uint8_t*)workspace;
  uint8_t* pad = ws + p.total;
  int rc;
  if ((rc = launch_embed_ln(tokens, R, C, m->tok_emb, m->vocab, m->pos_emb, m->n_pos, m->row_pos, m->ln_before_w,
                            m->ln_before_b, D, m->pad_idx, m->ln_eps, x, pad, st)))
    return rc;
  const uint8_t* pad_arg = has_pad ? pad : nullptr;
  for (int l = 0; l < N; ++l) {
    if (rep_out && rep_out[l]) RNAMSM_CHECK_CUDA(cudaMemcpyAsync(rep_out[l], x, T * D * 4, cudaMemcpyDeviceToDevice, st));
    float* map = row_attn_out ? row_attn_out + (size_t)l * H * C * C : nullptr;
    // layer l's fc2 epilogue also writes layer l+1's row-block LayerNorm (16-bit path)
    const bool fuse = is16(dtype) && fuse_ln_enabled();
    NextLn next{nullptr, nullptr, dtype};
    if (fuse && l + 1 < N)
      next = NextLn{m->layers[l + 1].row.ln_w, m->layers[l + 1].row.ln_b, block_dtype(m->layers[l + 1].row.dtype, dtype)};
    if ((rc = layer_forward(&m->layers[l], D, H, F, m->ln_eps, x, R, C, pad_arg, dtype, map, ws, p, st, fuse && l > 0,
                            next.w ? &next : nullptr, f32_tensor)))
      return rc;
  }
  // 16-bit path with logits wanted: the LM head's dense GEMM runs on the tensor cores, so it needs the final
  // LayerNorm's output in 16 bits as well -- written by a second (16-bit output) LayerNorm launch on the same rows
  // before the in-place fp32 one
  const bool lm16 = logits_out && is16(dtype) && m->lm_dense_w16 != nullptr;
  if (lm16 && (rc = launch_layernorm(x, m->ln_after_w, m->ln_after_b, ws + p.off_xn, dtype, (long long)T, D, m->ln_eps, st)))
    return rc;
  // emb_layer_norm_after in place (model.py:396): fp32 -> fp32, each warp reads its row before writing it
  if ((rc = launch_layernorm(x, m->ln_after_w, m->ln_after_b, x, RNAMSM_F32, (long long)T, D, m->ln_eps, st))) return rc;
  if (logits_out) {
    // RobertaLMHead (modules.py:313-319): dense -> erf-GELU -> LayerNorm -> tied projection + bias.
    // 16-bit path: dense + GELU on tcgen05 (16-bit h), LayerNorm with fp32 statistics -> fp32, 12-wide tied projection in
    // fp32.  fp32 path: FFMA throughout.  Scratch: the q|k|v|ctx region.
    uint8_t* hbuf = ws + p.off_qkvh;
    float* h32 = reinterpret_cast<float*>(hbuf + align256(T * D * 4));
    LinearEpilogue e{RNAMSM_EPI_BIAS_GELU, m->lm_dense_b, 1.f, 0, nullptr};
    if (lm16) {
      if ((rc = launch_linear_16(ws + p.off_xn, m->lm_dense_w16, (long long)T, D, D, dtype == RNAMSM_F16, e, hbuf, st))) return rc;
      if ((rc = launch_layernorm((const float*)hbuf, m->lm_ln_w, m->lm_ln_b, h32, RNAMSM_F32, (long long)T, D, m->ln_eps, st, 0,
                                 0, dtype)))
        return rc;
    } else {
      if ((rc = linear_any(x, m->lm_dense_w, (long long)T, D, D, RNAMSM_F32, e, hbuf, st, nullptr,
                           (dtype == RNAMSM_F32 && f32_tensor) ? ws + p.off_split : nullptr)))
        return rc;
      if ((rc = launch_layernorm((const float*)hbuf, m->lm_ln_w, m->lm_ln_b, h32, RNAMSM_F32, (long long)T, D, m->ln_eps, st)))
        return rc;
    }
    if ((rc = launch_vocab_proj(h32, m->tok_emb, m->lm_bias, (long long)T, m->vocab, D, logits_out, st))) return rc;
  }
  return 0;
}

size_t rnamsm_batch_workspace_bytes(int n_msa, const int* R, const int* C, int D, int H, int F, int dtype) {
  BatchPlan bp;
  if (n_msa <= 0 || !is16(dtype) || make_batch_plan(n_msa, R, C, D, H, F, dtype, &bp)) return 0;
  return bp.total;
}

int rnamsm_msa_forward_batch(const rnamsm_model_weights* m, int n_msa, const int64_t* tokens, const int* R, const int* C,
                             const uint8_t* has_pad, int dtype, float* x, float* const* row_attn_out, void* workspace,
                             size_t workspace_bytes, void* stream) {
  cudaStream_t st = (cudaStream_t)stream;
  const int D = m->embed_dim, H = m->num_heads, F = m->ffn_dim, N = m->num_layers;
  RNAMSM_REQUIRE(D == H * 64, "msa_forward_batch: head_dim must be 64 (D=%d H=%d)", D, H);
  RNAMSM_REQUIRE(is16(dtype), "msa_forward_batch: 16-bit operand types only (dtype %d); the fp32 parity path runs one MSA per call",
                 dtype);
  RNAMSM_REQUIRE(n_msa >= 1 && R && C, "msa_forward_batch: empty batch");
  BatchPlan bp;
  int rc;
  if ((rc = make_batch_plan(n_msa, R, C, D, H, F, dtype, &bp))) return rc;
  RNAMSM_REQUIRE(workspace_bytes >= bp.total, "msa_forward_batch: workspace %zu < required %zu", workspace_bytes, bp.total);
  RNAMSM_REQUIRE((reinterpret_cast<uintptr_t>(workspace) & 255) == 0, "msa_forward_batch: workspace must be 256 B aligned");
  uint8_t* ws = (uint8_t*)workspace;
  uint8_t* pad = ws + bp.off_pad;
  bool any_pad = false;
  long long off = 0;
  for (int i = 0; i < n_msa; ++i) {
    any_pad = any_pad || (has_pad && has_pad[i]);
    if ((rc = launch_embed_ln(tokens + off, R[i], C[i], m->tok_emb, m->vocab, m->pos_emb, m->n_pos, m->row_pos,
                              m->ln_before_w, m->ln_before_b, D, m->pad_idx, m->ln_eps, x + (size_t)off * D, pad + off, st)))
      return rc;
    off += (long long)R[i] * C[i];
  }
  for (int l = 0; l < N; ++l) {
    NextLn next{nullptr, nullptr, dtype};
    if (l + 1 < N)
      next = NextLn{m->layers[l + 1].row.ln_w, m->layers[l + 1].row.ln_b, block_dtype(m->layers[l + 1].row.dtype, dtype)};
    if ((rc = layer_forward_batch(&m->layers[l], D, H, F, m->ln_eps, x, n_msa, R, C, has_pad, any_pad, dtype, l,
                                  row_attn_out, ws, bp, st, l > 0, next.w ? &next : nullptr)))
      return rc;
  }
  return launch_layernorm(x, m->ln_after_w, m->ln_after_b, x, RNAMSM_F32, bp.T, D, m->ln_eps, st);
}

}  // extern "C"
