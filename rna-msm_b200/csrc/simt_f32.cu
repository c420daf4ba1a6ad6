// fp32 parity path (dtype = RNAMSM_F32): every contraction of the forward in plain FFMA with fp32
// accumulation, so the result tracks the reference's fp32 eager forward to ~1e-6 norm-relative
// (gate: 1e-4).  One shared 128x128x16 register-tiled SGEMM core; the four contractions differ
// only in how operand elements are addressed and how results are stored:
//   linear      : out[m,n]      = sum_k x[m,k] W[n,k]                       (modules.py:760-766 etc.)
//   tied logits : L[h,i,j]      = sum_{r,d} q[r,i,h,d] k[r,j,h,d]           (modules.py:774)
//   tied AV     : ctx[r,i,h,d]  = sum_j P[h,i,j] v[r,j,h,d]                 (modules.py:797)
//   column attn : flash-style online softmax over rows per (column, head)   (modules.py:896-923)
// The bf16 production path lives in umma_gemm.cu / col_attn_umma.cu.
#include "../../include/rnamsm_b200.h"
#include "common.cuh"
#include "launch.h"

namespace rnamsm {

constexpr int BM = 128, BN = 128, BK = 16, NT = 256;

template <class P>
__global__ void __launch_bounds__(NT) sgemm_kernel(const P p) {
  __shared__ __align__(16) float As[BK][BM + 4];
  __shared__ __align__(16) float Bs[BK][BN + 4];
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int n0 = blockIdx.x * BN, m0 = blockIdx.y * BM, bz = blockIdx.z;
  const int M = p.M, N = p.N;
  int k_begin, k_end;
  p.k_range(bz, k_begin, k_end);
  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  for (int k0 = k_begin; k0 < k_end; k0 += BK) {
#pragma unroll
    for (int i = 0; i < (BM * BK) / NT; ++i) {
      const int idx = t + i * NT;
      int mm, kk;
      if (P::kAContigK) { kk = idx & (BK - 1); mm = idx >> 4; } else { mm = idx & (BM - 1); kk = idx >> 7; }
      const int m = m0 + mm, k = k0 + kk;
      As[kk][mm] = (m < M && k < k_end) ? p.load_a(bz, m, k) : 0.f;
    }
#pragma unroll
    for (int i = 0; i < (BN * BK) / NT; ++i) {
      const int idx = t + i * NT;
      int nn, kk;
      if (P::kBContigK) { kk = idx & (BK - 1); nn = idx >> 4; } else { nn = idx & (BN - 1); kk = idx >> 7; }
      const int n = n0 + nn, k = k0 + kk;
      Bs[kk][nn] = (n < N && k < k_end) ? p.load_b(bz, n, k) : 0.f;
    }
    __syncthreads();
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4*>(&As[kk][ty * 4]);
      const float4 a1 = *reinterpret_cast<const float4*>(&As[kk][64 + ty * 4]);
      const float4 b0 = *reinterpret_cast<const float4*>(&Bs[kk][tx * 4]);
      const float4 b1 = *reinterpret_cast<const float4*>(&Bs[kk][64 + tx * 4]);
      const float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    __syncthreads();
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const int m = m0 + (i < 4 ? ty * 4 + i : 64 + ty * 4 + (i - 4));
    if (m >= M) continue;
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int n = n0 + (j < 4 ? tx * 4 + j : 64 + tx * 4 + (j - 4));
      if (n < N) p.store(bz, m, n, acc[i][j]);
    }
  }
}

// ---- linear ----------------------------------------------------------------------------------
struct LinearProb {
  static constexpr bool kAContigK = true, kBContigK = true;
  int M, N, K;
  const float* x;
  const float* W;
  float* out;
  LinearEpilogue e;
  __device__ void k_range(int, int& b, int& en) const { b = 0; en = K; }
  __device__ float load_a(int, int m, int k) const { return x[(size_t)m * K + k]; }
  __device__ float load_b(int, int n, int k) const { return W[(size_t)n * K + k]; }
  __device__ void store(int, int m, int n, float v) const {
    v += e.bias ? e.bias[n] : 0.f;
    const size_t o = (size_t)m * N + n;
    if (e.kind == RNAMSM_EPI_BIAS) {
      if (n < e.q_cols) {
        v *= e.q_scale;
        if (e.row_mask && e.row_mask[m]) v = 0.f;
      }
      out[o] = v;
    } else if (e.kind == RNAMSM_EPI_BIAS_GELU) {
      out[o] = gelu_erf(v);
    } else {
      out[o] += v;
    }
  }
};

int launch_linear_f32(const float* x, const float* W, long long M, int N, int K, const LinearEpilogue& epi, float* out,
                      cudaStream_t st) {
  RNAMSM_REQUIRE(M > 0 && M < (1LL << 31) && N > 0 && K > 0, "linear_f32: bad shape M=%lld N=%d K=%d", M, N, K);
  LinearProb p{(int)M, N, K, x, W, out, epi};
  dim3 grid(ceil_div(N, BN), ceil_div(M, BM), 1);
  RNAMSM_REQUIRE(grid.y <= 65535, "linear_f32: M=%lld too large for this path", M);
  ProfScope prof(linear_class(epi.kind, N, K), st);
  sgemm_kernel<<<grid, NT, 0, st>>>(p);
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- tied row-attention logits ------------------------------------------------------------------
struct TiedLogitsProb {
  static constexpr bool kAContigK = true, kBContigK = true;
  int M, N;  // = C
  int R, C, H, ld;  // ld = 3*H*64
  int n_splits, rows_per_split;
  const float* qkv;
  float* partial;
  __device__ void k_range(int bz, int& b, int& en) const {
    const int s = bz % n_splits;
    b = s * rows_per_split * 64;
    en = min(R, (s + 1) * rows_per_split) * 64;
  }
  __device__ float load_a(int bz, int i, int k) const {
    const int h = bz / n_splits;
    return qkv[((size_t)(k >> 6) * C + i) * ld + h * 64 + (k & 63)];
  }
  __device__ float load_b(int bz, int j, int k) const {
    const int h = bz / n_splits;
    return qkv[((size_t)(k >> 6) * C + j) * ld + H * 64 + h * 64 + (k & 63)];
  }
  __device__ void store(int bz, int i, int j, float v) const {
    const int h = bz / n_splits, s = bz % n_splits;
    partial[(((size_t)s * H + h) * C + i) * C + j] = v;
  }
};

int launch_row_logits_f32(const float* qkv, int R, int C, int H, float* partial, int n_splits, cudaStream_t st) {
  RNAMSM_REQUIRE(n_splits >= 1 && n_splits <= R, "row_logits_f32: n_splits=%d out of range for R=%d", n_splits, R);
  TiedLogitsProb p{C, C, R, C, H, 3 * H * 64, n_splits, ceil_div(R, n_splits), qkv, partial};
  dim3 grid(ceil_div(C, BN), ceil_div(C, BM), H * n_splits);
  ProfScope prof(KC_ROW_LOGITS, st);
  sgemm_kernel<<<grid, NT, 0, st>>>(p);
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- tied AV --------------------------------------------------------------------------------------
struct TiedAvProb {
  static constexpr bool kAContigK = true, kBContigK = false;
  int M, N;  // M = C (i), N = R*64 (r,d)
  int R, C, H, ld, ldp;
  const float* probs;
  const float* qkv;
  float* ctx;
  __device__ void k_range(int, int& b, int& en) const { b = 0; en = C; }
  __device__ float load_a(int h, int i, int j) const { return probs[((size_t)h * C + i) * ldp + j]; }
  __device__ float load_b(int h, int n, int j) const {
    return qkv[((size_t)(n >> 6) * C + j) * ld + 2 * H * 64 + h * 64 + (n & 63)];
  }
  __device__ void store(int h, int i, int n, float v) const {
    ctx[((size_t)(n >> 6) * C + i) * (H * 64) + h * 64 + (n & 63)] = v;
  }
};

int launch_row_av_f32(const float* probs, int ldp, const float* qkv, int R, int C, int H, float* ctx, cudaStream_t st) {
  RNAMSM_REQUIRE((long long)R * 64 < (1LL << 31), "row_av_f32: R too large");
  TiedAvProb p{C, R * 64, R, C, H, 3 * H * 64, ldp, probs, qkv, ctx};
  dim3 grid(ceil_div(R * 64, BN), ceil_div(C, BM), H);
  RNAMSM_REQUIRE(grid.y <= 65535 && grid.z <= 65535, "row_av_f32: grid too large");
  ProfScope prof(KC_ROW_AV, st);
  sgemm_kernel<<<grid, NT, 0, st>>>(p);
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---- column attention (fp32 flash) ------------------------------------------------------------------
// One block per (column c, head h, 64-query tile).  256 threads as a 16x16 grid; thread (ty,tx)
// owns score rows ty*4.. and score cols / output dims tx*4...  The 16 threads sharing a row live
// in one half-warp, so the row max / sum are 4 xor-shuffles.
constexpr int CQ = 64, CK = 64, HD = 64;

__global__ void __launch_bounds__(256)
col_attn_f32_kernel(const float* __restrict__ qkv, int R, int C, int H, const uint8_t* __restrict__ pad,
                    float* __restrict__ ctx) {
  extern __shared__ __align__(16) float col_smem[];
  float (*Qs)[CQ + 4] = reinterpret_cast<float (*)[CQ + 4]>(col_smem);                      // [d][i]
  float (*Ks)[CK + 4] = reinterpret_cast<float (*)[CK + 4]>(col_smem + HD * (CQ + 4));      // [d][j]
  float (*Vs)[HD + 4] = reinterpret_cast<float (*)[HD + 4]>(col_smem + 2 * HD * (CQ + 4));  // [j][d]
  float (*Ps)[CQ + 4] = reinterpret_cast<float (*)[CQ + 4]>(col_smem + 3 * HD * (CQ + 4));  // [j][i]
  const int t = threadIdx.x, tx = t & 15, ty = t >> 4;
  const int c = blockIdx.x, h = blockIdx.y, i0 = blockIdx.z * CQ;
  const int ld = 3 * H * HD;
  const size_t col_off = (size_t)c * ld + h * HD;
  const size_t row_stride = (size_t)C * ld;

  for (int idx = t; idx < CQ * HD; idx += 256) {
    const int d = idx & 63, i = idx >> 6;
    Qs[d][i] = (i0 + i < R) ? qkv[(size_t)(i0 + i) * row_stride + col_off + d] : 0.f;
  }
  float m_run[4], l_run[4], o[4][4];
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    m_run[a] = -INFINITY;
    l_run[a] = 0.f;
#pragma unroll
    for (int b = 0; b < 4; ++b) o[a][b] = 0.f;
  }
  for (int j0 = 0; j0 < R; j0 += CK) {
    __syncthreads();  // previous tile fully consumed (and Qs visible on the first pass)
    for (int idx = t; idx < CK * HD; idx += 256) {
      const int d = idx & 63, j = idx >> 6;
      const bool ok = j0 + j < R;
      const size_t base = (size_t)(j0 + j) * row_stride + col_off + d;
      Ks[d][j] = ok ? qkv[base + H * HD] : 0.f;
      Vs[j][d] = ok ? qkv[base + 2 * H * HD] : 0.f;
    }
    __syncthreads();
    float s[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) s[a][b] = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; ++d) {
      const float4 qa = *reinterpret_cast<const float4*>(&Qs[d][ty * 4]);
      const float4 kb = *reinterpret_cast<const float4*>(&Ks[d][tx * 4]);
      const float q[4] = {qa.x, qa.y, qa.z, qa.w}, k[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) s[a][b] = fmaf(q[a], k[b], s[a][b]);
    }
    // mask: padded keys -> -10000 (masked_fill, modules.py:911-915); keys past R do not exist.
#pragma unroll
    for (int b = 0; b < 4; ++b) {
      const int j = j0 + tx * 4 + b;
      const bool exists = j < R;
      const bool masked = exists && pad && pad[(size_t)j * C + c];
#pragma unroll
      for (int a = 0; a < 4; ++a) s[a][b] = !exists ? -INFINITY : (masked ? -10000.f : s[a][b]);
    }
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float mx = fmaxf(fmaxf(s[a][0], s[a][1]), fmaxf(s[a][2], s[a][3]));
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      const float m_new = fmaxf(m_run[a], mx);  // finite: every tile has >= 1 existing key
      const float corr = __expf(m_run[a] - m_new);
      float ps = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        const float pv = __expf(s[a][b] - m_new);
        ps += pv;
        Ps[tx * 4 + b][ty * 4 + a] = pv;
      }
#pragma unroll
      for (int off = 8; off > 0; off >>= 1) ps += __shfl_xor_sync(0xffffffffu, ps, off);
      l_run[a] = l_run[a] * corr + ps;
      m_run[a] = m_new;
#pragma unroll
      for (int b = 0; b < 4; ++b) o[a][b] *= corr;
    }
    __syncthreads();
#pragma unroll 8
    for (int j = 0; j < CK; ++j) {
      const float4 pa = *reinterpret_cast<const float4*>(&Ps[j][ty * 4]);
      const float4 vb = *reinterpret_cast<const float4*>(&Vs[j][tx * 4]);
      const float pp[4] = {pa.x, pa.y, pa.z, pa.w}, vv[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) o[a][b] = fmaf(pp[a], vv[b], o[a][b]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    const int i = i0 + ty * 4 + a;
    if (i >= R) continue;
    const float inv = 1.f / l_run[a];
    float4 r4 = make_float4(o[a][0] * inv, o[a][1] * inv, o[a][2] * inv, o[a][3] * inv);
    *reinterpret_cast<float4*>(&ctx[((size_t)i * C + c) * (H * HD) + h * HD + tx * 4]) = r4;
  }
}

int launch_col_attn_f32(const float* qkv, int R, int C, int H, const uint8_t* pad, float* ctx, cudaStream_t st) {
  RNAMSM_REQUIRE(R >= 1 && C >= 1 && H >= 1 && H <= 65535, "col_attn_f32: bad shape");
  dim3 grid(C, H, ceil_div(R, CQ));
  RNAMSM_REQUIRE(grid.z <= 65535, "col_attn_f32: R too large");
  ProfScope prof(KC_COL_ATTN, st);
  constexpr int smem_bytes = 4 * HD * (CQ + 4) * (int)sizeof(float);
  static bool attr_set = false;
  if (!attr_set) {
    RNAMSM_CHECK_CUDA(cudaFuncSetAttribute(col_attn_f32_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem_bytes));
    attr_set = true;
  }
  col_attn_f32_kernel<<<grid, 256, smem_bytes, st>>>(qkv, R, C, H, pad, ctx);
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rnamsm
