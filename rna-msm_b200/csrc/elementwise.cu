// HBM-bound kernels of the RNA-MSM forward: embedding gather + LayerNorm (K1), LayerNorm (K2/K9),
// tied-attention softmax / map export (K5) and the 12-wide tied LM-head projection (K9b).
// All are one-warp-per-row, 128-bit vectorised, fp32 statistics.
#include <cuda_fp16.h>

#include "common.cuh"
#include "launch.h"

namespace rnamsm {

constexpr int kMaxVec = 8;  // D <= 8 * 128 = 1024 features per token

// -------------------------------------------------------------------------------------------
// K1: grid (ceil(C / kColsPerBlock), R); each block first scans its row's non-pad prefix up to
// its column range (positions = cumsum(tok != pad) * (tok != pad) + pad_idx, modules.py:286-291)
// then one warp per token gathers E_tok[tok] + E_pos[pos] + p_row[r], LayerNorms, zeroes pads.
// -------------------------------------------------------------------------------------------
constexpr int kEmbWarps = 8;
constexpr int kColsPerBlock = 64;

__global__ void __launch_bounds__(kEmbWarps * 32)
embed_ln_kernel(const int64_t* __restrict__ tokens, int R, int C, const float* __restrict__ tok_emb, int vocab,
                const float* __restrict__ pos_emb, int n_pos, const float* __restrict__ row_pos,
                const float* __restrict__ ln_w, const float* __restrict__ ln_b, int D, int pad_idx, float eps,
                float* __restrict__ x_out, uint8_t* __restrict__ pad_out) {
  __shared__ int s_pos[kColsPerBlock];
  __shared__ int s_tok[kColsPerBlock];
  pdl_launch_dependents();
  pdl_wait();
  const int r = blockIdx.y;
  const int c0 = blockIdx.x * kColsPerBlock;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t* trow = tokens + (size_t)r * C;
  if (warp == 0) {
    int running = 0;
    const int c_end = min(C, c0 + kColsPerBlock);
    for (int base = 0; base < c_end; base += 32) {
      const int c = base + lane;
      const int tok = (c < c_end) ? (int)trow[c] : pad_idx;
      const bool keep = (c < c_end) && (tok != pad_idx);
      const unsigned m = __ballot_sync(0xffffffffu, keep);
      const int incl = running + __popc(m & (0xffffffffu >> (31 - lane)));
      if (c >= c0 && c < c_end) {
        s_pos[c - c0] = keep ? incl + pad_idx : pad_idx;
        s_tok[c - c0] = tok;
      }
      running += __popc(m);
    }
  }
  __syncthreads();
  const int nv = D / 128;
  const float prow = row_pos ? row_pos[r] : 0.f;
  for (int cc = warp; cc < kColsPerBlock; cc += kEmbWarps) {
    const int c = c0 + cc;
    if (c >= C) break;
    int tok = s_tok[cc];
    int pos = s_pos[cc];
    const bool is_pad = (tok == pad_idx);
    tok = min(max(tok, 0), vocab - 1);
    pos = min(pos, n_pos - 1);  // host validates; clamp keeps the gather in bounds regardless
    float4 v[kMaxVec];
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nv) {
        const int f = lane * 4 + i * 128;
        const float4 a = *reinterpret_cast<const float4*>(tok_emb + (size_t)tok * D + f);
        const float4 p = *reinterpret_cast<const float4*>(pos_emb + (size_t)pos * D + f);
        v[i] = make_float4(a.x + p.x + prow, a.y + p.y + prow, a.z + p.z + prow, a.w + p.w + prow);
      }
    warp_layernorm(v, nv, D, eps, ln_w, ln_b, lane);
    float* dst = x_out + ((size_t)r * C + c) * D;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nv) {
        if (is_pad) v[i] = make_float4(0.f, 0.f, 0.f, 0.f);  // x * (1 - pad), model.py:366-367
        *reinterpret_cast<float4*>(dst + lane * 4 + i * 128) = v[i];
      }
    if (pad_out && lane == 0) pad_out[(size_t)r * C + c] = is_pad ? 1 : 0;
  }
}

// -------------------------------------------------------------------------------------------
// K2: LayerNorm rows of fp32 x -> fp32 or bf16 y.  One warp per row, kLnWarps rows per block.
// -------------------------------------------------------------------------------------------
constexpr int kLnWarps = 8;

// kOut / kIn: 0 = fp32, 1 = bf16, 2 = fp16
template <int kOut, int kIn = 0>
__global__ void __launch_bounds__(kLnWarps * 32)
layernorm_kernel(const void* x, const float* __restrict__ w, const float* __restrict__ b,
                 void* y, long long n_rows, int D, float eps, int tr_R, int tr_C) {  // x may alias y (in-place final LN)
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = D / 128;
  pdl_launch_dependents();
  pdl_wait();
  for (long long row = (long long)blockIdx.x * kLnWarps + warp; row < n_rows;
       row += (long long)gridDim.x * kLnWarps) {
    float4 v[kMaxVec];
    if constexpr (kIn == 0) {
      const float* src = reinterpret_cast<const float*>(x) + (size_t)row * D;
#pragma unroll
      for (int i = 0; i < kMaxVec; ++i)
        if (i < nv) v[i] = *reinterpret_cast<const float4*>(src + lane * 4 + i * 128);
    } else {
      const uint16_t* src = reinterpret_cast<const uint16_t*>(x) + (size_t)row * D;
#pragma unroll
      for (int i = 0; i < kMaxVec; ++i)
        if (i < nv) {
          const uint2 u = *reinterpret_cast<const uint2*>(src + lane * 4 + i * 128);
          if constexpr (kIn == 1) {
            v[i] = make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16),
                               __uint_as_float(u.y & 0xffff0000u));
          } else {
            const __half2 h0 = *reinterpret_cast<const __half2*>(&u.x), h1 = *reinterpret_cast<const __half2*>(&u.y);
            v[i] = make_float4(__low2float(h0), __high2float(h0), __low2float(h1), __high2float(h1));
          }
        }
    }
    warp_layernorm(v, nv, D, eps, w, b, lane);
    // optional token transpose: input row r * C + c -> output row c * R + r (column attention layout)
    const long long orow = tr_C > 0 ? (row % tr_C) * tr_R + row / tr_C : row;
    if constexpr (kOut != 0) {
      uint16_t* dst = reinterpret_cast<uint16_t*>(y) + (size_t)orow * D;
#pragma unroll
      for (int i = 0; i < kMaxVec; ++i)
        if (i < nv) {
          uint2 pk = kOut == 1 ? make_uint2(pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w))
                               : make_uint2(pack_f16(v[i].x, v[i].y), pack_f16(v[i].z, v[i].w));
          *reinterpret_cast<uint2*>(dst + lane * 4 + i * 128) = pk;
        }
    } else {
      float* dst = reinterpret_cast<float*>(y) + (size_t)orow * D;
#pragma unroll
      for (int i = 0; i < kMaxVec; ++i)
        if (i < nv) *reinterpret_cast<float4*>(dst + lane * 4 + i * 128) = v[i];
    }
  }
}

// -------------------------------------------------------------------------------------------
// x += delta (16-bit contribution of a block computed elsewhere, e.g. the column block of the sharded
// forward), x written back in fp32, and LayerNorm(x) written in 16 bits for the next GEMM: the residual
// add rides on the LayerNorm pass that reads x anyway.  kDelta / kOut: 1 = bf16, 2 = fp16.
// -------------------------------------------------------------------------------------------
template <int kDelta, int kOut>
__global__ void __launch_bounds__(kLnWarps * 32)
add_layernorm_kernel(float* __restrict__ x, const uint16_t* __restrict__ delta, const float* __restrict__ w,
                     const float* __restrict__ b, uint16_t* __restrict__ y, long long n_rows, int D, float eps) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = D / 128;
  for (long long row = (long long)blockIdx.x * kLnWarps + warp; row < n_rows; row += (long long)gridDim.x * kLnWarps) {
    float* xr = x + (size_t)row * D;
    const uint16_t* dr = delta + (size_t)row * D;
    float4 v[kMaxVec];
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nv) {
        const int f = lane * 4 + i * 128;
        v[i] = *reinterpret_cast<const float4*>(xr + f);
        const uint2 d2 = *reinterpret_cast<const uint2*>(dr + f);
        float d0, d1, d2f, d3;
        if (kDelta == 1) {
          d0 = __uint_as_float(d2.x << 16); d1 = __uint_as_float(d2.x & 0xffff0000u);
          d2f = __uint_as_float(d2.y << 16); d3 = __uint_as_float(d2.y & 0xffff0000u);
        } else {
          const __half2 h0 = *reinterpret_cast<const __half2*>(&d2.x), h1 = *reinterpret_cast<const __half2*>(&d2.y);
          d0 = __low2float(h0); d1 = __high2float(h0); d2f = __low2float(h1); d3 = __high2float(h1);
        }
        v[i].x += d0; v[i].y += d1; v[i].z += d2f; v[i].w += d3;
        *reinterpret_cast<float4*>(xr + f) = v[i];
      }
    warp_layernorm(v, nv, D, eps, w, b, lane);
    uint16_t* dst = y + (size_t)row * D;
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nv) {
        const uint2 pk = kOut == 1 ? make_uint2(pack_bf16(v[i].x, v[i].y), pack_bf16(v[i].z, v[i].w))
                                   : make_uint2(pack_f16(v[i].x, v[i].y), pack_f16(v[i].z, v[i].w));
        *reinterpret_cast<uint2*>(dst + lane * 4 + i * 128) = pk;
      }
  }
}

// -------------------------------------------------------------------------------------------
// K5: one warp per (head, query column i).  Sums the split-K partial logits, applies the key
// mask, softmax over j, writes the fp32 map (the exported attention map) and the low-precision
// copy that feeds the AV GEMM.  Three passes over the (L2-resident) partial rows keep register
// use independent of C.
// -------------------------------------------------------------------------------------------
constexpr int kSmWarps = 4;

// kLp: element type of the low-precision copy (0 = fp32, 1 = bf16, 2 = fp16).  logit_scale multiplies
// the summed logits before the mask: the 16-bit path keeps q at 64^-1/2 scale (an exact power of
// two) in 16 bits and applies the 1/sqrt(R) of align_scaling (modules.py:713-715) here in fp32.
// kQ > 0: the row's summed logits stay in registers (kQ per lane, C <= 32 kQ) and every split is read ONCE, all loads
// of a row in flight together.  (Round 1 recomputed the split sum in each of the three passes through a loop the
// compiler could not pipeline: one load in flight per lane, 38 us at 12 x 256 x 256 and 470 us at 12 x 1024 x 1024 --
// 0.3 TB/s.)  kQ == 0: any C, the recomputing form.  Same operations in the same order either way.
template <int kLp, int kQ>
__global__ void __launch_bounds__(kSmWarps * 32)
row_softmax_kernel(const float* __restrict__ partial, int n_splits, int H, int C,
                   const uint8_t* __restrict__ key_pad, float logit_scale, float* __restrict__ probs_out,
                   void* __restrict__ probs_lp, int ld_lp) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  pdl_launch_dependents();
  pdl_wait();
  const long long row = (long long)blockIdx.x * kSmWarps + warp;  // h * C + i
  if (row >= (long long)H * C) return;
  const size_t split_stride = (size_t)H * C * C;
  const float* src = partial + (size_t)row * C;
  auto store_lp = [&](int j, float p) {
    if constexpr (kLp == 1)
      reinterpret_cast<__nv_bfloat16*>(probs_lp)[(size_t)row * ld_lp + j] = __float2bfloat16(p);
    else if constexpr (kLp == 2)
      reinterpret_cast<__half*>(probs_lp)[(size_t)row * ld_lp + j] = __float2half_rn(p);
    else
      reinterpret_cast<float*>(probs_lp)[(size_t)row * ld_lp + j] = p;
  };
  float* dst = probs_out + (size_t)row * C;
  if constexpr (kQ > 0) {
    float v[kQ];
#pragma unroll
    for (int q = 0; q < kQ; ++q) v[q] = 0.f;
    for (int s0 = 0; s0 < n_splits; s0 += 4) {          // four splits of every column in flight, added in split order
      float t[4][kQ];
#pragma unroll
      for (int u = 0; u < 4; ++u)
#pragma unroll
        for (int q = 0; q < kQ; ++q) {
          const int j = q * 32 + lane;
          t[u][q] = (s0 + u < n_splits && j < C) ? __ldg(src + (size_t)(s0 + u) * split_stride + j) : 0.f;
        }
#pragma unroll
      for (int u = 0; u < 4; ++u)
        if (s0 + u < n_splits) {
#pragma unroll
          for (int q = 0; q < kQ; ++q) v[q] += t[u][q];
        }
    }
    float mx = -INFINITY;
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const int j = q * 32 + lane;
      float a = v[q] * logit_scale;
      if (j < C && key_pad && key_pad[j]) a = -10000.f;  // masked_fill, modules.py:780-784
      v[q] = a;
      if (j < C) mx = fmaxf(mx, a);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      v[q] = __expf(v[q] - mx);
      if (q * 32 + lane < C) sum += v[q];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
#pragma unroll
    for (int q = 0; q < kQ; ++q) {
      const int j = q * 32 + lane;
      const float p = (j < C) ? v[q] * inv : 0.f;
      if (j < C) dst[j] = p;
      if (probs_lp && j < ld_lp) store_lp(j, p);
    }
    for (int j = kQ * 32 + lane; j < ld_lp; j += 32)     // (ld_lp may exceed 32 kQ by the row padding)
      if (probs_lp) store_lp(j, 0.f);
  } else {
    auto logit = [&](int j) -> float {
      float a = 0.f;
      for (int s = 0; s < n_splits; ++s) a += src[s * split_stride + j];
      a *= logit_scale;
      if (key_pad && key_pad[j]) a = -10000.f;  // masked_fill, modules.py:780-784
      return a;
    };
    float mx = -INFINITY;
    for (int j = lane; j < C; j += 32) mx = fmaxf(mx, logit(j));
    mx = warp_max(mx);
    float sum = 0.f;
    for (int j = lane; j < C; j += 32) sum += __expf(logit(j) - mx);
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    for (int j = lane; j < ld_lp || j < C; j += 32) {
      const float p = (j < C) ? __expf(logit(j) - mx) * inv : 0.f;
      if (j < C) dst[j] = p;
      if (probs_lp && j < ld_lp) store_lp(j, p);
    }
  }
}

// -------------------------------------------------------------------------------------------
// K9b: logits[m, v] = h[m, :] . E[v, :] + bias[v]; one warp per token, V <= 32 outputs.
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256)
vocab_proj_kernel(const float* __restrict__ h, const float* __restrict__ E, const float* __restrict__ bias,
                  long long M, int V, int D, float* __restrict__ out) {
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int nv = D / 128;
  for (long long m = (long long)blockIdx.x * 8 + warp; m < M; m += (long long)gridDim.x * 8) {
    float4 v[kMaxVec];
#pragma unroll
    for (int i = 0; i < kMaxVec; ++i)
      if (i < nv) v[i] = *reinterpret_cast<const float4*>(h + (size_t)m * D + lane * 4 + i * 128);
    float mine = 0.f;
    for (int o = 0; o < V; ++o) {
      float acc = 0.f;
#pragma unroll
      for (int i = 0; i < kMaxVec; ++i)
        if (i < nv) {
          const float4 e = *reinterpret_cast<const float4*>(E + (size_t)o * D + lane * 4 + i * 128);
          acc += v[i].x * e.x + v[i].y * e.y + v[i].z * e.z + v[i].w * e.w;
        }
      acc = warp_sum(acc);
      if (lane == o) mine = acc + bias[o];
    }
    if (lane < V) out[(size_t)m * V + lane] = mine;
  }
}

// -------------------------------------------------------------------------------------------
// fp32 -> (hi, lo) for the three-MMA tf32 product: hi = x rounded TO NEAREST at tf32 precision (10 mantissa bits; low 13
// bits zero, so it is the same number whether the tensor core truncates or rounds its operands), lo = x - hi (exact in
// fp32, |lo| <= 2^-11 |x|, either sign) rounded to nearest at tf32 precision as well.  Representation error
// |x - hi - lo| <= 2^-23 |x| with no systematic sign (plain truncation would bias every term the same way: measured 6e-6
// instead of 6e-7 norm-relative on a 768-long dot product).
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ float rn_tf32(float v) {
  return __uint_as_float((__float_as_uint(v) + 0x00001000u) & 0xffffe000u);
}
__global__ void __launch_bounds__(256)
split_tf32_kernel(const float4* __restrict__ x, float4* __restrict__ hi, float4* __restrict__ lo, long long n4) {
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (long long)gridDim.x * blockDim.x) {
    const float4 v = x[i];
    const float4 h = make_float4(rn_tf32(v.x), rn_tf32(v.y), rn_tf32(v.z), rn_tf32(v.w));
    hi[i] = h;
    lo[i] = make_float4(rn_tf32(v.x - h.x), rn_tf32(v.y - h.y), rn_tf32(v.z - h.z), rn_tf32(v.w - h.w));
  }
}

int launch_split_tf32(const float* x, float* hi, float* lo, long long n, cudaStream_t st) {
  RNAMSM_REQUIRE(n % 4 == 0 && n > 0, "split_tf32: element count %lld must be a positive multiple of 4", n);
  const long long n4 = n / 4;
  const int blocks = (int)std::min<long long>((n4 + 255) / 256, 148LL * 16);
  ProfScope prof(KC_LAYERNORM, st);
  split_tf32_kernel<<<blocks, 256, 0, st>>>(reinterpret_cast<const float4*>(x), reinterpret_cast<float4*>(hi),
                                           reinterpret_cast<float4*>(lo), n4);
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------
// Debug: how close did a 16-bit tensor come to the edge of its range?  counters[0] += elements that are non-finite or
// sit AT the largest finite value (where cvt.rn.satfinite clamps), counters[1] = max(counters[1], bits of max |v|) as
// a float.  Grid-stride, one atomic per warp.
// -------------------------------------------------------------------------------------------
template <bool kFp16>
__global__ void __launch_bounds__(256)
range_scan_kernel(const uint16_t* __restrict__ p, long long n, unsigned long long* __restrict__ counters) {
  unsigned long long hits = 0;
  float mx = 0.f;
  for (long long i = (long long)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (long long)gridDim.x * blockDim.x) {
    const uint16_t a = p[i] & 0x7fffu;
    const bool edge = kFp16 ? a >= 0x7bffu : a >= 0x7f7fu;          // max finite (65504 / 3.39e38), inf, nan
    hits += edge ? 1u : 0u;
    const float v = kFp16 ? __half2float(__ushort_as_half(a)) : __uint_as_float((uint32_t)a << 16);
    if (!edge) mx = fmaxf(mx, v);
  }
  mx = warp_max(mx);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
  if ((threadIdx.x & 31) == 0) {
    if (hits) atomicAdd(&counters[0], hits);
    atomicMax(&counters[1], (unsigned long long)__float_as_uint(mx));
  }
}

int launch_range_scan(const void* p, long long n, int dtype, unsigned long long* counters, cudaStream_t st) {
  RNAMSM_REQUIRE(dtype == 1 || dtype == 2, "range_scan: 16-bit tensors only (dtype %d)", dtype);
  if (n <= 0) return 0;
  const int blocks = (int)std::min<long long>((n + 255) / 256, 148LL * 8);
  if (dtype == 2) range_scan_kernel<true><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(p), n, counters);
  else range_scan_kernel<false><<<blocks, 256, 0, st>>>(reinterpret_cast<const uint16_t*>(p), n, counters);
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// -------------------------------------------------------------------------------------------
// Launchers
// -------------------------------------------------------------------------------------------
int launch_embed_ln(const int64_t* tokens, int R, int C, const float* tok_emb, int vocab, const float* pos_emb,
                    int n_pos, const float* row_pos, const float* ln_w, const float* ln_b, int D, int pad_idx,
                    float eps, float* x_out, uint8_t* pad_out, cudaStream_t st) {
  RNAMSM_REQUIRE(D % 128 == 0 && D <= 128 * kMaxVec, "embed_layernorm: D=%d must be a multiple of 128 <= 1024", D);
  RNAMSM_REQUIRE(R > 0 && C > 0 && R <= 65535, "embed_layernorm: bad shape R=%d C=%d", R, C);
  dim3 grid(ceil_div(C, kColsPerBlock), R);
  ProfScope prof(KC_EMBED_LN, st);
  RNAMSM_CHECK_CUDA(launch_pdl(embed_ln_kernel, grid, dim3(kEmbWarps * 32), 0, st, tokens, R, C, tok_emb, vocab, pos_emb, n_pos,
                               row_pos, ln_w, ln_b, D, pad_idx, eps, x_out, pad_out));
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_layernorm(const float* x, const float* w, const float* b, void* y, int y_dtype, long long n_rows, int D,
                     float eps, cudaStream_t st, int tr_R, int tr_C, int x_dtype) {
  RNAMSM_REQUIRE(tr_C <= 0 || (long long)tr_R * tr_C == n_rows, "layernorm: transpose shape %d x %d != %lld rows", tr_R, tr_C, n_rows);
  RNAMSM_REQUIRE(tr_C <= 0 || (const void*)x != (const void*)y, "layernorm: the transposing form cannot run in place");
  RNAMSM_REQUIRE(D % 128 == 0 && D <= 128 * kMaxVec, "layernorm: D=%d must be a multiple of 128 <= 1024", D);
  if (n_rows <= 0) return 0;
  const int blocks = (int)std::min<long long>((n_rows + kLnWarps - 1) / kLnWarps, 148LL * 32);
  ProfScope prof(KC_LAYERNORM, st);
  if (x_dtype != 0) {   // 16-bit input -> fp32 output (the LM head's LayerNorm behind a 16-bit dense GEMM)
    RNAMSM_REQUIRE(y_dtype == 0 && tr_C <= 0 && (x_dtype == 1 || x_dtype == 2), "layernorm: 16-bit input supports fp32 output only");
    if (x_dtype == 1) RNAMSM_CHECK_CUDA(launch_pdl(layernorm_kernel<0, 1>, dim3(blocks), dim3(kLnWarps * 32), 0, st, x, w, b, y, n_rows, D, eps, 0, 0));
    else RNAMSM_CHECK_CUDA(launch_pdl(layernorm_kernel<0, 2>, dim3(blocks), dim3(kLnWarps * 32), 0, st, x, w, b, y, n_rows, D, eps, 0, 0));
  } else if (y_dtype == 1)
    RNAMSM_CHECK_CUDA(launch_pdl(layernorm_kernel<1, 0>, dim3(blocks), dim3(kLnWarps * 32), 0, st, x, w, b, y, n_rows, D, eps, tr_R, tr_C));
  else if (y_dtype == 2)
    RNAMSM_CHECK_CUDA(launch_pdl(layernorm_kernel<2, 0>, dim3(blocks), dim3(kLnWarps * 32), 0, st, x, w, b, y, n_rows, D, eps, tr_R, tr_C));
  else
    RNAMSM_CHECK_CUDA(launch_pdl(layernorm_kernel<0, 0>, dim3(blocks), dim3(kLnWarps * 32), 0, st, x, w, b, y, n_rows, D, eps, tr_R, tr_C));
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_add_layernorm(float* x, const void* delta, int delta_dtype, const float* w, const float* b, void* y,
                         int y_dtype, long long n_rows, int D, float eps, cudaStream_t st) {
  RNAMSM_REQUIRE(D % 128 == 0 && D <= 128 * kMaxVec, "add_layernorm: D=%d must be a multiple of 128 <= 1024", D);
  RNAMSM_REQUIRE((delta_dtype == 1 || delta_dtype == 2) && (y_dtype == 1 || y_dtype == 2), "add_layernorm: 16-bit delta / output only");
  if (n_rows <= 0) return 0;
  const int blocks = (int)std::min<long long>((n_rows + kLnWarps - 1) / kLnWarps, 148LL * 32);
  ProfScope prof(KC_LAYERNORM, st);
  const uint16_t* d = reinterpret_cast<const uint16_t*>(delta);
  uint16_t* yy = reinterpret_cast<uint16_t*>(y);
  if (delta_dtype == 1 && y_dtype == 1) add_layernorm_kernel<1, 1><<<blocks, kLnWarps * 32, 0, st>>>(x, d, w, b, yy, n_rows, D, eps);
  else if (delta_dtype == 1) add_layernorm_kernel<1, 2><<<blocks, kLnWarps * 32, 0, st>>>(x, d, w, b, yy, n_rows, D, eps);
  else if (y_dtype == 1) add_layernorm_kernel<2, 1><<<blocks, kLnWarps * 32, 0, st>>>(x, d, w, b, yy, n_rows, D, eps);
  else add_layernorm_kernel<2, 2><<<blocks, kLnWarps * 32, 0, st>>>(x, d, w, b, yy, n_rows, D, eps);
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_row_softmax(const float* partial, int n_splits, int H, int C, const uint8_t* key_pad, float logit_scale,
                       float* probs_out, void* probs_lp, int ld_lp, int dtype, cudaStream_t st) {
  RNAMSM_REQUIRE(n_splits >= 1 && H > 0 && C > 0, "row_softmax: bad shape");
  const long long rows = (long long)H * C;
  const int blocks = (int)((rows + kSmWarps - 1) / kSmWarps);
  if (probs_lp == nullptr) ld_lp = 0;
  ProfScope prof(KC_ROW_SOFTMAX, st);
#define RNAMSM_SOFTMAX_LAUNCH(LP, Q)                                                                                     \
  RNAMSM_CHECK_CUDA(launch_pdl(row_softmax_kernel<LP, Q>, dim3(blocks), dim3(kSmWarps * 32), 0, st, partial, n_splits, H, C, \
                               key_pad, logit_scale, probs_out, probs_lp, ld_lp))
#define RNAMSM_SOFTMAX_BY_C(LP)                            \
  do {                                                     \
    if (C <= 128) RNAMSM_SOFTMAX_LAUNCH(LP, 4);            \
    else if (C <= 256) RNAMSM_SOFTMAX_LAUNCH(LP, 8);       \
    else if (C <= 512) RNAMSM_SOFTMAX_LAUNCH(LP, 16);      \
    else if (C <= 1024) RNAMSM_SOFTMAX_LAUNCH(LP, 32);     \
    else RNAMSM_SOFTMAX_LAUNCH(LP, 0);                     \
  } while (0)
  if (dtype == 1) RNAMSM_SOFTMAX_BY_C(1);
  else if (dtype == 2) RNAMSM_SOFTMAX_BY_C(2);
  else RNAMSM_SOFTMAX_BY_C(0);
#undef RNAMSM_SOFTMAX_BY_C
#undef RNAMSM_SOFTMAX_LAUNCH
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

int launch_vocab_proj(const float* h, const float* E, const float* bias, long long M, int V, int D, float* out,
                      cudaStream_t st) {
  RNAMSM_REQUIRE(V <= 32 && D % 128 == 0 && D <= 128 * kMaxVec, "vocab_proj: V=%d D=%d unsupported", V, D);
  if (M <= 0) return 0;
  const int blocks = (int)std::min<long long>((M + 7) / 8, 148LL * 16);
  ProfScope prof(KC_VOCAB_PROJ, st);
  vocab_proj_kernel<<<blocks, 256, 0, st>>>(h, E, bias, M, V, D, out);
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace rnamsm
