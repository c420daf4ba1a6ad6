// Contact head on the device (SURVEY.md 8f "next" row 2): ContactPredictionHead.forward, modules.py:347-366,
// with utils/tensor.py:98-113 (symmetrize, average-product correction) fused with the 120 -> 1 logistic
// regression.  For every map k over the L x L block left after stripping BOS (/EOS):
//     S_k = A_k + A_k^T,  a1_k[i] = sum_j S_k[i,j],  a12_k = sum_i a1_k[i],
//     contacts[i,j] = sigmoid( b + sum_k w_k (S_k[i,j] - a1_k[i] a1_k[j] / a12_k) )
// HBM-bound: the maps are read twice for the margins (rows + columns) and twice for the output
// (tile + transposed tile); nothing of size [K, L, L] is written.  Deterministic (no atomics).
#include "../../include/rnamsm_b200.h"
#include "common.cuh"
#include "launch.h"

namespace rnamsm {

// grid (ceil(L/32), K), block (32, 8): a1[k, i] for the 32 indices of the tile = row sum + column sum.
__global__ void __launch_bounds__(256)
contact_margins_kernel(const float* __restrict__ maps, int C, int start, int L, float* __restrict__ a1) {
  __shared__ float red[8][33];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int k = blockIdx.y, i0 = blockIdx.x * 32;
  const float* A = maps + (size_t)k * C * C + (size_t)start * C + start;   // the L x L block, row stride C
  // column sums of columns i0 + tx: rows ty, ty + 8, ... (each warp reads 128 contiguous bytes per row)
  float cs = 0.f;
  if (i0 + tx < L)
    for (int j = ty; j < L; j += 8) cs += A[(size_t)j * C + i0 + tx];
  red[ty][tx] = cs;
  __syncthreads();
  // row sums of rows i0 + ty*4 + r: lanes stride over the columns
  float rs[4];
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty * 4 + r;
    float s = 0.f;
    if (i < L)
      for (int j = tx; j < L; j += 32) s += A[(size_t)i * C + j];
    rs[r] = warp_sum(s);
  }
  if (ty == 0 && i0 + tx < L) {
    float c = 0.f;
#pragma unroll
    for (int y = 0; y < 8; ++y) c += red[y][tx];
    red[0][tx] = c;
  }
  __syncthreads();
  if (tx < 4) {
    const int i = i0 + ty * 4 + tx;
    if (i < L) a1[(size_t)k * L + i] = rs[tx] + red[0][ty * 4 + tx];
  }
}

// one warp per map: a12[k] = sum_i a1[k, i]
__global__ void __launch_bounds__(128)
contact_total_kernel(const float* __restrict__ a1, int K, int L, float* __restrict__ a12) {
  const int k = blockIdx.x * 4 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (k >= K) return;
  float s = 0.f;
  for (int i = lane; i < L; i += 32) s += a1[(size_t)k * L + i];
  s = warp_sum(s);
  if (lane == 0) a12[k] = s;
}

// grid (ceil(L/32), ceil(L/32)), block (32, 8): thread (tx, ty) owns (i = i0 + ty + 8r, j = j0 + tx), r < 4.
__global__ void __launch_bounds__(256)
contact_out_kernel(const float* __restrict__ maps, int K, int C, int start, int L, const float* __restrict__ a1,
                   const float* __restrict__ a12, const float* __restrict__ w, const float* __restrict__ bias,
                   float* __restrict__ out) {
  __shared__ float T[32][33];
  __shared__ float ai[32], aj[32];
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int i0 = blockIdx.y * 32, j0 = blockIdx.x * 32;
  float acc[4] = {0.f, 0.f, 0.f, 0.f};
  for (int k = 0; k < K; ++k) {
    const float* A = maps + (size_t)k * C * C + (size_t)start * C + start;
    __syncthreads();                                   // previous iteration's smem fully consumed
#pragma unroll
    for (int r = 0; r < 4; ++r) {                      // transposed tile: rows j0.., columns i0..
      const int jj = j0 + ty + 8 * r, ii = i0 + tx;
      T[ty + 8 * r][tx] = (jj < L && ii < L) ? A[(size_t)jj * C + ii] : 0.f;
    }
    if (ty == 0) ai[tx] = (i0 + tx < L) ? a1[(size_t)k * L + i0 + tx] : 0.f;
    if (ty == 1) aj[tx] = (j0 + tx < L) ? a1[(size_t)k * L + j0 + tx] : 0.f;
    __syncthreads();
    const float wk = w[k], inv = 1.f / a12[k];
#pragma unroll
    for (int r = 0; r < 4; ++r) {
      const int i = i0 + ty + 8 * r, j = j0 + tx;
      if (i < L && j < L) {
        const float s = A[(size_t)i * C + j] + T[tx][ty + 8 * r];             // A_ij + A_ji
        acc[r] = fmaf(wk, s - ai[ty + 8 * r] * aj[tx] * inv, acc[r]);
      }
    }
  }
  const float b = bias ? bias[0] : 0.f;
#pragma unroll
  for (int r = 0; r < 4; ++r) {
    const int i = i0 + ty + 8 * r, j = j0 + tx;
    if (i < L && j < L) out[(size_t)i * L + j] = 1.f / (1.f + __expf(-(acc[r] + b)));
  }
}

}  // namespace rnamsm

using namespace rnamsm;

extern "C" int rnamsm_contact_head(const float* maps, int K, int C, int start, int L, const float* w, const float* bias,
                                   float* out, float* workspace, void* stream) {
  RNAMSM_REQUIRE(K > 0 && C > 0 && start >= 0 && L > 0 && start + L <= C, "contact_head: bad block start=%d L=%d of C=%d", start, L, C);
  RNAMSM_REQUIRE(workspace != nullptr, "contact_head: workspace of (K*L + K) floats required");
  cudaStream_t st = (cudaStream_t)stream;
  float* a1 = workspace;
  float* a12 = workspace + (size_t)K * L;
  const int tiles = ceil_div(L, 32);
  ProfScope prof(KC_CONTACT, st);
  contact_margins_kernel<<<dim3(tiles, K), dim3(32, 8), 0, st>>>(maps, C, start, L, a1);
  contact_total_kernel<<<ceil_div(K, 4), 128, 0, st>>>(a1, K, L, a12);
  contact_out_kernel<<<dim3(tiles, tiles), dim3(32, 8), 0, st>>>(maps, K, C, start, L, a1, a12, w, bias, out);
  count_launch(3);
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// SS-predictor input packing: what _downstream_tasks/SS builds on the CPU from the saved
// *_atp.npy -- DataProcess.feature_load (code/pre_processing/data_processing.py:32-48: outer concatenation of the
// ACGU one-hot [L,L,8] + maps transposed to [L,L,120]) followed by format_input_shape (data_fomat.py:38-59:
// -> [1,128,L,L] float) -- straight from the device-resident maps: channel c < 4: onehot(seq[i])[c];
// 4 <= c < 8: onehot(seq[j])[c-4]; c >= 8: map c-8 over the BOS-stripped block.  Pure HBM copy/gather.
// ---------------------------------------------------------------------------------------------------------
namespace rnamsm {
__global__ void __launch_bounds__(256)
ss_pack_kernel(const float* __restrict__ maps, int K, int C, int start, int L, const uint8_t* __restrict__ codes,
               float* __restrict__ out) {
  const int ch = blockIdx.z, i = blockIdx.y;
  float* dst = out + ((size_t)ch * L + i) * L;
  if (ch < 8) {
    const int ci = codes[i];
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x)
      dst[j] = (ch < 4) ? (ci == ch ? 1.f : 0.f) : (codes[j] == ch - 4 ? 1.f : 0.f);
  } else {
    const float* src = maps + (size_t)(ch - 8) * C * C + (size_t)(start + i) * C + start;
    for (int j = blockIdx.x * blockDim.x + threadIdx.x; j < L; j += gridDim.x * blockDim.x) dst[j] = src[j];
  }
}
}  // namespace rnamsm

extern "C" int rnamsm_ss_pack(const float* maps, int K, int C, int start, int L, const uint8_t* seq_codes, float* out,
                              void* stream) {
  RNAMSM_REQUIRE(K > 0 && L > 0 && start >= 0 && start + L <= C && L <= 65535 && K + 8 <= 65535, "ss_pack: bad shape");
  dim3 grid(rnamsm::ceil_div(L, 256), L, K + 8);
  rnamsm::ss_pack_kernel<<<grid, 256, 0, (cudaStream_t)stream>>>(maps, K, C, start, L, seq_codes, out);
  rnamsm::count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

// ---------------------------------------------------------------------------------------------------------
// RSA-predictor input packing (_downstream_tasks/RSA/predict.py:131-141): the
// [4 + D + 1, L] tensor (z-scored one-hot | z-scored embedding | ones, transposed to channels-first) written
// from the device-resident hidden states.  The reference computes the embedding z-score in fp32
// ((emb - mu) / std, numpy float32 arrays) and the one-hot z-score in float64 before the final cast to fp32;
// both are reproduced operation for operation (__fsub_rn / __fdiv_rn: no contraction, IEEE division; the
// eight possible one-hot values are rounded from float64 on the host), so the result is bit-identical.
// 32 x 32 tiles through shared memory: reads coalesced along the feature axis, writes along the residue axis.
// ---------------------------------------------------------------------------------------------------------
namespace rnamsm {
struct RsaOneHot { float v[8]; };  // [channel][seq[i] == channel]

__global__ void __launch_bounds__(256)
rsa_pack_kernel(const float* __restrict__ emb, int ld, int L, int D, int n_oh, const uint8_t* __restrict__ codes,
                const float* __restrict__ mu, const float* __restrict__ sd, RsaOneHot oh, float* __restrict__ out) {
  __shared__ float tile[32][33];
  const int i0 = blockIdx.x * 32, ch0 = blockIdx.y * 32, n_ch = n_oh + D + 1;
  const int tx = threadIdx.x, ty = threadIdx.y;
  const int ch = ch0 + tx;
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int i = i0 + ty + k;
    float v = 0.f;
    if (i < L && ch < n_ch) {
      if (ch < n_oh) {
        v = oh.v[ch * 2 + (codes[i] == ch ? 1 : 0)];
      } else if (ch < n_oh + D) {
        const int f = ch - n_oh;
        v = __fdiv_rn(__fsub_rn(emb[(size_t)i * ld + f], mu[f]), sd[f]);
      } else {
        v = 1.f;
      }
    }
    tile[ty + k][tx] = v;
  }
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 32; k += 8) {
    const int c = ch0 + ty + k, i = i0 + tx;
    if (c < n_ch && i < L) out[(size_t)c * L + i] = tile[tx][ty + k];
  }
}
}  // namespace rnamsm

extern "C" int rnamsm_rsa_pack(const float* emb, int ld, int L, int D, const uint8_t* seq_codes, const double* mu_oh,
                               const double* std_oh, const float* mu_emb, const float* std_emb, float* out,
                               void* stream) {
  RNAMSM_REQUIRE(L > 0 && D > 0 && ld >= D, "rsa_pack: bad shape (L=%d D=%d ld=%d)", L, D, ld);
  RNAMSM_REQUIRE((mu_oh == nullptr) == (std_oh == nullptr), "rsa_pack: mu_oh and std_oh go together");
  RNAMSM_REQUIRE(mu_oh == nullptr || seq_codes != nullptr, "rsa_pack: one-hot channels need the sequence codes");
  rnamsm::RsaOneHot oh{};
  const int n_oh = mu_oh ? 4 : 0;
  for (int c = 0; c < n_oh; ++c)
    for (int hit = 0; hit < 2; ++hit) oh.v[c * 2 + hit] = (float)(((double)hit - mu_oh[c]) / std_oh[c]);
  dim3 grid(rnamsm::ceil_div(L, 32), rnamsm::ceil_div(n_oh + D + 1, 32));
  RNAMSM_REQUIRE(grid.y <= 65535, "rsa_pack: D too large");
  rnamsm::rsa_pack_kernel<<<grid, dim3(32, 8), 0, (cudaStream_t)stream>>>(emb, ld, L, D, n_oh, seq_codes, mu_emb, std_emb,
                                                                         oh, out);
  rnamsm::count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}
