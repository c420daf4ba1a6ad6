// bf16 tensor-core path: one persistent, warp-specialised tcgen05 GEMM serving the three dense
// contractions of the forward.  Operands are staged in shared memory by TMA (SWIZZLE_128B),
// multiplied by tcgen05.mma (cta_group::1, M=128, N=256, K=16 per instruction, fp32 accumulators
// in TMEM, double-buffered so the epilogue of tile t overlaps the MMAs of tile t+1), and read
// back with tcgen05.ld for the fused epilogues.
//
//   DENSE : out[m,n] = epi( sum_k x[m,k] W[n,k] + b[n] )                 nn.Linear sites,
//           modules.py:760-766 (q/k/v), :799 (out_proj), :424-426 (fc1/GELU/fc2), :314 (lm dense)
//   TIED  : partial[s,h,i,j] = sum_{r in split s} sum_d q[r,i,h,d] k[r,j,h,d]   modules.py:774
//           (K-loop walks MSA rows; one 64-wide head slice per k-block, split-K over row ranges)
//   AV    : ctx[r,i,h,:] = sum_j P[h,i,j] v[r,j,h,:]                            modules.py:797
//           (P tile is the stationary A operand; B = V read in place as an MN-major operand,
//            four MSA rows x 64 head dims per 256-wide N tile)
//
// Warp roles (192 threads): warp 0 = TMA producer (one elected lane), warp 1 = MMA issuer (one
// elected lane), warps 2..5 = epilogue (TMEM lane quadrant = warp_idx % 4); warp 2 also owns the
// TMEM allocation.  Pipelines: smem full/empty ring (kStages), TMEM full/empty (2 accumulators).
#include "../../include/rnamsm_b200.h"
#include "common.cuh"
#include "launch.h"

namespace rnamsm {

namespace {

constexpr int BLOCK_M = 128;
constexpr int BLOCK_N = 256;
constexpr int BLOCK_K = 64;                      // one SWIZZLE_128B atom of bf16 along K
constexpr int UMMA_K = 16;
constexpr int kStages = 4;
constexpr int A_BYTES = BLOCK_M * BLOCK_K * 2;   // 16 KiB
constexpr int B_BYTES = BLOCK_N * BLOCK_K * 2;   // 32 KiB
constexpr int STAGE_BYTES = A_BYTES + B_BYTES;   // 48 KiB
constexpr int kTmemCols = 512;                   // 2 accumulator stages x 256 fp32 columns
constexpr int kEpiWarps = 4;
constexpr int kThreads = 64 + 32 * kEpiWarps;
constexpr int kSmemBytes = kStages * STAGE_BYTES + 256 + 1024;  // tiles + barriers + align slack

enum { V_DENSE = 0, V_TIED = 1, V_AV = 2 };

struct GemmArgs {
  int m_tiles, n_tiles, batches, splits;
  int k_blocks;          // DENSE: K/64, AV: ceil(C/64); TIED: rows of the split (computed per tile)
  int M, N;              // logical bounds for the epilogue (DENSE: M x N; TIED: C x C; AV: C x R)
  int R, C, H;
  int rows_per_split;
  int epi_kind;
  const float* bias;
  float q_scale;
  int q_cols;
  const uint8_t* row_mask;
  void* out;
  int ld_out;
};

struct TileCoord {
  int m0, n0;            // element offsets of the tile inside the logical M / N extents
  int batch, split;
  int kb_begin, kb_count;
};

template <int kVariant>
__device__ __forceinline__ TileCoord decode_tile(const GemmArgs& g, int tile) {
  TileCoord t;
  if (kVariant == V_DENSE) {
    t.n0 = (tile % g.n_tiles) * BLOCK_N;
    t.m0 = (tile / g.n_tiles) * BLOCK_M;
    t.batch = 0; t.split = 0; t.kb_begin = 0; t.kb_count = g.k_blocks;
  } else if (kVariant == V_TIED) {
    t.n0 = (tile % g.n_tiles) * BLOCK_N;  tile /= g.n_tiles;
    t.m0 = (tile % g.m_tiles) * BLOCK_M;  tile /= g.m_tiles;
    t.split = tile % g.splits;
    t.batch = tile / g.splits;
    t.kb_begin = t.split * g.rows_per_split;
    t.kb_count = min(g.R, t.kb_begin + g.rows_per_split) - t.kb_begin;
  } else {
    t.m0 = (tile % g.m_tiles) * BLOCK_M;  tile /= g.m_tiles;
    t.n0 = (tile % g.n_tiles) * 4;        // first MSA row of the 4-row group
    t.batch = tile / g.n_tiles;
    t.split = 0; t.kb_begin = 0; t.kb_count = g.k_blocks;
  }
  return t;
}

// ---- epilogues: 32 consecutive accumulator columns of one row per call ---------------------------
template <int kVariant>
__device__ __forceinline__ void epilogue_chunk(const GemmArgs& g, const TileCoord& t, int row_in_tile, int chunk,
                                               const uint32_t (&acc)[32]) {
  if (kVariant == V_DENSE) {
    const long long m = (long long)t.m0 + row_in_tile;
    const int n = t.n0 + chunk * 32;
    if (m >= g.M || n >= g.N) return;
    float v[32];
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
      const float4 b = *reinterpret_cast<const float4*>(g.bias + n + i);
      v[i] = __uint_as_float(acc[i]) + b.x;
      v[i + 1] = __uint_as_float(acc[i + 1]) + b.y;
      v[i + 2] = __uint_as_float(acc[i + 2]) + b.z;
      v[i + 3] = __uint_as_float(acc[i + 3]) + b.w;
    }
    if (g.epi_kind == RNAMSM_EPI_BIAS_RESIDUAL) {
      float* dst = reinterpret_cast<float*>(g.out) + (size_t)m * g.ld_out + n;
#pragma unroll
      for (int i = 0; i < 32; i += 4) {
        float4 r = *reinterpret_cast<float4*>(dst + i);
        r.x += v[i]; r.y += v[i + 1]; r.z += v[i + 2]; r.w += v[i + 3];
        *reinterpret_cast<float4*>(dst + i) = r;
      }
      return;
    }
    if (g.epi_kind == RNAMSM_EPI_BIAS_GELU) {
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] = gelu_erf(v[i]);
    } else if (n < g.q_cols) {  // q_cols is a multiple of 32: a chunk is entirely q or not
      const float s = (g.row_mask && g.row_mask[m]) ? 0.f : g.q_scale;
#pragma unroll
      for (int i = 0; i < 32; ++i) v[i] *= s;
    }
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.out) + (size_t)m * g.ld_out + n;
#pragma unroll
    for (int i = 0; i < 32; i += 8) {
      uint4 pk = make_uint4(pack_bf16(v[i], v[i + 1]), pack_bf16(v[i + 2], v[i + 3]),
                            pack_bf16(v[i + 4], v[i + 5]), pack_bf16(v[i + 6], v[i + 7]));
      *reinterpret_cast<uint4*>(dst + i) = pk;
    }
  } else if (kVariant == V_TIED) {
    const int i = t.m0 + row_in_tile;
    const int j0 = t.n0 + chunk * 32;
    if (i >= g.C || j0 >= g.C) return;
    float* dst = reinterpret_cast<float*>(g.out) + (((size_t)t.split * g.H + t.batch) * g.C + i) * g.C + j0;
    if ((g.C & 3) == 0 && j0 + 32 <= g.C) {
#pragma unroll
      for (int k = 0; k < 32; k += 4)
        *reinterpret_cast<float4*>(dst + k) = make_float4(__uint_as_float(acc[k]), __uint_as_float(acc[k + 1]),
                                                          __uint_as_float(acc[k + 2]), __uint_as_float(acc[k + 3]));
    } else {
#pragma unroll
      for (int k = 0; k < 32; ++k)
        if (j0 + k < g.C) dst[k] = __uint_as_float(acc[k]);
    }
  } else {
    const int i = t.m0 + row_in_tile;
    const int r = t.n0 + (chunk >> 1);
    if (i >= g.C || r >= g.R) return;
    __nv_bfloat16* dst = reinterpret_cast<__nv_bfloat16*>(g.out) + ((size_t)r * g.C + i) * g.ld_out + t.batch * 64 +
                         (chunk & 1) * 32;
#pragma unroll
    for (int k = 0; k < 32; k += 8) {
      uint4 pk = make_uint4(pack_bf16(__uint_as_float(acc[k]), __uint_as_float(acc[k + 1])),
                            pack_bf16(__uint_as_float(acc[k + 2]), __uint_as_float(acc[k + 3])),
                            pack_bf16(__uint_as_float(acc[k + 4]), __uint_as_float(acc[k + 5])),
                            pack_bf16(__uint_as_float(acc[k + 6]), __uint_as_float(acc[k + 7])));
      *reinterpret_cast<uint4*>(dst + k) = pk;
    }
  }
}

template <int kVariant>
__global__ void __launch_bounds__(kThreads, 1)
umma_gemm_kernel(const __grid_constant__ CUtensorMap tmap_a, const __grid_constant__ CUtensorMap tmap_b,
                 const GemmArgs g) {
  extern __shared__ uint8_t smem_raw[];
  // SWIZZLE_128B tiles need 1024 B alignment (descriptor base_offset = 0).
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + kStages * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + kStages;
  uint64_t* tmem_full = empty_bar + kStages;
  uint64_t* tmem_empty = tmem_full + 2;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(tmem_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int total_tiles = g.m_tiles * g.n_tiles * g.batches * g.splits;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tmap_a);
    tma_prefetch_desc(&tmap_b);
  }
  if (warp == 1 && elect_one()) {
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tmem_full[a], 1);
      mbar_init(&tmem_empty[a], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ================================ TMA producer ================================
    if (elect_one()) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile<kVariant>(g, tile);
        for (int kb = 0; kb < t.kb_count; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sa = smem + stage * STAGE_BYTES;
          uint8_t* sb = sa + A_BYTES;
          mbar_expect_tx(&full_bar[stage], STAGE_BYTES);
          if (kVariant == V_DENSE) {
            tma_load_3d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, t.m0, 0);
            tma_load_3d(sb, &tmap_b, &full_bar[stage], kb * BLOCK_K, t.n0, 0);
          } else if (kVariant == V_TIED) {
            const int r = t.kb_begin + kb;
            tma_load_3d(sa, &tmap_a, &full_bar[stage], t.batch * 64, t.m0, r);
            tma_load_3d(sb, &tmap_b, &full_bar[stage], (g.H + t.batch) * 64, t.n0, r);
          } else {
            tma_load_3d(sa, &tmap_a, &full_bar[stage], kb * BLOCK_K, t.m0, t.batch);
            tma_load_3d(sb, &tmap_b, &full_bar[stage], (2 * g.H + t.batch) * 64, kb * BLOCK_K, t.n0);
          }
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ==================================
    if (elect_one()) {
      constexpr uint32_t idesc = make_idesc_bf16(BLOCK_M, BLOCK_N, 0, kVariant == V_AV ? 1 : 0);
      int stage = 0, acc = 0;
      uint32_t phase = 0, acc_phase = 0;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const TileCoord t = decode_tile<kVariant>(g, tile);
        mbar_wait(&tmem_empty[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (int kb = 0; kb < t.kb_count; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * STAGE_BYTES);
          const uint32_t b_addr = a_addr + A_BYTES;
#pragma unroll
          for (int k = 0; k < BLOCK_K / UMMA_K; ++k) {
            // A: K-major, 128 B rows, 8-row groups 1024 B apart; +32 B per 16-element k step.
            const uint64_t adesc = make_smem_desc_sw128(a_addr + k * (UMMA_K * 2), 16, 1024);
            uint64_t bdesc;
            if (kVariant == V_AV)  // B: MN-major [4 r][64 j][64 d]: 64-wide N chunks 8 KiB apart,
              bdesc = make_smem_desc_sw128(b_addr + k * (UMMA_K * 128), 64 * 128, 1024);  // 8-row k groups 1 KiB
            else
              bdesc = make_smem_desc_sw128(b_addr + k * (UMMA_K * 2), 16, 1024);
            umma_bf16(d_tmem, adesc, bdesc, idesc, (uint32_t)((kb | k) != 0));
          }
          umma_commit(&empty_bar[stage]);  // smem slot reusable once these MMAs have read it
          if (++stage == kStages) { stage = 0; phase ^= 1; }
        }
        umma_commit(&tmem_full[acc]);      // accumulator complete -> epilogue
        acc ^= 1;
        if (acc == 0) acc_phase ^= 1;
      }
    }
  } else {
    // ================================ epilogue ====================================
    const int quad = warp & 3;  // TMEM lanes [32*quad, 32*quad+32) are accessible to this warp
    int acc = 0;
    uint32_t acc_phase = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
      const TileCoord t = decode_tile<kVariant>(g, tile);
      mbar_wait(&tmem_full[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + acc * BLOCK_N + ((uint32_t)(quad * 32) << 16);
      const int row_in_tile = quad * 32 + lane;
#pragma unroll 1
      for (int chunk = 0; chunk < BLOCK_N / 32; ++chunk) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + chunk * 32, v);
        tmem_ld_wait();
        epilogue_chunk<kVariant>(g, t, row_in_tile, chunk, v);
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&tmem_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, kTmemCols);
}

int num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
    if (n <= 0) n = 148;
  }
  return n;
}

template <int kVariant>
int launch_variant(const CUtensorMap& ta, const CUtensorMap& tb, const GemmArgs& g, int prof_class, cudaStream_t st) {
  RNAMSM_CHECK_CUDA(cudaFuncSetAttribute(umma_gemm_kernel<kVariant>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         kSmemBytes));
  const long long total = (long long)g.m_tiles * g.n_tiles * g.batches * g.splits;
  RNAMSM_REQUIRE(total > 0 && total < (1LL << 31), "umma_gemm: tile count %lld out of range", total);
  const int grid = (int)std::min<long long>(total, num_sms());
  ProfScope prof(prof_class, st);
  umma_gemm_kernel<kVariant><<<grid, kThreads, kSmemBytes, st>>>(ta, tb, g);
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

// ---------------------------------------------------------------------------------------------
int launch_linear_bf16(const void* x, const void* W, long long M, int N, int K, const LinearEpilogue& epi, void* out,
                       cudaStream_t st) {
  RNAMSM_REQUIRE(M > 0 && M < (1LL << 31), "linear_bf16: M=%lld out of range", M);
  RNAMSM_REQUIRE(N % 32 == 0 && K % BLOCK_K == 0 && K >= BLOCK_K, "linear_bf16: N=%d must be a multiple of 32, K=%d of 64", N, K);
  RNAMSM_REQUIRE(epi.bias != nullptr, "linear_bf16: bias required");
  RNAMSM_REQUIRE(epi.q_cols % 32 == 0, "linear_bf16: q_cols=%d must be a multiple of 32", epi.q_cols);
  CUtensorMap ta, tb;
  {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)M, 1};
    uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)M * K * 2};
    uint32_t box[3] = {BLOCK_K, BLOCK_M, 1};
    if (encode_tmap_bf16(&ta, x, 3, dims, strides, box)) return 3;
  }
  {
    uint64_t dims[3] = {(uint64_t)K, (uint64_t)N, 1};
    uint64_t strides[2] = {(uint64_t)K * 2, (uint64_t)N * K * 2};
    uint32_t box[3] = {BLOCK_K, BLOCK_N, 1};
    if (encode_tmap_bf16(&tb, W, 3, dims, strides, box)) return 3;
  }
  GemmArgs g{};
  g.m_tiles = ceil_div(M, BLOCK_M);
  g.n_tiles = ceil_div(N, BLOCK_N);
  g.batches = 1; g.splits = 1;
  g.k_blocks = K / BLOCK_K;
  g.M = (int)M; g.N = N;
  g.epi_kind = epi.kind; g.bias = epi.bias; g.q_scale = epi.q_scale; g.q_cols = epi.q_cols; g.row_mask = epi.row_mask;
  g.out = out; g.ld_out = N;
  return launch_variant<V_DENSE>(ta, tb, g, linear_class(epi.kind, N, K), st);
}

int row_logits_splits_bf16(int R, int C, int H) {
  const long long tiles = (long long)H * ceil_div(C, BLOCK_M) * ceil_div(C, BLOCK_N);
  int want = (int)std::max<long long>(1, num_sms() / std::max<long long>(1, tiles));
  want = std::min(want, std::max(1, R / 8));  // keep >= 8 rows (k-blocks) per split
  want = std::max(1, std::min(want, R));
  const int rps = ceil_div(R, want);
  return ceil_div(R, rps);  // every split non-empty
}

int launch_row_logits_bf16(const void* qkv, int R, int C, int H, float* partial, int n_splits, cudaStream_t st) {
  RNAMSM_REQUIRE(n_splits >= 1 && n_splits <= R, "row_logits_bf16: n_splits=%d out of range for R=%d", n_splits, R);
  const int rps = ceil_div(R, n_splits);
  RNAMSM_REQUIRE((n_splits - 1) * rps < R, "row_logits_bf16: n_splits=%d leaves an empty split for R=%d", n_splits, R);
  const int ld = 3 * H * 64;
  CUtensorMap ta, tb;
  uint64_t dims[3] = {(uint64_t)ld, (uint64_t)C, (uint64_t)R};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)C * ld * 2};
  uint32_t box_a[3] = {BLOCK_K, BLOCK_M, 1};
  uint32_t box_b[3] = {BLOCK_K, BLOCK_N, 1};
  if (encode_tmap_bf16(&ta, qkv, 3, dims, strides, box_a)) return 3;
  if (encode_tmap_bf16(&tb, qkv, 3, dims, strides, box_b)) return 3;
  GemmArgs g{};
  g.m_tiles = ceil_div(C, BLOCK_M);
  g.n_tiles = ceil_div(C, BLOCK_N);
  g.batches = H; g.splits = n_splits;
  g.rows_per_split = rps;
  g.R = R; g.C = C; g.H = H; g.M = C; g.N = C;
  g.out = partial;
  return launch_variant<V_TIED>(ta, tb, g, KC_ROW_LOGITS, st);
}

int launch_row_av_bf16(const void* probs, int ldp, const void* qkv, int R, int C, int H, void* ctx, cudaStream_t st) {
  RNAMSM_REQUIRE(ldp % 8 == 0 && ldp >= C, "row_av_bf16: ldp=%d must be a multiple of 8 and >= C=%d", ldp, C);
  const int ld = 3 * H * 64;
  CUtensorMap ta, tb;
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)C, (uint64_t)H};
    uint64_t strides[2] = {(uint64_t)ldp * 2, (uint64_t)C * ldp * 2};
    uint32_t box[3] = {BLOCK_K, BLOCK_M, 1};
    if (encode_tmap_bf16(&ta, probs, 3, dims, strides, box)) return 3;
  }
  {
    uint64_t dims[3] = {(uint64_t)ld, (uint64_t)C, (uint64_t)R};
    uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)C * ld * 2};
    uint32_t box[3] = {64, BLOCK_K, 4};
    if (encode_tmap_bf16(&tb, qkv, 3, dims, strides, box)) return 3;
  }
  GemmArgs g{};
  g.m_tiles = ceil_div(C, BLOCK_M);
  g.n_tiles = ceil_div(R, 4);
  g.batches = H; g.splits = 1;
  g.k_blocks = ceil_div(C, BLOCK_K);
  g.R = R; g.C = C; g.H = H; g.M = C; g.N = R;
  g.out = ctx; g.ld_out = H * 64;
  return launch_variant<V_AV>(ta, tb, g, KC_ROW_AV, st);
}

}  // namespace rnamsm
