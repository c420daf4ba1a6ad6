// K5s, 16-bit path: the whole tied row attention of a SHORT alignment (C <= 128 columns) in ONE launch
// (modules.py:752-821: compute_attention_weights + compute_attention_update, align_scaling :713-715).
//
//   phase 1   partial[s,h,i,j] = sum_{r in chunk s} sum_d q[r,i,h,d] k[r,j,h,d]            modules.py:774
//   -- grid barrier --
//   phase 2   P[h,i,:] = softmax_j( logit_scale * sum_s partial[s,h,i,:]  (key mask -> -10000) )   modules.py:780-788
//             (fp32 rows -> the exported map, 16-bit rows -> the operand of phase 3)
//   -- grid barrier --
//   phase 3   ctx[r,i,h,:] = sum_j P[h,i,j] v[r,j,h,:]   for the rows r of the same chunk            modules.py:797
//
// Why a kernel of its own: with C = 36 (the 2DRB_1 example, BASELINE configs[0]) the general path -- the 256-row pair
// tile of umma_gemm_kernel<TIED>, one TMA box per MSA row and k-block, a separate softmax launch, the AV GEMM --
// spends 23.6 + 6.1 + 29.3 us per layer (ncu, profiles/r02b_launches_cfg1.md) on 2 GFLOP and 110 MB of traffic:
// it is bound by the latency of ~90 dependent 9 KiB TMA boxes per CTA and by three launches.  Here a CTA is one
// (head, row chunk): H x n_chunks <= #SMs CTAs, all co-resident (cooperative launch), each
//   - streams its rows' q and k head slices as [GR rows][CP columns][64] boxes (several MSA rows per TMA box, so a
//     box carries up to 32 KiB instead of 9), multiplies them as M128 x N=CP x K16 tcgen05.mma (single CTA;
//     accumulator rows >= CP are never read: the A descriptor simply runs on into the neighbouring rows' tiles),
//   - takes part in the softmax of its head (rows i = chunk, chunk + n_chunks, ...: one warp per row, the split sums in
//     a fixed order, so the result does not depend on timing),
//   - loads the head's 16-bit P once (TMA, K-major A operand) and runs P V over its own rows with V read in place as an
//     MN-major B operand (four MSA rows = 256 accumulator columns per MMA group), 16-bit rows leaving through swizzled
//     staging + TMA stores clipped at column C by the hardware.
// The two grid-wide barriers are one global counter (release add / acquire spin) that the last CTA to leave resets.
//
// Warp roles: warp 0 lane 0 = TMA producer, warp 1 = TMEM allocation + MMA issuer (lane 0), warps 4..11 = epilogue
// (TMEM lane quadrant = warp & 3, two warps per quadrant taking alternate MSA rows of a group, each with a two-deep
// staging ring: with C = 36 only quadrant 0 and 1 hold rows, and one warp per quadrant with one staging box spent
// ~1000 cycles per MSA row waiting for its previous TMA store -- 22 of the first version's 43 us); all 12 warps take
// softmax rows in phase 2.
#include <stdlib.h>

#include <atomic>
#include <mutex>

#include "../../include/rnamsm_b200.h"
#include "common.cuh"
#include "launch.h"

namespace rnamsm {

namespace {

constexpr int kMaxC = 128;
constexpr int kThreads = 384;
constexpr int kEpiWarps = 8;
constexpr int RING_BYTES = 128 * 1024;       // operand ring, re-cut for phase 3
constexpr int P_BYTES = 2 * 128 * 128;       // P: two K atoms (64 keys each) of 128 rows x 128 B
constexpr int STG_BYTES = kEpiWarps * 2 * 32 * 128;   // two 32-row x 128 B staging boxes per epilogue warp
constexpr int OFF_RING = 0, OFF_P = OFF_RING + RING_BYTES, OFF_STG = OFF_P + P_BYTES, OFF_BAR = OFF_STG + STG_BYTES;
constexpr int kSmem = OFF_BAR + 512 + 1024;
constexpr int kMaxStages = 8;
constexpr int kTmemCols = 512;               // phase 1: S in [0, CP); phase 3: two 256-column accumulators
constexpr int GV = 4;                        // MSA rows per P V group (4 x 64 head dims = 256 accumulator columns)
static_assert(kSmem <= 227 * 1024, "row_attn_short: shared memory budget");

struct ShortArgs {
  int R, C, H, CP;            // CP = C rounded up to 16: box rows, MMA N of phase 1, MMA K of phase 3
  int GR;                     // MSA rows per q / k box
  int n_chunks, rows_per_chunk;
  int ldp;                    // row pitch of probs_lp in elements
  float* partial;             // [n_chunks, H, C, C] fp32
  float* map;                 // [H, C, C] fp32 probabilities (the exported row-attention map)
  void* probs_lp;             // [H, C, ldp] 16-bit probabilities
  const uint8_t* key_pad;     // [C] or nullptr (MSA row 0 of the padding mask, modules.py:780-784)
  float logit_scale;          // 1 / sqrt(R) (q carries 64^-1/2)
  unsigned* sync;             // grid barrier counter, zero between launches
  long long* trace;           // debug (RNAMSM_SHORT_TRACE=1): globaltimer of CTA 0 / thread 0 at the phase boundaries
};

__device__ __forceinline__ void trace_mark(const ShortArgs& a, int slot) {
  if (a.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0) {
    long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    a.trace[slot] = t;
  }
}

__device__ __forceinline__ unsigned ld_acquire_gpu(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void red_release_gpu(unsigned* p) {
  asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(p) : "memory");
}

// All CTAs are co-resident (cooperative launch).  Arrival n of the launch waits for the counter to reach `target`.
__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned target) {
  fence_proxy_async_all();                   // every thread's generic-proxy global writes -> other CTAs' TMA (async proxy) reads
  __syncthreads();
  if (threadIdx.x == 0) {
    red_release_gpu(counter);
    unsigned spins = 0;
    while (ld_acquire_gpu(counter) < target)
      if (++spins > (1u << 24)) __trap();    // a missing CTA must not hang the GPU
    fence_proxy_async_all();
  }
  __syncthreads();
}

__device__ __forceinline__ uint32_t pack16_sel(float lo, float hi, bool fp16) { return fp16 ? pack_f16(lo, hi) : pack_bf16(lo, hi); }

template <bool kFp16>
__global__ void __launch_bounds__(kThreads, 1)
row_attn_short_kernel(const __grid_constant__ CUtensorMap tm_qk, const __grid_constant__ CUtensorMap tm_p,
                      const __grid_constant__ CUtensorMap tm_v, const __grid_constant__ CUtensorMap tm_o, const ShortArgs a) {
  extern __shared__ uint8_t smem_raw[];
  trace_mark(a, 0);
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + OFF_BAR);
  uint64_t* full1 = bars;                       // [kMaxStages] phase 1 ring
  uint64_t* empty1 = bars + kMaxStages;
  uint64_t* full3 = bars + 2 * kMaxStages;      // [kMaxStages] phase 3 ring
  uint64_t* empty3 = bars + 3 * kMaxStages;
  uint64_t* s_done = bars + 4 * kMaxStages;     // partial logits of this CTA complete in TMEM
  uint64_t* p_full = s_done + 1;                // P landed in shared memory
  uint64_t* acc_full = p_full + 1;              // [2]
  uint64_t* acc_empty = acc_full + 2;           // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(acc_empty + 2);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int h = blockIdx.x % a.H, chunk = blockIdx.x / a.H;
  const int r0 = chunk * a.rows_per_chunk, r1 = min(a.R, r0 + a.rows_per_chunk);
  const int D = a.H * 64;
  const int CP = a.CP, GR = a.GR;
  const int tile_bytes = CP * 128;                              // one MSA row's [CP][64] head slice
  const int stg1 = 2 * GR * tile_bytes, ns1 = min(kMaxStages, RING_BYTES / stg1);
  const int stg3 = GV * tile_bytes, ns3 = min(kMaxStages, RING_BYTES / stg3);
  const int n_grp1 = (r1 - r0 + GR - 1) / GR, n_grp3 = (r1 - r0 + GV - 1) / GV;
  const int n_katoms = (CP + 63) / 64;

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_qk);
    tma_prefetch_desc(&tm_p);
    tma_prefetch_desc(&tm_v);
    tma_prefetch_desc(&tm_o);
    for (int s = 0; s < kMaxStages; ++s) {
      mbar_init(&full1[s], 1);
      mbar_init(&empty1[s], 1);
      mbar_init(&full3[s], 1);
      mbar_init(&empty3[s], 1);
    }
    mbar_init(s_done, 1);
    mbar_init(p_full, 1);
    for (int b = 0; b < 2; ++b) {
      mbar_init(&acc_full[b], 1);
      mbar_init(&acc_empty[b], kEpiWarps);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  trace_mark(a, 1);

  // =========================== phase 1: partial logits of (head h, rows r0..r1) ===========================
  if (warp == 0 && lane == 0) {
    int stage = 0;
    uint32_t phase = 0;
    for (int g = 0; g < n_grp1; ++g) {
      mbar_wait(&empty1[stage], phase ^ 1);
      uint8_t* sq = smem + OFF_RING + stage * stg1;
      mbar_expect_tx(&full1[stage], stg1);
      tma_load_3d(sq, &tm_qk, &full1[stage], h * 64, 0, r0 + g * GR);                        // q: [GR][CP][64]
      tma_load_3d(sq + GR * tile_bytes, &tm_qk, &full1[stage], D + h * 64, 0, r0 + g * GR);  // k
      if (++stage == ns1) { stage = 0; phase ^= 1; }
    }
    // V does not depend on P: once this CTA's last logit MMA has read the ring, the first P V groups are fetched
    // behind the two grid barriers and the softmax
    mbar_wait(s_done, 0);
    trace_mark(a, 2);
    for (int g = 0; g < min(n_grp3, ns3); ++g) {
      mbar_expect_tx(&full3[g], stg3);
      tma_load_3d(smem + OFF_RING + g * stg3, &tm_v, &full3[g], 2 * D + h * 64, 0, r0 + g * GV);   // v: [GV][CP][64]
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_16(128, CP, kFp16 ? 1 : 0, 0, 0);
    int stage = 0;
    uint32_t phase = 0, first = 0;
    for (int g = 0; g < n_grp1; ++g) {
      mbar_wait(&full1[stage], phase);
      tc_fence_after();
      const uint32_t qa = smem_u32(smem + OFF_RING + stage * stg1);
      const uint32_t ka = qa + GR * tile_bytes;
      const int nr = min(GR, r1 - (r0 + g * GR));               // rows of the NEXT chunk in the box are not ours
      for (int rr = 0; rr < nr; ++rr) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          umma_16(tmem_base, make_smem_desc_sw128(qa + rr * tile_bytes + k * 32, 16, 1024),
                  make_smem_desc_sw128(ka + rr * tile_bytes + k * 32, 16, 1024), idesc, first);
          first = 1;
        }
      }
      umma_commit(&empty1[stage]);
      if (++stage == ns1) { stage = 0; phase ^= 1; }
    }
    umma_commit(s_done);
  } else if (warp >= 4) {
    const int quad = warp & 3, sub = (warp - 4) >> 2;
    const int i = quad * 32 + lane;
    mbar_wait(s_done, 0);
    tc_fence_after();
    float* dst = a.partial + (((size_t)chunk * a.H + h) * a.C + i) * a.C;
    if (quad * 32 < a.C) {                                      // warp-uniform
      for (int c = sub * 32; c < CP; c += 64) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(quad * 32) << 16) + c, v);
        tmem_ld_wait();
        if (i < a.C) {
#pragma unroll
          for (int e = 0; e < 32; ++e)
            if (c + e < a.C) dst[c + e] = __uint_as_float(v[e]);
        }
      }
    }
    tc_fence_before();
  }
  trace_mark(a, 3);
  grid_barrier(a.sync, gridDim.x);
  trace_mark(a, 4);

  // =========================== phase 2: softmax rows i = chunk, chunk + n_chunks, ... of head h ===========
  {
    const size_t split_stride = (size_t)a.H * a.C * a.C;
    for (int i = chunk + warp * a.n_chunks; i < a.C; i += (kThreads / 32) * a.n_chunks) {
      const float* src = a.partial + ((size_t)h * a.C + i) * a.C;
      float v[kMaxC / 32];
      float mx = -INFINITY;
#pragma unroll
      for (int q = 0; q < kMaxC / 32; ++q) {
        const int j = q * 32 + lane;
        float acc = -INFINITY;
        if (j < a.C) {
          acc = 0.f;
          for (int s = 0; s < a.n_chunks; s += 4) {             // four loads in flight, summed in split order
            float t[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) t[u] = s + u < a.n_chunks ? __ldcg(src + (s + u) * split_stride + j) : 0.f;
#pragma unroll
            for (int u = 0; u < 4; ++u)
              if (s + u < a.n_chunks) acc += t[u];
          }
          acc *= a.logit_scale;
          if (a.key_pad && a.key_pad[j]) acc = -10000.f;        // masked_fill, modules.py:780-784
        }
        v[q] = acc;
        mx = fmaxf(mx, acc);
      }
      mx = warp_max(mx);
      float sum = 0.f;
#pragma unroll
      for (int q = 0; q < kMaxC / 32; ++q)
        if (q * 32 + lane < a.C) sum += __expf(v[q] - mx);
      sum = warp_sum(sum);
      const float inv = 1.f / sum;
      float* dst = a.map + ((size_t)h * a.C + i) * a.C;
      uint16_t* dlp = reinterpret_cast<uint16_t*>(a.probs_lp) + ((size_t)h * a.C + i) * a.ldp;
#pragma unroll
      for (int q = 0; q < kMaxC / 32; ++q) {
        const int j = q * 32 + lane;
        const float p = j < a.C ? __expf(v[q] - mx) * inv : 0.f;
        if (j < a.C) dst[j] = p;
        if (j < a.ldp) dlp[j] = kFp16 ? __half_as_ushort(__float2half_rn(p)) : __bfloat16_as_ushort(__float2bfloat16(p));
      }
    }
  }
  trace_mark(a, 5);
  grid_barrier(a.sync, 2 * gridDim.x);
  trace_mark(a, 6);

  // =========================== phase 3: ctx rows r0..r1 of head h = P V ===================================
  if (warp == 0 && lane == 0) {
    mbar_expect_tx(p_full, n_katoms * tile_bytes);
    for (int ka = 0; ka < n_katoms; ++ka) tma_load_3d(smem + OFF_P + ka * (128 * 128), &tm_p, p_full, ka * 64, 0, h);
    int stage = 0;
    uint32_t phase = 1;                          // the first pass over the ring was issued before the barriers
    for (int g = ns3; g < n_grp3; ++g) {
      mbar_wait(&empty3[stage], phase ^ 1);
      mbar_expect_tx(&full3[stage], stg3);
      tma_load_3d(smem + OFF_RING + stage * stg3, &tm_v, &full3[stage], 2 * D + h * 64, 0, r0 + g * GV);
      if (++stage == ns3) { stage = 0; phase ^= 1; }
    }
  } else if (warp == 1 && lane == 0) {
    const uint32_t idesc = make_idesc_16(128, GV * 64, kFp16 ? 1 : 0, 0, 1);   // P (K-major) x V (MN-major)
    const uint32_t pa = smem_u32(smem + OFF_P);
    mbar_wait(p_full, 0);
    int stage = 0, acc = 0;
    uint32_t phase = 0, acc_phase = 0;
    for (int g = 0; g < n_grp3; ++g) {
      mbar_wait(&acc_empty[acc], acc_phase ^ 1);
      mbar_wait(&full3[stage], phase);
      tc_fence_after();
      const uint32_t va = smem_u32(smem + OFF_RING + stage * stg3);
      for (int k = 0; k < CP / 16; ++k)   // 16 keys per instruction; V: 64-wide N chunks (MSA rows) tile_bytes apart
        umma_16(tmem_base + acc * (GV * 64), make_smem_desc_sw128(pa + (k >> 2) * (128 * 128) + (k & 3) * 32, 16, 1024),
                make_smem_desc_sw128(va + k * 2048, (uint32_t)tile_bytes, 1024), idesc, (uint32_t)(k != 0));
      umma_commit(&empty3[stage]);
      umma_commit(&acc_full[acc]);
      if (++stage == ns3) { stage = 0; phase ^= 1; }
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
  } else if (warp >= 4) {
    const int quad = warp & 3, sub = (warp - 4) >> 2;
    uint8_t* bufs = smem + OFF_STG + (warp - 4) * (2 * 32 * 128);
    const bool has_rows = quad * 32 < a.C;
    int acc = 0, n_st = 0;
    uint32_t acc_phase = 0;
    for (int g = 0; g < n_grp3; ++g) {
      mbar_wait(&acc_full[acc], acc_phase);
      tc_fence_after();
      const int nr = min(GV, r1 - (r0 + g * GV));
      if (has_rows) {
        for (int rr = sub; rr < nr; rr += 2, ++n_st) {
          uint8_t* buf = bufs + (n_st & 1) * (32 * 128);
          uint32_t v0[32], v1[32], w[32];
          const uint32_t taddr = tmem_base + ((uint32_t)(quad * 32) << 16) + acc * (GV * 64) + rr * 64;
          tmem_ld_32x32(taddr, v0);
          tmem_ld_32x32(taddr + 32, v1);
          tmem_ld_wait();
#pragma unroll
          for (int k = 0; k < 32; k += 2) {
            w[k / 2] = pack16_sel(__uint_as_float(v0[k]), __uint_as_float(v0[k + 1]), kFp16);
            w[16 + k / 2] = pack16_sel(__uint_as_float(v1[k]), __uint_as_float(v1[k + 1]), kFp16);
          }
          if (lane == 0) asm volatile("cp.async.bulk.wait_group.read 1;" ::: "memory");   // the store before last has left this box
          __syncwarp();
          uint8_t* row = buf + lane * 128;
#pragma unroll
          for (int c = 0; c < 8; ++c)
            *reinterpret_cast<uint4*>(row + ((c ^ (lane & 7)) << 4)) = make_uint4(w[4 * c], w[4 * c + 1], w[4 * c + 2], w[4 * c + 3]);
          fence_proxy_async_smem();
          __syncwarp();
          if (lane == 0) {
            tma_store_3d(&tm_o, buf, h * 64, quad * 32, r0 + g * GV + rr);   // rows i >= C are clipped by the hardware
            bulk_commit();
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&acc_empty[acc]);
      acc ^= 1;
      if (acc == 0) acc_phase ^= 1;
    }
    if (lane == 0) bulk_wait_all0();
  }

  if (warp == 0 && lane == 0) trace_mark(a, 7);     // producer done issuing
  tc_fence_before();
  __syncthreads();
  trace_mark(a, 8);
  if (warp == 1) tmem_dealloc(tmem_base, kTmemCols);
  // the counter goes back to zero for the next launch: every CTA passed the second barrier before its third arrival
  if (threadIdx.x == 0 && atomicAdd(a.sync, 1u) == 3u * gridDim.x - 1u) atomicExch(a.sync, 0u);
}

// Grid-barrier counters: a ring of 64 per device (128 B apart), one per launch in round-robin order, so launches of this
// kernel that overlap in time (two streams, two model instances) never share a counter; each is zero between uses
// (the last CTA of a launch resets its own).
unsigned* sync_counter() {
  constexpr int kRing = 64, kStride = 32;                    // 32 x 4 B = 128 B
  static unsigned* ptr[64] = {};
  static std::atomic<unsigned> next[64];
  int dev = 0;
  cudaGetDevice(&dev);
  if (dev < 0 || dev >= 64) return nullptr;
  if (ptr[dev] == nullptr) {
    static std::mutex mu;
    std::lock_guard<std::mutex> lock(mu);
    if (ptr[dev] == nullptr) {
      unsigned* p = nullptr;
      if (cudaMalloc(&p, kRing * kStride * sizeof(unsigned)) != cudaSuccess) return nullptr;
      cudaMemset(p, 0, kRing * kStride * sizeof(unsigned));
      ptr[dev] = p;
    }
  }
  return ptr[dev] + (size_t)(next[dev].fetch_add(1, std::memory_order_relaxed) % kRing) * kStride;
}

int num_sms_dev() {
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  return sms > 0 ? sms : 148;
}

template <bool kFp16>
int launch_short(const CUtensorMap& tqk, const CUtensorMap& tp, const CUtensorMap& tv, const CUtensorMap& to, const ShortArgs& a,
                 cudaStream_t st) {
  static bool attr_set = false;
  if (!attr_set) {
    RNAMSM_CHECK_CUDA(cudaFuncSetAttribute(row_attn_short_kernel<kFp16>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem));
    attr_set = true;
  }
  static int want_trace = -1;
  if (want_trace < 0) want_trace = getenv("RNAMSM_SHORT_TRACE") ? 1 : 0;
  ShortArgs args = a;
  if (want_trace) {
    static long long* dbuf = nullptr;
    if (!dbuf) { cudaMalloc(&dbuf, 16 * sizeof(long long)); }
    args.trace = dbuf;
  }
  void* params[] = {(void*)&tqk, (void*)&tp, (void*)&tv, (void*)&to, (void*)&args};
  {
    ProfScope prof(KC_ROW_SHORT, st);
    RNAMSM_CHECK_CUDA(cudaLaunchCooperativeKernel((const void*)row_attn_short_kernel<kFp16>, dim3(a.H * a.n_chunks),
                                                  dim3(kThreads), params, (size_t)kSmem, st));
  }
  count_launch();
  if (want_trace) {       // debug: phase boundaries of CTA 0 in ns since its first instruction
    long long h[16];
    RNAMSM_CHECK_CUDA(cudaStreamSynchronize(st));
    RNAMSM_CHECK_CUDA(cudaMemcpy(h, args.trace, sizeof(h), cudaMemcpyDeviceToHost));
    fprintf(stderr, "row_attn_short R=%d C=%d: prologue %lld | logits %lld | epi1 %lld | barrier1 %lld | softmax %lld | barrier2 %lld | "
            "PV issue %lld | drain %lld | total %lld ns\n", a.R, a.C, h[1] - h[0], h[2] - h[1], h[3] - h[2], h[4] - h[3], h[5] - h[4],
            h[6] - h[5], h[7] - h[6], h[8] - h[7], h[8] - h[0]);
  }
  return 0;
}

}  // namespace

// Row chunks (= split count of the partial logits) the one-launch path uses for this shape; 0 = not applicable
// (C > 128, more heads than SMs, or RNAMSM_ROW_SHORT=0).
int row_attn_short_chunks(int R, int C, int H) {
  static int enabled = -1;
  if (enabled < 0) {
    const char* e = getenv("RNAMSM_ROW_SHORT");
    enabled = (e && e[0] == '0') ? 0 : 1;
  }
  if (!enabled || C > kMaxC || C < 1 || R < 1 || H < 1) return 0;
  // CTAs that can be co-resident (one per SM unless the device is partitioned or shares its shared memory): resolved once;
  // if the kernel cannot be resident at all the three-kernel chain stays in use
  static int resident = -1;
  if (resident < 0) {
    int per_sm = 0;
    cudaFuncSetAttribute(row_attn_short_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    cudaFuncSetAttribute(row_attn_short_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kSmem);
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, row_attn_short_kernel<true>, kThreads, kSmem) != cudaSuccess) {
      per_sm = 0;
    }
    cudaGetLastError();     // (no device, as in the CPU-only build container: the chain stays selected, nothing is launched)
    resident = std::max(0, per_sm) * num_sms_dev();
  }
  const int sms = std::min(num_sms_dev(), resident);
  if (H > sms) return 0;
  const int want = std::min(R, sms / H);
  const int rpc = ceil_div(R, want);
  return ceil_div(R, rpc);                  // every chunk non-empty
}

int launch_row_attn_short_16(const void* qkv, int R, int C, int H, int fp16, float* partial, int n_chunks, const uint8_t* key_pad,
                             float logit_scale, float* map, void* probs_lp, int ldp, void* ctx, cudaStream_t st) {
  RNAMSM_REQUIRE(C >= 1 && C <= kMaxC, "row_attn_short: C=%d outside [1, %d]", C, kMaxC);
  RNAMSM_REQUIRE(n_chunks >= 1 && n_chunks <= R && (long long)H * n_chunks <= num_sms_dev() && row_attn_short_chunks(R, C, H) > 0,
                 "row_attn_short: %d heads x %d chunks do not fit the device (R=%d)", H, n_chunks, R);
  RNAMSM_REQUIRE(ldp % 8 == 0 && ldp >= C, "row_attn_short: ldp=%d must be a multiple of 8 and >= C=%d", ldp, C);
  RNAMSM_REQUIRE(partial && map && probs_lp && ctx, "row_attn_short: null buffer");
  ShortArgs a{};
  a.R = R; a.C = C; a.H = H;
  a.CP = (C + 15) / 16 * 16;
  a.GR = std::max(1, std::min(8, 128 / a.CP));
  a.n_chunks = n_chunks;
  a.rows_per_chunk = ceil_div(R, n_chunks);
  RNAMSM_REQUIRE((n_chunks - 1) * a.rows_per_chunk < R, "row_attn_short: n_chunks=%d leaves an empty chunk for R=%d", n_chunks, R);
  a.ldp = ldp;
  a.partial = partial; a.map = map; a.probs_lp = probs_lp; a.key_pad = key_pad; a.logit_scale = logit_scale;
  a.sync = sync_counter();
  RNAMSM_REQUIRE(a.sync != nullptr, "row_attn_short: could not allocate the barrier counter");
  const int ld = 3 * H * 64;
  const int in_dt = fp16 ? TMAP_F16 : TMAP_BF16;
  CUtensorMap tqk, tp, tv, to;
  {
    uint64_t dims[3] = {(uint64_t)ld, (uint64_t)C, (uint64_t)R};
    uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)C * ld * 2};
    uint32_t box_qk[3] = {64, (uint32_t)a.CP, (uint32_t)a.GR};
    uint32_t box_v[3] = {64, (uint32_t)a.CP, GV};
    if (encode_tmap(&tqk, in_dt, qkv, 3, dims, strides, box_qk)) return 3;
    if (encode_tmap(&tv, in_dt, qkv, 3, dims, strides, box_v)) return 3;
  }
  {
    uint64_t dims[3] = {(uint64_t)C, (uint64_t)C, (uint64_t)H};
    uint64_t strides[2] = {(uint64_t)ldp * 2, (uint64_t)C * ldp * 2};
    uint32_t box[3] = {64, (uint32_t)a.CP, 1};
    if (encode_tmap(&tp, in_dt, probs_lp, 3, dims, strides, box)) return 3;
  }
  {
    uint64_t dims[3] = {(uint64_t)(H * 64), (uint64_t)C, (uint64_t)R};
    uint64_t strides[2] = {(uint64_t)H * 64 * 2, (uint64_t)C * H * 64 * 2};
    uint32_t box[3] = {64, 32, 1};
    if (encode_tmap(&to, in_dt, ctx, 3, dims, strides, box)) return 3;
  }
  return fp16 ? launch_short<true>(tqk, tp, tv, to, a, st) : launch_short<false>(tqk, tp, tv, to, a, st);
}

}  // namespace rnamsm
