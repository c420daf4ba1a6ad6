// K7, 16-bit path, production kernel: warp-specialised flash-style column attention on tcgen05
// (modules.py:896-923).  For every alignment column c and head h the MSA depth R is the sequence
// axis:   ctx[i,c,h,:] = sum_j softmax_j(q[i,c,h,:] . k[j,c,h,:]) v[j,c,h,:]     (q pre-scaled by 64^-1/2)
// The R x R probabilities are never materialised (the reference keeps [H,C,B,R,R] per layer).
//
// Persistent CTAs (one per SM).  A work item is (column c, head h, block of NT x 128 queries): NT
// independent 128-row query tiles that share one K/V stream.  Each tile is its own pipeline
// (its own MMA-issuing thread, softmax warpgroup, S / O accumulators in TMEM and P buffer).
// Measured (ncu + ablations): time per step ~= 1000 + 525 * NT cycles whatever is removed
// (MMAs, TMA loads, P stores, proxy fences, even the ex2 themselves) -- the softmax warps are bound by
// the XU (MUFU.EX2 at 16/clk/SM = 2048 cycles per 4-tile step, 55-66 % busy) and by issue slots
// (~2200 per 4-tile step) at once, with 4-5 warps per scheduler to hide the FFMA -> MUFU -> FADD
// chains.  NT = 4 fills TMEM (4 x (64 S + 64 O) = 512 columns); NT = 2 serves shallow MSAs (R <= 256).
// Variants tried and measured equal or slower: two threads per row, one issuing thread for all
// tiles, P through tensor memory (tcgen05.st + A-from-TMEM PV MMA), back-off in the role threads,
// offsetting the tiles' phases by 400-1600 cycles.  What did cost time was the ITEM boundary
// (t_item ~= (3.3 + steps) * t_step): direct per-thread global stores of O were 19 % of the kernel at
// R = 512 -- O now leaves through shared memory and one asynchronous TMA store per warp.
// Items are numbered with the query block innermost, so CTAs working at the same moment share K/V
// through L2.
// Round 2, measured and dropped: (a) an XU token ring between the NT softmax warps of an SM sub-partition (one warp in
// its exponential section at a time, token handed on 3/4 of the way through): 593 vs 622 TF/s at 512 x 256, 666 vs
// 731 at 4096 x 128, 638 vs 704 at 1024 x 1024 -- exclusive use of the XU makes it worse, so the warps are not
// convoying on it; (b) writing the section as three straight runs (32 FFMA2, 64 MUFU pinned with asm volatile, 32 packs
// + adds): ptxas re-interleaves them into the same MUFU, MUFU, FADD2, F2FP pattern; (c) ex2.approx.ftz.f16x2 runs at 16
// RESULTS/clk/SM like the f32 form (tools/micro/mufu_bench.cu on a B200: 15.96 vs 15.98), so it halves MUFU
// instructions but not XU time; (d) programmatic dependent launch for the prologue: no change.  ncu at 4096 x 128
// (profiles/r02_ncu_col_attn_cfg4.md): XU 70.5 %, tensor 34.6 %, issue 58 %; top stalls wait 24.7 % (fixed-latency
// dependencies), MIO 17.6 %, long scoreboard 13.5 % -- every softmax warp is latency-bound at ~2900 cycles per 64-key
// step with 4 of them per sub-partition (TMEM and the 96-register budget allow no more).
//
//   warp 0            TMA producer: Q tiles of the item and a K/V ring, 3-D boxes (64 d x 1 column
//                     x rows) straight out of the packed q|k|v activation
//   warps 1..3 (+0)   MMA issuers, one thread per tile: S_t = Q_t K_j^T (M128 N64 K64, fp32
//                     in TMEM) issued ONE STEP AHEAD of the softmax, O_t (+)= P_t V_j accumulating in
//                     TMEM; warp 1 also owns the TMEM allocation
//   warps 4..4+4NT-1  softmax, one warpgroup per tile (thread = query row = TMEM lane):
//                     S -> registers, running max with LAZY rescaling (O and l are only rescaled when
//                     the row max grows by more than 2^8, FlashAttention-4 style: the stale reference
//                     max cancels in O / l), p = ex2(s * log2e - m), P -> smem as the 16-bit K-major
//                     SWIZZLE_128B A operand of the PV MMA.  O never leaves TMEM until the item ends.
#include <stdlib.h>

#include "../../include/rnamsm_b200.h"
#include "common.cuh"
#include "launch.h"

namespace rnamsm {

namespace {

constexpr int BQ = 128, BKV = 64, HD = 64;
constexpr int Q_BYTES = BQ * HD * 2;        // 16 KiB per tile
constexpr int KV_BYTES = BKV * HD * 2;      // 8 KiB each for K and V
constexpr int P_BYTES = BQ * BKV * 2;       // 16 KiB per tile
constexpr float kLog2e = 1.4426950408889634f;
constexpr bool kPolyPairs = false;          // true: 25 % of the exponentials as polynomials (see exp2_poly2); measured
                                            // identical time (452 / 685 / 578 TF/s either way): the XU is not the limiter
constexpr float kRescaleThreshold = 8.0f;   // log2 units: P stays <= 2^8, exact in fp16 / bf16 range

// G = item groups per CTA: the NT tiles form G groups of NT / G tiles; a group shares one K/V ring and works
// through its own sequence of items, so with G = 2 the two groups' item boundaries (Q reload, PV(last), O
// read-out) fall at different times and each hides behind the other group's steps.  G = 1 is the production
// configuration; G = 2 (RNAMSM_COL_GROUPS=2) is written but NOT yet measured or validated on hardware.
template <int NT, int G = 1>
struct Cfg {
  static constexpr int kQBufs = NT == 2 ? 2 : 1;        // Q double-buffered across items when smem allows
  static constexpr int kKvStages = NT == 2 ? 4 : 3;
  static constexpr int OFF_Q = 0;                        // [kQBufs][NT tiles]
  static constexpr int OFF_KV = OFF_Q + kQBufs * NT * Q_BYTES;   // [G groups][stages][K | V]
  static constexpr int OFF_P = OFF_KV + G * kKvStages * 2 * KV_BYTES;
  static constexpr int OFF_BAR = OFF_P + NT * P_BYTES;
  static constexpr int kSmem = OFF_BAR + 512 + 1024;
  static constexpr int kWarps = 4 + 4 * NT;              // 4 role warps + one softmax warpgroup per tile
  static constexpr int kThreads = 32 * kWarps;           // NT = 4: 640 threads -> 96 registers each
  static constexpr int kTmemCols = NT * 128;
};

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// 2^x for a packed pair on the FMA / ALU pipes instead of the XU (FlashAttention-4's trick): round-to-nearest split
// x = n + f via the 1.5 * 2^23 magic add, degree-3 minimax polynomial for 2^f on [-0.5, 0.5] (max relative error
// 7.5e-5, below the 16-bit rounding of P), n added into the exponent field.  ~10 issue slots per pair and no MUFU:
// one pair in four goes this way so that the XU (64 ex2 per row per step otherwise) and the issue slots balance.
__device__ __forceinline__ void exp2_poly2(uint64_t x, float& r0, float& r1) {
  float x0, x1;
  f32x2_unpack(x, x0, x1);
  const uint64_t xc = f32x2_pack(fmaxf(x0, -126.f), fmaxf(x1, -126.f));
  const uint64_t t = f32x2_add(xc, f32x2_pack(12582912.f, 12582912.f));
  const uint64_t fl = f32x2_add(t, f32x2_pack(-12582912.f, -12582912.f));
  const uint64_t fr = f32x2_fma(fl, f32x2_pack(-1.f, -1.f), xc);
  uint64_t p = f32x2_fma(f32x2_pack(0.05517144873738289f, 0.05517144873738289f), fr,
                         f32x2_pack(0.2426108419895172f, 0.2426108419895172f));
  p = f32x2_fma(p, fr, f32x2_pack(0.6932609677314758f, 0.6932609677314758f));
  p = f32x2_fma(p, fr, f32x2_pack(0.9999281167984009f, 0.9999281167984009f));
  float p0, p1, t0, t1;
  f32x2_unpack(p, p0, p1);
  f32x2_unpack(t, t0, t1);
  r0 = __uint_as_float(__float_as_uint(p0) + (__float_as_uint(t0) << 23));
  r1 = __uint_as_float(__float_as_uint(p1) + (__float_as_uint(t1) << 23));
}

struct Item { int c, h, i0; };

template <int NT>
__device__ __forceinline__ Item decode_item(int item, int nqb, int H) {
  Item it;
  it.i0 = (item % nqb) * (NT * BQ);
  item /= nqb;
  it.h = item % H;
  it.c = item / H;
  return it;
}

template <int NT, bool kFp16, int G>
__global__ void __launch_bounds__(Cfg<NT, G>::kThreads, 1)
col_attn_ws_kernel(const __grid_constant__ CUtensorMap tm_q, const __grid_constant__ CUtensorMap tm_kv,
                   const __grid_constant__ CUtensorMap tm_o, int R, int C, int H, int col_major, int n_items,
                   const uint8_t* __restrict__ pad) {
  constexpr int fp16 = kFp16 ? 1 : 0;
  using K = Cfg<NT, G>;
  constexpr int kKvStages = K::kKvStages;
  constexpr int TG = NT / G;                    // tiles per item group
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + K::OFF_BAR);
  uint64_t* q_full = bars;                      // [G][2]
  uint64_t* q_empty = bars + 2 * G;             // [G][2]
  uint64_t* kv_full = bars + 4 * G;             // [G][4]
  uint64_t* kv_empty = bars + 8 * G;            // [G][4]
  uint64_t* s_full = bars + 12 * G;             // [4] per tile
  uint64_t* s_free = bars + 12 * G + 4;
  uint64_t* p_full = bars + 12 * G + 8;
  uint64_t* pv_done = bars + 12 * G + 12;
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12 * G + 16);

  const int warp = __shfl_sync(0xffffffffu, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const int D = H * HD;
  const int nblk = (R + BKV - 1) / BKV;
  const int nqb = (R + TG * BQ - 1) / (TG * BQ);

  if (warp == 0 && elect_one()) {
    tma_prefetch_desc(&tm_q);
    tma_prefetch_desc(&tm_kv);
    for (int b = 0; b < 2 * G; ++b) {
      mbar_init(&q_full[b], 1);
      mbar_init(&q_empty[b], TG);               // one commit per MMA issuer of the group
    }
    if (G == 2)
      for (int b = 4; b < 8; ++b) {
        mbar_init(&kv_full[b], 1);
        mbar_init(&kv_empty[b], TG);
      }
    for (int b = 0; b < 4; ++b) {
      mbar_init(&kv_full[b], 1);
      mbar_init(&kv_empty[b], TG);              // one commit per MMA issuer of the group
      mbar_init(&s_full[b], 1);
      mbar_init(&s_free[b], 4);
      mbar_init(&p_full[b], 4);
      mbar_init(&pv_done[b], 1);
    }
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(tmem_ptr, K::kTmemCols);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  pdl_launch_dependents();
  pdl_wait();                // RNAMSM_PDL=1: the prologue above overlapped the predecessor's tail

  // Single-thread roles: TMA producer = warp 0 lane 0 (group 1's producer, G = 2: warp 1 lane 1); MMA issuer of
  // tile t = lane 0 of warp 1 + t for t < 3 and lane 1 of warp 0 for t = 3 (two roles on divergent lanes of one
  // warp: independent thread scheduling interleaves them, and it keeps the CTA at 20 warps = 96 registers per thread).
  const int mma_tile = (lane == 0 && warp >= 1 && warp <= 3 && warp - 1 < NT) ? warp - 1
                       : ((NT == 4 && warp == 0 && lane == 1) ? 3 : -1);

  if ((warp == 0 && lane == 0) || (G == 2 && warp == 1 && lane == 1)) {
    // ================================ TMA producer (one per item group) ===========
    {
      const int gr = (G == 2 && warp == 1) ? 1 : 0;
      int kv_stage = 0;
      uint32_t kv_phase = 0;
      int li = 0;
      for (int item = blockIdx.x * G + gr; item < n_items; item += gridDim.x * G, ++li) {
        const Item it = decode_item<TG>(item, nqb, H);
        const int qb = K::kQBufs == 2 ? (li & 1) : 0;
        const int quse = K::kQBufs == 2 ? (li >> 1) : li;     // how many times this buffer was used before
        mbar_wait_relaxed(&q_empty[gr * 2 + qb], (quse & 1) ^ 1);
        mbar_expect_tx(&q_full[gr * 2 + qb], TG * Q_BYTES);
#pragma unroll
        for (int tl = 0; tl < TG; ++tl)
          tma_load_3d(smem + K::OFF_Q + (qb * NT + gr * TG + tl) * Q_BYTES, &tm_q, &q_full[gr * 2 + qb], it.h * HD,
                      col_major ? it.i0 + tl * BQ : it.c, col_major ? it.c : it.i0 + tl * BQ);
        for (int j = 0; j < nblk; ++j) {
          mbar_wait_relaxed(&kv_empty[gr * 4 + kv_stage], kv_phase ^ 1);
          uint8_t* sk = smem + K::OFF_KV + (gr * kKvStages + kv_stage) * 2 * KV_BYTES;
          mbar_expect_tx(&kv_full[gr * 4 + kv_stage], 2 * KV_BYTES);
          tma_load_3d(sk, &tm_kv, &kv_full[gr * 4 + kv_stage], D + it.h * HD, col_major ? j * BKV : it.c,
                      col_major ? it.c : j * BKV);
          tma_load_3d(sk + KV_BYTES, &tm_kv, &kv_full[gr * 4 + kv_stage], 2 * D + it.h * HD, col_major ? j * BKV : it.c,
                      col_major ? it.c : j * BKV);
          if (++kv_stage == kKvStages) { kv_stage = 0; kv_phase ^= 1; }
        }
      }
    }
  } else if (mma_tile >= 0) {
    // ================================ MMA issuer of tile t =========================
    {
      const int t = mma_tile;
      const int gr = t / TG;                                        // item group of this tile
      const uint32_t idesc_s = make_idesc_16(BQ, BKV, fp16, 0, 0);  // Q (K-major) x K (K-major)
      const uint32_t idesc_o = make_idesc_16(BQ, HD, fp16, 0, 1);   // P (K-major) x V (MN-major)
      const int first = (int)blockIdx.x * G + gr, stride = (int)gridDim.x * G;
      const int my_items = (G == 1 || first < n_items) ? (n_items - first + stride - 1) / stride : 0;
      const long long total_steps = (long long)my_items * nblk;
      const uint32_t tmem_S = tmem_base + t * BKV, tmem_O = tmem_base + NT * BKV + t * HD;
      const uint32_t pa = smem_u32(smem + K::OFF_P + t * P_BYTES);
      // cursor of the S issue (runs one step ahead of the PV issue)
      long long gs = 0;
      int s_li = 0, s_j = 0, s_stage = 0;
      uint32_t s_phase = 0;
      auto issue_s = [&]() {
        const int qb = K::kQBufs == 2 ? (s_li & 1) : 0;
        const int quse = K::kQBufs == 2 ? (s_li >> 1) : s_li;
        if (s_j == 0) mbar_wait_relaxed(&q_full[gr * 2 + qb], quse & 1);
        mbar_wait_relaxed(&kv_full[gr * 4 + s_stage], s_phase);
        if (gs > 0) mbar_wait_relaxed(&s_free[t], (uint32_t)((gs - 1) & 1));   // softmax t has S(gs-1) in registers
        tc_fence_after();
        const uint32_t ka = smem_u32(smem + K::OFF_KV + (gr * kKvStages + s_stage) * 2 * KV_BYTES);
        const uint32_t qa = smem_u32(smem + K::OFF_Q + (qb * NT + t) * Q_BYTES);
#pragma unroll
        for (int k = 0; k < HD / 16; ++k)
          umma_16(tmem_S, make_smem_desc_sw128(qa + k * 32, 16, 1024), make_smem_desc_sw128(ka + k * 32, 16, 1024),
                  idesc_s, (uint32_t)(k != 0));
        umma_commit(&s_full[t]);
        if (s_j == nblk - 1) umma_commit(&q_empty[gr * 2 + qb]);   // this tile's Q fully consumed
        ++gs;
        if (++s_j == nblk) { s_j = 0; ++s_li; }
        if (++s_stage == kKvStages) { s_stage = 0; s_phase ^= 1; }
      };
      int o_j = 0, o_stage = 0;
      if (total_steps > 0) issue_s();
      for (long long g = 0; g < total_steps; ++g) {
        // S(g+1) goes out before PV(g) -- except across an item boundary: the next item's first S waits for its Q
        // tile, and PV(last) (which the softmax warps need for their epilogue) must not queue behind that wait
        const bool s_new_item = K::kQBufs == 1 && s_j == 0;      // (double-buffered Q is already there: keep S ahead)
        if (g + 1 < total_steps && !s_new_item) issue_s();
        const uint32_t va = smem_u32(smem + K::OFF_KV + (gr * kKvStages + o_stage) * 2 * KV_BYTES + KV_BYTES);
        mbar_wait_relaxed(&p_full[t], (uint32_t)(g & 1));                  // P_t(g) in smem, O_t rescaled if needed
        tc_fence_after();
#pragma unroll
        for (int k = 0; k < BKV / 16; ++k)
          umma_16(tmem_O, make_smem_desc_sw128(pa + k * 32, 16, 1024), make_smem_desc_sw128(va + k * 2048, 8192, 1024),
                  idesc_o, (uint32_t)((o_j | k) != 0));
        umma_commit(&pv_done[t]);
        umma_commit(&kv_empty[gr * 4 + o_stage]);                  // this tile is done with K_j and V_j
        if (++o_j == nblk) o_j = 0;
        if (++o_stage == kKvStages) o_stage = 0;
        if (g + 1 < total_steps && s_new_item) issue_s();
      }
    }
  } else if (warp >= 4 && warp < 4 + 4 * NT) {
    // ================================ softmax warpgroups ==========================
    const int t = (warp - 4) >> 2;                 // tile
    const int gr = t / TG, tl = t % TG;            // item group, tile inside the group's item
    const int quad = warp & 3;                     // TMEM lane quadrant of this warp
    const int row = quad * 32 + lane;              // query row inside the tile == TMEM lane
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t tmem_S = tmem_base + t * BKV + lane_off;
    const uint32_t tmem_O = tmem_base + NT * BKV + t * HD + lane_off;
    uint8_t* prow = smem + K::OFF_P + t * P_BYTES + row * 128;
    const float neg_masked = -10000.f;             // masked_fill value, modules.py:911-915
    long long g = 0;
    for (int item = blockIdx.x * G + gr; item < n_items; item += gridDim.x * G) {
      const Item it = decode_item<TG>(item, nqb, H);
      float m_ref = -INFINITY, l_run = 0.f;
      for (int j = 0; j < nblk; ++j, ++g) {
        const int j0 = j * BKV;
        mbar_wait(&s_full[t], (uint32_t)(g & 1));
        tc_fence_after();
        uint32_t sv[2][32];
        tmem_ld_32x32(tmem_S, sv[0]);
        tmem_ld_32x32(tmem_S + 32, sv[1]);
        tmem_ld_wait();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[t]);    // S buffer may be overwritten by the next step's S

        if (pad != nullptr || j0 + BKV > R) {      // warp-uniform slow path: key masks
          uint32_t mask_lo = 0, mask_hi = 0;
          if (pad != nullptr) {
            const int ja = j0 + lane, jb = j0 + 32 + lane;
            mask_lo = __ballot_sync(0xffffffffu, ja < R && pad[(size_t)ja * C + it.c] != 0);
            mask_hi = __ballot_sync(0xffffffffu, jb < R && pad[(size_t)jb * C + it.c] != 0);
          }
          const int n_valid = R - j0;
#pragma unroll
          for (int e = 0; e < BKV; ++e) {
            const uint32_t mbits = (e < 32) ? mask_lo : mask_hi;
            float v = __uint_as_float(sv[e >> 5][e & 31]);
            if ((mbits >> (e & 31)) & 1u) v = neg_masked;
            if (e >= n_valid) v = -INFINITY;       // key row does not exist
            sv[e >> 5][e & 31] = __float_as_uint(v);
          }
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(sv[0][e]));
          mx1 = fmaxf(mx1, __uint_as_float(sv[0][e + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(sv[1][e]));
          mx3 = fmaxf(mx3, __uint_as_float(sv[1][e + 1]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * kLog2e;   // log2 domain; finite (>= 1 real key)
        // lazy rescale: keep the old reference unless the max grew by more than the threshold
        float factor = 1.f;
        const bool grow = mx > m_ref + kRescaleThreshold;   // true on the first block (m_ref = -inf)
        if (grow) {
          factor = ex2(m_ref - mx);                          // 0 on the first block
          m_ref = mx;
          l_run *= factor;
        }
        const bool rescale = (j > 0) && __any_sync(0xffffffffu, grow);
        // exponentials: (s * log2e - m) and the row sum run as packed fp32 pairs (FFMA2 / FADD2: the
        // kernel is bound by issue slots and the XU together); the 16-bit P words (keys 2w, 2w+1 ->
        // word w) overwrite the low half of each sv[hlf] as they are produced
        const uint64_t c_l2e = f32x2_pack(kLog2e, kLog2e), c_negm = f32x2_pack(-m_ref, -m_ref);
        uint64_t acc0 = f32x2_pack(0.f, 0.f), acc1 = acc0;
#pragma unroll
        for (int hlf = 0; hlf < 2; ++hlf) {
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const uint64_t x0 = f32x2_fma(f32x2_pack(__uint_as_float(sv[hlf][e]), __uint_as_float(sv[hlf][e + 1])), c_l2e, c_negm);
            const uint64_t x1 = f32x2_fma(f32x2_pack(__uint_as_float(sv[hlf][e + 2]), __uint_as_float(sv[hlf][e + 3])), c_l2e, c_negm);
            float a0, a1, a2, a3;
            f32x2_unpack(x0, a0, a1);
            a0 = ex2(a0); a1 = ex2(a1);
            if (kPolyPairs && ((e >> 2) & 1)) {       // compile-time pattern: one pair in four on the FMA pipe
              exp2_poly2(x1, a2, a3);
            } else {
              f32x2_unpack(x1, a2, a3);
              a2 = ex2(a2); a3 = ex2(a3);
            }
            acc0 = f32x2_add(acc0, f32x2_pack(a0, a1));
            acc1 = f32x2_add(acc1, f32x2_pack(a2, a3));
            sv[hlf][e >> 1] = kFp16 ? pack_f16(a0, a1) : pack_bf16(a0, a1);
            sv[hlf][(e >> 1) + 1] = kFp16 ? pack_f16(a2, a3) : pack_bf16(a2, a3);
          }
        }
        {
          float s0, s1, s2, s3;
          f32x2_unpack(acc0, s0, s1);
          f32x2_unpack(acc1, s2, s3);
          l_run += (s0 + s1) + (s2 + s3);
        }
        if (g > 0) {                                // PV(g-1) done: P buffer free, O_t stable
          mbar_wait(&pv_done[t], (uint32_t)((g - 1) & 1));
          tc_fence_after();
        }
        if (rescale) {                              // warp-uniform; rare once the max has settled
          uint32_t ov[32];
#pragma unroll 1
          for (int hlf = 0; hlf < 2; ++hlf) {
            tmem_ld_32x32(tmem_O + hlf * 32, ov);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; ++d) ov[d] = __float_as_uint(__uint_as_float(ov[d]) * factor);
            tmem_st_32x32(tmem_O + hlf * 32, ov);
          }
          tmem_st_wait();
        }
        if (j == 0) {                               // the previous item's O rows (staged in this warp's P rows) have
          if (lane == 0) bulk_wait_read0();         // been read by their TMA store
          __syncwarp();
        }
        // P row -> smem, K-major SWIZZLE_128B: 16-byte chunk ch of row r lives at chunk (ch ^ (r & 7)).
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          const uint32_t* w = &sv[ch >> 2][(ch & 3) * 4];
          *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
        fence_proxy_async_smem();   // generic-proxy P writes -> visible to the tensor-core (async) proxy
        tc_fence_before();          // our tcgen05.ld / st precede the MMA that follows the barrier
        __syncwarp();
        if (lane == 0) mbar_arrive(&p_full[t]);
      }
      // ---- item epilogue: O / l -> ctx -----------------------------------------------------------
      // The 16-bit rows are staged in this warp's 32 rows of the tile's P buffer (free once PV(last) is done,
      // same SWIZZLE_128B pattern) and leave through ONE TMA store per warp: asynchronous, so the warp goes
      // straight on to the next item.  (Direct st.global of 128 B per thread -- 32 lines per instruction --
      // cost 19 % of the kernel at R = 512 and 13 % at R = 256: measured by ablation.)
      mbar_wait(&pv_done[t], (uint32_t)((g - 1) & 1));
      tc_fence_after();
      const float inv = 1.f / l_run;
#pragma unroll 1
      for (int hlf = 0; hlf < 2; ++hlf) {
        uint32_t ov[32];
        tmem_ld_32x32(tmem_O + hlf * 32, ov);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; d += 8) {
          uint4 val;
          if (fp16)
            val = make_uint4(pack_f16(__uint_as_float(ov[d]) * inv, __uint_as_float(ov[d + 1]) * inv),
                             pack_f16(__uint_as_float(ov[d + 2]) * inv, __uint_as_float(ov[d + 3]) * inv),
                             pack_f16(__uint_as_float(ov[d + 4]) * inv, __uint_as_float(ov[d + 5]) * inv),
                             pack_f16(__uint_as_float(ov[d + 6]) * inv, __uint_as_float(ov[d + 7]) * inv));
          else
            val = make_uint4(pack_bf16(__uint_as_float(ov[d]) * inv, __uint_as_float(ov[d + 1]) * inv),
                             pack_bf16(__uint_as_float(ov[d + 2]) * inv, __uint_as_float(ov[d + 3]) * inv),
                             pack_bf16(__uint_as_float(ov[d + 4]) * inv, __uint_as_float(ov[d + 5]) * inv),
                             pack_bf16(__uint_as_float(ov[d + 6]) * inv, __uint_as_float(ov[d + 7]) * inv));
          const int ch = hlf * 4 + (d >> 3);
          *reinterpret_cast<uint4*>(prow + ((ch ^ (row & 7)) << 4)) = val;
        }
      }
      fence_proxy_async_smem();     // staged rows -> visible to the TMA (async proxy)
      __syncwarp();
      const int i_warp = it.i0 + tl * BQ + quad * 32;          // first query row of this warp; rows >= R are clipped
      if (lane == 0 && i_warp < R) {
        tma_store_3d(&tm_o, smem + K::OFF_P + t * P_BYTES + quad * 32 * 128, it.h * HD, it.c, i_warp);
        bulk_commit();
      }
      tc_fence_before();            // O loads precede the next item's first PV (ordered via p_full)
    }
  }

  if (warp >= 4 && lane == 0) bulk_wait_all0();   // the last items' TMA stores have left shared memory
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem_base, K::kTmemCols);
}

template <int NT, bool kFp16, int G = 1>
int launch_nt(const CUtensorMap& tq, const CUtensorMap& tkv, const CUtensorMap& to, int R, int C, int H, int col_major,
              const uint8_t* pad, cudaStream_t st) {
  using K = Cfg<NT, G>;
  static_assert(K::kSmem <= 227 * 1024, "column attention: shared memory budget");
  constexpr int TG = NT / G;
  static bool attr_set = false;
  if (!attr_set) {
    RNAMSM_CHECK_CUDA(cudaFuncSetAttribute(col_attn_ws_kernel<NT, kFp16, G>, cudaFuncAttributeMaxDynamicSharedMemorySize, K::kSmem));
    attr_set = true;
  }
  const long long n_items = (long long)C * H * ((R + TG * BQ - 1) / (TG * BQ));
  RNAMSM_REQUIRE(n_items < (1LL << 31), "col_attn_ws: too many work items");
  int sms = 0, dev = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  if (sms <= 0) sms = 148;
  const int grid = (int)std::min<long long>((n_items + G - 1) / G, sms);
  ProfScope prof(KC_COL_ATTN, st);
  RNAMSM_CHECK_CUDA(launch_pdl(col_attn_ws_kernel<NT, kFp16, G>, dim3(grid), dim3(K::kThreads), K::kSmem, st, tq, tkv, to, R, C, H,
                               col_major, (int)n_items, pad));
  count_launch();
  RNAMSM_CHECK_CUDA(cudaGetLastError());
  return 0;
}

}  // namespace

int launch_col_attn_ws_16(const void* qkv, int R, int C, int H, int fp16, int col_major, const uint8_t* pad, void* ctx,
                          cudaStream_t st) {
  RNAMSM_REQUIRE(R >= 1 && C >= 1 && H >= 1, "col_attn_ws: bad shape");
  const int ld = 3 * H * HD;
  CUtensorMap tq, tkv;
  // token-major q|k|v [R, C, 3D]: the rows of one column are C*3D elements apart (every 128 B row of a
  // box in a different 2 MiB page once the activation outgrows the TLB: measured ~27 cycles per row in
  // the TMA unit, the bound of this kernel).  column-major [C, R, 3D]: rows 3D elements apart.
  uint64_t dims[3] = {(uint64_t)ld, (uint64_t)(col_major ? R : C), (uint64_t)(col_major ? C : R)};
  uint64_t strides[2] = {(uint64_t)ld * 2, (uint64_t)(col_major ? R : C) * ld * 2};
  uint32_t box_q[3] = {HD, (uint32_t)(col_major ? BQ : 1), (uint32_t)(col_major ? 1 : BQ)};
  uint32_t box_kv[3] = {HD, (uint32_t)(col_major ? BKV : 1), (uint32_t)(col_major ? 1 : BKV)};
  const int in_dt = fp16 ? TMAP_F16 : TMAP_BF16;
  if (encode_tmap(&tq, in_dt, qkv, 3, dims, strides, box_q)) return 3;
  if (encode_tmap(&tkv, in_dt, qkv, 3, dims, strides, box_kv)) return 3;
  // ctx [R, C, D] token-major: one 32-row x 64-column box per softmax warp and item
  CUtensorMap to;
  uint64_t odims[3] = {(uint64_t)H * HD, (uint64_t)C, (uint64_t)R};
  uint64_t ostrides[2] = {(uint64_t)H * HD * 2, (uint64_t)C * H * HD * 2};
  uint32_t box_o[3] = {HD, 1, 32};
  if (encode_tmap(&to, in_dt, ctx, 3, odims, ostrides, box_o)) return 3;
  static int forced = -1;
  if (forced < 0) {
    const char* e = getenv("RNAMSM_COL_NT");
    forced = e ? atoi(e) : 0;
  }
  // EXPERIMENTAL, off by default, not yet run on hardware: RNAMSM_COL_GROUPS=2 -> four tiles as two item groups
  static int groups = -1;
  if (groups < 0) {
    const char* e = getenv("RNAMSM_COL_GROUPS");
    groups = (e && atoi(e) == 2) ? 2 : 1;
  }
  if (groups == 2)
    return fp16 ? launch_nt<4, true, 2>(tq, tkv, to, R, C, H, col_major, pad, st)
                : launch_nt<4, false, 2>(tq, tkv, to, R, C, H, col_major, pad, st);
  const int nt = forced == 2 || forced == 4 ? forced : (R > 2 * BQ ? 4 : 2);
  if (nt == 4)
    return fp16 ? launch_nt<4, true>(tq, tkv, to, R, C, H, col_major, pad, st)
                : launch_nt<4, false>(tq, tkv, to, R, C, H, col_major, pad, st);
  return fp16 ? launch_nt<2, true>(tq, tkv, to, R, C, H, col_major, pad, st)
              : launch_nt<2, false>(tq, tkv, to, R, C, H, col_major, pad, st);
}

}  // namespace rnamsm
