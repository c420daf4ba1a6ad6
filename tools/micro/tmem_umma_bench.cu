// Microbenchmark: per-SM throughput of the resources the column-attention kernel shares, alone and together:
//   LDTM  (tcgen05.ld 32x32b.x32: 4 KiB per warp instruction), STS.128, MUFU.EX2, and tcgen05.mma M128 x N x K16 with
//   the A operand in shared memory (SS) or in tensor memory (TS).
// One CTA per SM: 16 worker warps in 4 groups (group g = warps 4g..4g+3, one warp per SM sub-partition) and one
// MMA-issuing warp.  Every group runs one job; the table in main() lists the combinations.  Cycles are clock64()
// deltas of each warp around its own loop (mean over CTAs), so overlapping jobs show each other's slow-down directly.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rna-msm_b200/csrc -o tools/micro/tmem_umma_bench \
//        tools/micro/tmem_umma_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

using namespace rnamsm;

enum Job { IDLE = 0, LDTM = 1, STS = 2, MUFU = 3, STTM = 4, SOFTMAXISH = 5, SMX_MATH = 6, SMX_SYNC = 7 };
enum Mma { NONE = 0, SS64 = 1, TS64 = 2, SS128 = 3, TS128 = 4, SS256 = 5 };

struct Params {
  int job[4];
  int mma;
  int iters;       // worker iterations
  int mma_iters;   // MMA iterations (4 instructions each)
  int mma_threads; // 1..4 issuing threads (lane 0 of warps 16..19), each with its own accumulator columns
  int alt_d;       // 1: consecutive iterations alternate between two accumulators
  int commit_each; // 1: tcgen05.commit after every 4 instructions and wait for it (latency of a 4-instruction chain)
};

__global__ void __launch_bounds__(640, 1) bench_kernel(Params p, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                    // 128 x 64 16-bit, SW128: 16 KiB
  uint8_t* sB = smem + 16384;            // 256 x 64 16-bit: 32 KiB
  uint8_t* sP = smem + 49152;            // 4 groups x 16 KiB store targets
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 49152 + 65536);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 48);
  int* done = reinterpret_cast<int*>(bars + 49);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (49152 + 65536) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    *done = 0;
    for (int b = 0; b < 48; ++b) mbar_init(&bars[b], (b == 1 || b >= 16) ? 1000000 : 1);
    fence_mbar_init();
  }
  if (warp == 16) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  long long t0 = 0, t1 = 0;

  if (warp >= 16) {
    if (lane == 0 && p.mma != NONE && warp - 16 < p.mma_threads) {
      const int mt = warp - 16;
      const int N = (p.mma == SS64 || p.mma == TS64) ? 64 : (p.mma == SS256 ? 256 : 128);
      const bool ts = p.mma == TS64 || p.mma == TS128;
      const uint32_t idesc = make_idesc_16(128, N, 1, 0, 0);
      const uint32_t a = smem_u32(sA), b = smem_u32(sB);
      t0 = clock64();
      int n_active = 0;
      for (int g = 0; g < 4; ++g) n_active += p.job[g] != IDLE ? 4 : 0;
      int it = 0;
      // keep issuing until every worker warp has finished, so that the workers are timed under MMA load throughout
      for (; it < p.mma_iters || *reinterpret_cast<volatile int*>(done) < n_active; ++it) {
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const bool odd = p.alt_d && (it & 1);
          const uint32_t d = tmem_base + (N == 256 ? (odd ? 0 : 256) : N == 128 ? (odd ? 384 : 256) : 256 + mt * 64 + (odd ? 64 : 0));
          if (ts)
            umma_16_ts(d, tmem_base + k * 8, make_smem_desc_sw128(b + k * 32, 16, 1024), idesc, (uint32_t)(k != 0));
          else
            umma_16(d, make_smem_desc_sw128(a + k * 32, 16, 1024), make_smem_desc_sw128(b + k * 32, 16, 1024),
                    idesc, (uint32_t)(k != 0));
        }
        if (p.commit_each) {
          umma_commit(&bars[8 + mt]);
          mbar_wait(&bars[8 + mt], (uint32_t)(it & 1));
        } else if ((it & 15) == 15) {
          umma_commit(&bars[1]);
        }
      }
      umma_commit(&bars[4 + mt]);
      mbar_wait(&bars[4 + mt], 0);
      t1 = clock64();
      out[(blockIdx.x * 20 + warp) * 2 + 1] = it;
    }
  } else {
    const int g = warp >> 2, quad = warp & 3;
    const int job = p.job[g];
    const uint32_t lane_off = (uint32_t)(quad * 32) << 16;
    const uint32_t taddr = tmem_base + lane_off + 64 + (g & 1) * 64;
    const int row = quad * 32 + lane;
    uint8_t* prow = sP + g * 16384 + row * 128;
    uint32_t v[2][32];
#pragma unroll
    for (int i = 0; i < 32; ++i) { v[0][i] = 0x3c003c00u + i; v[1][i] = 0x3c003c00u + lane; }
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = 0.5f + i * 0.01f + lane * 1e-3f;
    __syncwarp();
    t0 = clock64();
    if (job == LDTM) {
      for (int it = 0; it < p.iters; ++it) {
        tmem_ld_32x32(taddr, v[0]);
        tmem_ld_32x32(taddr + 32, v[1]);
        tmem_ld_wait();
      }
    } else if (job == STTM) {
      for (int it = 0; it < p.iters; ++it) {
        tmem_st_32x32(taddr, v[0]);
        tmem_st_wait();
      }
    } else if (job == STS) {
      for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(prow + ((ch ^ (row & 7)) << 4))),
                       "r"(v[0][ch * 4]), "r"(v[0][ch * 4 + 1]), "r"(v[0][ch * 4 + 2]), "r"(v[0][ch * 4 + 3])
                       : "memory");
        }
      }
    } else if (job == MUFU) {
      for (int it = 0; it < p.iters; ++it) {
#pragma unroll
        for (int r = 0; r < 4; ++r)
#pragma unroll
          for (int i = 0; i < 16; ++i) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
      }
    } else if (job == SOFTMAXISH) {
      // the step body of the column attention without any barrier: LDTM 64 columns, 64 ex2, 8 STS.128, proxy fence
      for (int it = 0; it < p.iters; ++it) {
        tmem_ld_32x32(taddr, v[0]);
        tmem_ld_32x32(taddr + 32, v[1]);
        tmem_ld_wait();
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float a0 = __uint_as_float(v[h][i]), a1 = __uint_as_float(v[h][i + 1]);
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
            asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
            v[h][i >> 1] = pack_f16(a0, a1);
          }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(prow + ((ch ^ (row & 7)) << 4))),
                       "r"(v[ch >> 2][(ch & 3) * 4]), "r"(v[ch >> 2][(ch & 3) * 4 + 1]), "r"(v[ch >> 2][(ch & 3) * 4 + 2]),
                       "r"(v[ch >> 2][(ch & 3) * 4 + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
      }
    } else if (job == SMX_MATH || job == SMX_SYNC) {
      // the full arithmetic of one column-attention step (row max with lazy rescale, packed FFMA2 / FADD2, 64 ex2,
      // 32 packs, 8 STS.128, proxy fence); SMX_SYNC adds the step's mbarrier traffic on barriers that never block
      // (3 try_waits on a completed phase, 2 arrives by lane 0, 2 __syncwarp, the tcgen05 fences)
      float m_ref = -INFINITY, l_run = 0.f;
      uint64_t* dummy = bars + 16 + (warp & 15);
      for (int it = 0; it < p.iters; ++it) {
        if (job == SMX_SYNC) {
          mbar_wait_quiet(dummy, 1);
          tc_fence_after();
        }
        tmem_ld_32x32(taddr, v[0]);
        tmem_ld_32x32(taddr + 32, v[1]);
        tmem_ld_wait();
        if (job == SMX_SYNC) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dummy + 16);
        }
        float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
        for (int e = 0; e < 32; e += 2) {
          mx0 = fmaxf(mx0, __uint_as_float(v[0][e]));
          mx1 = fmaxf(mx1, __uint_as_float(v[0][e + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(v[1][e]));
          mx3 = fmaxf(mx3, __uint_as_float(v[1][e + 1]));
        }
        const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * 1.4426950408889634f;
        if (mx > m_ref + 8.f) {
          float f;
          float d = m_ref - mx;
          asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(f) : "f"(d));
          m_ref = mx;
          l_run *= f;
        }
        const uint64_t c_l2e = f32x2_pack(1.4426950408889634f, 1.4426950408889634f), c_negm = f32x2_pack(-m_ref, -m_ref);
        uint64_t acc0 = f32x2_pack(0.f, 0.f), acc1 = acc0;
#pragma unroll
        for (int h = 0; h < 2; ++h)
#pragma unroll
          for (int e = 0; e < 32; e += 4) {
            const uint64_t x0 = f32x2_fma(f32x2_pack(__uint_as_float(v[h][e]), __uint_as_float(v[h][e + 1])), c_l2e, c_negm);
            const uint64_t x1 = f32x2_fma(f32x2_pack(__uint_as_float(v[h][e + 2]), __uint_as_float(v[h][e + 3])), c_l2e, c_negm);
            float a0, a1, a2, a3;
            f32x2_unpack(x0, a0, a1);
            f32x2_unpack(x1, a2, a3);
            asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(a0));
            asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(a1));
            asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(a2));
            asm("ex2.approx.ftz.f32 %0, %0;" : "+f"(a3));
            acc0 = f32x2_add(acc0, f32x2_pack(a0, a1));
            acc1 = f32x2_add(acc1, f32x2_pack(a2, a3));
            v[h][e >> 1] = pack_f16(a0, a1);
            v[h][(e >> 1) + 1] = pack_f16(a2, a3);
          }
        {
          float s0, s1, s2, s3;
          f32x2_unpack(acc0, s0, s1);
          f32x2_unpack(acc1, s2, s3);
          l_run += (s0 + s1) + (s2 + s3);
        }
        if (job == SMX_SYNC) {
          mbar_wait_quiet(dummy, 1);
          tc_fence_after();
        }
#pragma unroll
        for (int ch = 0; ch < 8; ++ch) {
          asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(smem_u32(prow + ((ch ^ (row & 7)) << 4))),
                       "r"(v[ch >> 2][(ch & 3) * 4]), "r"(v[ch >> 2][(ch & 3) * 4 + 1]), "r"(v[ch >> 2][(ch & 3) * 4 + 2]),
                       "r"(v[ch >> 2][(ch & 3) * 4 + 3])
                       : "memory");
        }
        fence_proxy_async_smem();
        if (job == SMX_SYNC) {
          tc_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(dummy + 16);
        }
      }
      x[0] += l_run;
    }
    t1 = clock64();
    if (lane == 0 && job != IDLE) atomicAdd(done, 1);
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456f) out[0] = v[0][3] + v[1][5];
  }
  if (lane == 0) out[(blockIdx.x * 20 + warp) * 2] = t1 - t0;
  tc_fence_before();
  __syncthreads();
  if (warp == 16) tmem_dealloc(tmem_base, 512);
}

static const char* job_name[] = {"-", "LDTM", "STS", "MUFU", "STTM", "SMX", "SMXm", "SMXs"};
static const char* mma_name[] = {"-", "SS N64", "TS N64", "SS N128", "TS N128", "SS N256"};

static void run(Params p, long long* d_out) {
  const int smem = 49152 + 65536 + 512 + 1024;
  static bool set = false;
  if (!set) {
    cudaFuncSetAttribute(bench_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    set = true;
  }
  cudaMemset(d_out, 0, 148 * 20 * 2 * sizeof(long long));
  bench_kernel<<<148, 640, smem>>>(p, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(e));
    exit(1);
  }
  static long long h[148 * 20 * 2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("mma %-8s jobs %-4s %-4s %-4s %-4s |", mma_name[p.mma], job_name[p.job[0]], job_name[p.job[1]], job_name[p.job[2]],
         job_name[p.job[3]]);
  if (p.mma != NONE) {
    double s = 0;
    for (int b = 0; b < 148; ++b)
      for (int t = 0; t < p.mma_threads; ++t) s += (double)h[(b * 20 + 16 + t) * 2] / (4.0 * (double)h[(b * 20 + 16 + t) * 2 + 1]);
    printf(" mma x%d%s%s %.1f cyc/instr/thread |", p.mma_threads, p.alt_d ? " altD" : "", p.commit_each ? " commit+wait each 4" : "",
           s / 148 / p.mma_threads);
  }
  for (int g = 0; g < 4; ++g) {
    if (p.job[g] == IDLE) continue;
    double s = 0;
    for (int b = 0; b < 148; ++b)
      for (int q = 0; q < 4; ++q) s += (double)h[(b * 20 + g * 4 + q) * 2];
    printf(" g%d %s %.1f cyc/iter", g, job_name[p.job[g]], s / 148 / 4 / p.iters);
  }
  printf("\n");
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * 20 * 2 * sizeof(long long));
  const int IT = 2000, MIT = 4000;
  auto P = [&](int mma, int j0, int j1, int j2, int j3) {
    Params p;
    p.job[0] = j0; p.job[1] = j1; p.job[2] = j2; p.job[3] = j3;
    p.mma = mma; p.iters = IT; p.mma_iters = MIT; p.mma_threads = 1; p.alt_d = 0; p.commit_each = 0;
    return p;
  };
  printf("# iter = LDTM: 2 x (32 lanes x 32 col) = 8 KiB per warp; STS: 8 x STS.128 = 4 KiB per warp; MUFU: 64 ex2 per thread;\n"
         "# STTM: 4 KiB per warp; SMX: LDTM 8 KiB + 64 ex2 + 32 cvt + 8 STS.128 + proxy fence.  One warp per sub-partition per group.\n");
  run(P(NONE, LDTM, IDLE, IDLE, IDLE), d_out);
  run(P(NONE, LDTM, LDTM, IDLE, IDLE), d_out);
  run(P(NONE, LDTM, LDTM, LDTM, LDTM), d_out);
  run(P(NONE, STTM, IDLE, IDLE, IDLE), d_out);
  run(P(NONE, STTM, STTM, STTM, STTM), d_out);
  run(P(NONE, STS, IDLE, IDLE, IDLE), d_out);
  run(P(NONE, STS, STS, STS, STS), d_out);
  run(P(NONE, MUFU, IDLE, IDLE, IDLE), d_out);
  run(P(NONE, MUFU, MUFU, MUFU, MUFU), d_out);
  run(P(SS64, IDLE, IDLE, IDLE, IDLE), d_out);
  run(P(TS64, IDLE, IDLE, IDLE, IDLE), d_out);
  run(P(SS128, IDLE, IDLE, IDLE, IDLE), d_out);
  run(P(TS128, IDLE, IDLE, IDLE, IDLE), d_out);
  run(P(SS256, IDLE, IDLE, IDLE, IDLE), d_out);
  run(P(SS64, LDTM, LDTM, LDTM, LDTM), d_out);
  run(P(TS64, LDTM, LDTM, LDTM, LDTM), d_out);
  run(P(SS64, STS, STS, STS, STS), d_out);
  run(P(TS64, STS, STS, STS, STS), d_out);
  run(P(SS64, MUFU, MUFU, MUFU, MUFU), d_out);
  run(P(SS64, LDTM, MUFU, MUFU, STS), d_out);
  run(P(NONE, SOFTMAXISH, IDLE, IDLE, IDLE), d_out);
  run(P(NONE, SOFTMAXISH, SOFTMAXISH, IDLE, IDLE), d_out);
  run(P(NONE, SOFTMAXISH, SOFTMAXISH, SOFTMAXISH, SOFTMAXISH), d_out);
  run(P(SS64, SOFTMAXISH, SOFTMAXISH, SOFTMAXISH, SOFTMAXISH), d_out);
  run(P(TS64, SOFTMAXISH, SOFTMAXISH, SOFTMAXISH, SOFTMAXISH), d_out);
  {
    Params p = P(SS64, IDLE, IDLE, IDLE, IDLE);
    p.alt_d = 1; run(p, d_out);
    p.alt_d = 0; p.mma_threads = 2; run(p, d_out);
    p.mma_threads = 4; run(p, d_out);
    p.mma = TS64; run(p, d_out);
    p.mma = SS64; p.mma_threads = 1; p.commit_each = 1; run(p, d_out);
    p.mma_threads = 4; run(p, d_out);
    p = P(SS64, SOFTMAXISH, SOFTMAXISH, SOFTMAXISH, SOFTMAXISH);
    p.mma_threads = 4; run(p, d_out);
    p.mma = TS64; run(p, d_out);
    p = P(NONE, SMX_MATH, IDLE, IDLE, IDLE); run(p, d_out);
    p = P(NONE, SMX_MATH, SMX_MATH, SMX_MATH, SMX_MATH); run(p, d_out);
    p = P(NONE, SMX_SYNC, IDLE, IDLE, IDLE); run(p, d_out);
    p = P(NONE, SMX_SYNC, SMX_SYNC, SMX_SYNC, SMX_SYNC); run(p, d_out);
    p = P(SS64, SMX_SYNC, SMX_SYNC, SMX_SYNC, SMX_SYNC); p.mma_threads = 4; run(p, d_out);
  }
  cudaFree(d_out);
  return 0;
}
