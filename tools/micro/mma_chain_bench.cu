// Microbenchmark: how fast ONE thread can push tcgen05.mma (M128 N64 K16, kind::f16, cta_group::1, operands in shared
// memory) through the tensor pipe, depending on how the instructions depend on each other, and what it costs to give
// one attention tile two issuing threads (one for S = Q K^T, one for O += P V) instead of one.
// The column attention issues, per tile and 64-key step, one 4-instruction S chain (overwrite + 3 accumulates) and one
// 4-instruction PV chain (accumulates) from a single thread.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I rna-msm_b200/csrc -I include \
//        -o tools/micro/mma_chain_bench tools/micro/mma_chain_bench.cu -lcuda
#include <cstdio>
#include <cstdlib>
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "common.cuh"

using namespace rnamsm;

enum Pattern {
  ACC_CHAIN = 0,      // every instruction accumulates into one accumulator
  OVW_CHAIN = 1,      // every instruction overwrites one accumulator
  OVW_ROTATE = 2,     // overwrites, 4 accumulators in rotation (independent instructions)
  S_THEN_PV = 3,      // per iteration: S chain (ovw + 3 acc -> S_t) then PV chain (4 acc -> O_t): the production pattern
  S_PV_INTERLEAVED = 4,  // the same 8 instructions, alternating S_t / O_t
  S_ONLY = 5,         // per iteration: S chain only
  PV_ONLY = 6,        // per iteration: PV chain only
  S_THEN_PV_N128 = 7, // 128-key step: S chain of 4 with N = 128, PV chain of 8 (N = 64)
};

struct Params {
  int n_threads;        // issuing threads
  int lanes_per_warp;   // 1: thread i = lane 0 of warp i; 2: threads 2w, 2w+1 = lanes 0, 1 of warp w
  int pattern[8];
  int iters;
  int load_warps;       // warps 8.. : MUFU + FFMA load on every sub-partition (0, 8 or 16 warps)
};

__global__ void __launch_bounds__(768, 1) k(Params p, long long* out) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint8_t* sA = smem;                    // 128 x 64, SW128
  uint8_t* sB = smem + 16384;            // 128 x 64
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 32768);
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);
  int* done = reinterpret_cast<int*>(bars + 17);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) {
    *done = 0;
    for (int b = 0; b < 16; ++b) mbar_init(&bars[b], b < 8 ? 1 : 1000000);
    fence_mbar_init();
  }
  if (warp == 0) {
    tmem_alloc(tmem_ptr, 512);
    tmem_relinquish();
  }
  fence_proxy_async_smem();
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  int tid = -1;
  if (warp < 8) {
    if (p.lanes_per_warp == 1 && lane == 0) tid = warp;
    if (p.lanes_per_warp == 2 && lane < 2) tid = warp * 2 + lane;
    if (tid >= p.n_threads) tid = -1;
  }
  if (tid >= 0) {
    const int pat = p.pattern[tid];
    const int tile = tid & 3;
    const uint32_t d_s = tmem_base + tile * 64, d_o = tmem_base + 256 + tile * 64;
    const uint32_t idesc = make_idesc_16(128, 64, 1, 0, 0), idesc128 = make_idesc_16(128, 128, 1, 0, 0);
    const uint32_t a = smem_u32(sA), b = smem_u32(sB);
    auto mma = [&](uint32_t d, int kk, uint32_t acc, uint32_t id) {
      umma_16(d, make_smem_desc_sw128(a + (kk & 3) * 32, 16, 1024), make_smem_desc_sw128(b + (kk & 3) * 32, 16, 1024), id, acc);
    };
    long long n_inst = 0;
    const long long t0 = clock64();
    for (int it = 0; it < p.iters; ++it) {
      if (pat == ACC_CHAIN) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) mma(d_o, kk, 1u, idesc);
        n_inst += 8;
      } else if (pat == OVW_CHAIN) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) mma(d_o, kk, 0u, idesc);
        n_inst += 8;
      } else if (pat == OVW_ROTATE) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) mma(tmem_base + (kk & 3) * 64 + (tile >> 1) * 256, kk, 0u, idesc);
        n_inst += 8;
      } else if (pat == S_THEN_PV) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mma(d_s, kk, (uint32_t)(kk != 0), idesc);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mma(d_o, kk, 1u, idesc);
        n_inst += 8;
      } else if (pat == S_PV_INTERLEAVED) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          mma(d_s, kk, (uint32_t)(kk != 0), idesc);
          mma(d_o, kk, 1u, idesc);
        }
        n_inst += 8;
      } else if (pat == S_ONLY) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mma(d_s, kk, (uint32_t)(kk != 0), idesc);
        n_inst += 4;
      } else if (pat == PV_ONLY) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mma(d_o, kk, 1u, idesc);
        n_inst += 4;
      } else if (pat == S_THEN_PV_N128) {
        const uint32_t d_s2 = tmem_base + (tile & 1) * 128, d_o2 = tmem_base + 256 + (tile & 1) * 64;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) mma(d_s2, kk, (uint32_t)(kk != 0), idesc128);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk) mma(d_o2, kk, 1u, idesc);
        n_inst += 12;
      }
      if ((it & 7) == 7) umma_commit(&bars[8 + tid]);
    }
    umma_commit(&bars[tid]);
    mbar_wait(&bars[tid], 0);
    const long long t1 = clock64();
    out[(blockIdx.x * 16 + tid) * 2] = t1 - t0;
    out[(blockIdx.x * 16 + tid) * 2 + 1] = n_inst;
    atomicAdd(done, 1);
  } else if (warp >= 8 && warp - 8 < p.load_warps) {
    float x[16];
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = 0.5f + i * 0.01f + lane * 1e-3f;
    while (*reinterpret_cast<volatile int*>(done) < p.n_threads) {
#pragma unroll
      for (int r = 0; r < 4; ++r)
#pragma unroll
        for (int i = 0; i < 16; ++i) {
          asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(x[i]));
          x[i] = fmaf(x[i], 0.999f, 0.001f);
          x[i] = fmaf(x[i], 1.001f, -0.001f);
        }
    }
    float s = 0;
#pragma unroll
    for (int i = 0; i < 16; ++i) s += x[i];
    if (s == 123.456f) out[0] = 1;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, 512);
}

static const char* pat_name[] = {"acc-chain", "ovw-chain", "ovw-rotate", "S;PV", "S/PV interleaved", "S only", "PV only", "S128;PV (128 keys)"};

static void run(const char* what, Params p, long long* d_out) {
  const int smem = 32768 + 256 + 1024;
  cudaMemset(d_out, 0, 148 * 16 * 2 * sizeof(long long));
  k<<<148, 768, smem>>>(p, d_out);
  cudaError_t e = cudaDeviceSynchronize();
  if (e != cudaSuccess) {
    printf("CUDA error: %s\n", cudaGetErrorString(e));
    exit(1);
  }
  static long long h[148 * 16 * 2];
  cudaMemcpy(h, d_out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("%-46s threads %d (%d per warp) load warps %2d |", what, p.n_threads, p.lanes_per_warp, p.load_warps);
  for (int t = 0; t < p.n_threads; ++t) {
    double cyc = 0, n = 0;
    for (int b = 0; b < 148; ++b) {
      cyc += (double)h[(b * 16 + t) * 2];
      n += (double)h[(b * 16 + t) * 2 + 1];
    }
    if (t < 2 || t == 4) printf(" t%d %-18s %.0f cyc/iter %.0f cyc/instr |", t, pat_name[p.pattern[t]], cyc / 148 / p.iters, cyc / n);
  }
  printf("\n");
}

int main() {
  long long* d_out;
  cudaMalloc(&d_out, 148 * 16 * 2 * sizeof(long long));
  auto P = [&](int n, int lpw, int pat_even, int pat_odd_or_hi, bool split_hi, int load) {
    Params p;
    p.n_threads = n; p.lanes_per_warp = lpw; p.iters = 3000; p.load_warps = load;
    for (int t = 0; t < 8; ++t) {
      if (lpw == 2) p.pattern[t] = (t & 1) ? pat_odd_or_hi : pat_even;   // lanes 0 / 1 of warp t / 2... tile = t & 3
      else p.pattern[t] = (split_hi && t >= 4) ? pat_odd_or_hi : pat_even;
    }
    return p;
  };
  for (int load : {0, 16}) {
    run("1 thread, accumulate chain", P(1, 1, ACC_CHAIN, 0, false, load), d_out);
    run("1 thread, overwrite chain", P(1, 1, OVW_CHAIN, 0, false, load), d_out);
    run("1 thread, overwrites into 4 accumulators", P(1, 1, OVW_ROTATE, 0, false, load), d_out);
    run("1 thread, S chain then PV chain", P(1, 1, S_THEN_PV, 0, false, load), d_out);
    run("1 thread, S / PV interleaved", P(1, 1, S_PV_INTERLEAVED, 0, false, load), d_out);
    run("1 thread, S only", P(1, 1, S_ONLY, 0, false, load), d_out);
    run("1 thread, PV only", P(1, 1, PV_ONLY, 0, false, load), d_out);
    run("4 tiles, 1 thread each: S;PV", P(4, 1, S_THEN_PV, 0, false, load), d_out);
    run("4 tiles, 1 thread each: interleaved", P(4, 1, S_PV_INTERLEAVED, 0, false, load), d_out);
    run("4 tiles, 2 threads each in 8 warps: S | PV", P(8, 1, S_ONLY, PV_ONLY, true, load), d_out);
    {
      // lanes 0 / 1 of warps 0..3: tile = tid & 3 would pair (0,1) on tiles 0,1 -- give both lanes of a warp one tile instead
      Params p = P(8, 2, S_ONLY, PV_ONLY, false, load);
      run("4 tiles, 2 threads each as lanes 0/1 of 4 warps", p, d_out);
    }
    run("2 tiles, 128-key steps: S128;PV", P(2, 1, S_THEN_PV_N128, 0, false, load), d_out);
    run("4 threads, 128-key steps (2 S-tiles shared)", P(4, 1, S_THEN_PV_N128, 0, false, load), d_out);
  }
  cudaFree(d_out);
  return 0;
}
