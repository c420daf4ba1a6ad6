// Minimal stand-alone reproduction for the compute-sanitizer racecheck report on tcgen05.alloc.cta_group::2
// (profiles/r01_sanitizer.md): a cluster of two CTAs does NOTHING but the paired TMEM allocation, the canonical
// fence / barrier sequence, one read of the returned base address, and the paired deallocation.  No other shared-
// memory access exists in the kernel, so any hazard racecheck reports here is between the two halves of the paired
// allocation instruction itself (both CTAs issue it; the hardware writes the base into each CTA's shared memory).
//   nvcc -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -o tmem_alloc2_racecheck tmem_alloc2_racecheck.cu
//   compute-sanitizer --tool racecheck ./tmem_alloc2_racecheck
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

template <int kGroup>
__global__ void __cluster_dims__(kGroup, 1, 1) __launch_bounds__(128, 1) alloc_kernel(uint32_t* out) {
  __shared__ uint32_t tmem_ptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) {
    if (kGroup == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(64u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&tmem_ptr)), "r"(64u) : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
  __syncthreads();
  if (kGroup == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
  const uint32_t base = tmem_ptr;
  if (threadIdx.x == 0) out[blockIdx.x] = base;
  __syncthreads();
  if (kGroup == 2) {
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
  }
  if (warp == 0) {
    if (kGroup == 2) asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64u) : "memory");
    else asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(base), "r"(64u) : "memory");
  }
}

int main() {
  uint32_t* out;
  cudaMalloc(&out, 8 * sizeof(uint32_t));
  alloc_kernel<1><<<2, 128>>>(out);
  cudaError_t e1 = cudaDeviceSynchronize();
  alloc_kernel<2><<<2, 128>>>(out);
  cudaError_t e2 = cudaDeviceSynchronize();
  uint32_t h[2];
  cudaMemcpy(h, out, sizeof(h), cudaMemcpyDeviceToHost);
  printf("cta_group::1: %s   cta_group::2: %s   tmem base CTA0 0x%x CTA1 0x%x\n", cudaGetErrorString(e1), cudaGetErrorString(e2), h[0], h[1]);
  return (e1 != cudaSuccess || e2 != cudaSuccess) ? 1 : 0;
}
