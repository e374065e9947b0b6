// Microbenchmark: depth of the tcgen05.mma issue queue and the cost of commit / mbarrier wait between MMA groups.
#include <cstdio>
#include <cstdint>
#include "../nerf-sos_b200/csrc/tc_ptx.cuh"
using namespace nsos::ptx;

__global__ void __launch_bounds__(192, 1) k_queue(int variant, int group, int reps, long long* out, long long* stamps) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar, gbar[4];
  __shared__ uint32_t tptr;
  int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (3 * 32768) / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&gbar[i]), 1); fence_mbar_init(); mbar_arrive(smem_u32(&gbar[3])); }
  if (warp == 4) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = tptr;
  if (warp == 4) {
    uint32_t idesc = make_idesc_f16(256);
    long long t0 = clock64();
    if (variant == 0) {          // stamps after each of the first 32 MMA issues
      if (elect_one()) {
        for (int r = 0; r < 32; ++r) {
          umma_ts(tm, tm + 256 + (r % 16) * 8, make_sw128_desc(smem_u32(base) + (r % 4) * 32), idesc, 1);
          stamps[r] = clock64() - t0;
        }
        umma_commit(smem_u32(&bar));
      }
    } else {                     // groups of `group` MMAs, each followed by a commit to a per-group barrier; variant 2 also waits on an
                                 // already-completed barrier (2 groups back) between groups, like the weight-ring consumer does
      uint32_t g = 0;
      for (int r = 0; r < reps; ++r) {
        if (variant == 2) mbar_wait(smem_u32(&gbar[3]), 0, 5);   // a barrier whose phase 0 completed at start: returns at once
        if (elect_one()) {
          for (int k = 0; k < group; ++k)
            umma_ts(tm, tm + 256 + (k % 16) * 8, make_sw128_desc(smem_u32(base) + (k % 4) * 32), idesc, 1);
          umma_commit(smem_u32(&gbar[g & 1]));
        }
        __syncwarp();
        ++g;
      }
      if (elect_one()) umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, 1);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) out[0] = t1 - t0;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 512);
}

int main() {
  long long *d, *st; cudaMalloc(&d, 8); cudaMalloc(&st, 32 * 8);
  size_t smem = 3 * 32768 + 2048;
  cudaFuncSetAttribute(k_queue, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  k_queue<<<1, 192, smem>>>(0, 0, 0, d, st);
  cudaDeviceSynchronize();
  long long h[32], tot; cudaMemcpy(h, st, sizeof(h), cudaMemcpyDeviceToHost); cudaMemcpy(&tot, d, 8, cudaMemcpyDeviceToHost);
  printf("issue stamps (cycles since start):"); for (int i = 0; i < 32; ++i) printf(" %lld", h[i]); printf("\n total %lld for 32 MMAs\n", tot);
  for (int variant : {1, 2})
    for (int group : {2, 4, 8, 12, 16, 24}) {
      int reps = 1920 / group;
      k_queue<<<1, 192, smem>>>(variant, group, reps, d, st);
      cudaError_t e = cudaDeviceSynchronize();
      cudaMemcpy(&tot, d, 8, cudaMemcpyDeviceToHost);
      printf("variant %d group=%2d: %.1f cycles per MMA, %.0f per group (ideal %d) [%s]\n", variant, group, (double)tot / (reps * group),
             (double)tot / reps, group * 137, cudaGetErrorString(e));
    }
  return 0;
}
