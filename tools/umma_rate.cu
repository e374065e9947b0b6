// Microbenchmark: sustained tcgen05.mma issue rate (cycles per MMA) for the operand forms the render kernel uses.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o umma_rate tools/umma_rate.cu && ./umma_rate
#include <cstdio>
#include <cstdint>
#include "../nerf-sos_b200/csrc/tc_ptx.cuh"
using namespace nsos::ptx;

// mode 0: TS (A in TMEM), mode 1: SS (A in smem); nb = number of distinct B tiles cycled (smem traffic pattern)
__global__ void __launch_bounds__(192, 1) k_rate(int mode, int N, int reps, int same_b, long long* out, const uint8_t* gsrc, int copy_mode) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar;
  __shared__ uint64_t cbar[4];
  __shared__ volatile int done;
  __shared__ uint32_t tptr;
  int warp = threadIdx.x >> 5;
  for (int i = threadIdx.x; i < (3 * 32768 + 16384) / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;  // fp16 1.0
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&cbar[i]), 1); done = 0; fence_mbar_init(); }
  if (warp == 4) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = tptr;
  if (warp == 4) {
    uint32_t idesc = make_idesc_f16(N);
    uint32_t a_smem = smem_u32(base + 3 * 32768);
    long long t0 = clock64();
    if (elect_one()) {
      for (int r = 0; r < reps; ++r) {
        uint32_t b = smem_u32(base + (same_b ? 0 : (r % 3) * 32768));
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          uint64_t bd = make_sw128_desc(b + ks * 32);
          if (mode == 0) umma_ts(tm, tm + 256 + (r % 4) * 32 + ks * 8, bd, idesc, 1);
          else umma_ss(tm, make_sw128_desc(a_smem + ks * 32), bd, idesc, 1);
        }
      }
      umma_commit(smem_u32(&bar));
    }
    __syncwarp();
    mbar_wait(smem_u32(&bar), 0, 1);
    long long t1 = clock64();
    if ((threadIdx.x & 31) == 0) { out[blockIdx.x] = t1 - t0; done = 1; }
  } else if (warp == 5 && copy_mode) {
    // concurrent weight-stream traffic: bulk copies of 32 KB into a scratch region (not read by the MMAs), 2 in flight
    uint8_t* scratch = base + 3 * 32768 + 16384;
    uint32_t n = 0;
    if (elect_one()) {
      while (!done) {
        uint32_t slot = n & 1;
        if (n >= 2) mbar_wait(smem_u32(&cbar[slot]), ((n >> 1) - 1) & 1, 7);
        mbar_arrive_expect_tx(smem_u32(&cbar[slot]), 32768);
        bulk_g2s(smem_u32(scratch + slot * 32768), gsrc + (size_t)((n * 148 + blockIdx.x) % 64) * 32768, 32768, smem_u32(&cbar[slot]));
        ++n;
        if (copy_mode == 2) { long long t = clock64(); while (clock64() - t < 900) {} }   // throttle to ~32 KB / 1000 cycles
      }
      if (n >= 1) mbar_wait(smem_u32(&cbar[(n - 1) & 1]), ((n - 1) >> 1) & 1, 8);
      if (n >= 2) mbar_wait(smem_u32(&cbar[(n - 2) & 1]), ((n - 2) >> 1) & 1, 9);
    }
  }
  tc_fence_before(); __syncthreads();
  if (warp == 4) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  size_t smem = 3 * 32768 + 16384 + 2 * 32768 + 1024;
  uint8_t* gsrc; cudaMalloc(&gsrc, 64 * 32768); cudaMemset(gsrc, 0, 64 * 32768);
  cudaFuncSetAttribute(k_rate, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int reps = 2000;
  for (int grid : {1, 148})
    for (int mode : {0, 1})
      for (int N : {256, 128})
        for (int same_b : {0, 1, 2}) {
          int copy_mode = same_b;   // 0: no copies, 1: copies at full speed, 2: throttled copies
          k_rate<<<grid, 192, smem>>>(mode, N, reps, 0, d, gsrc, copy_mode);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[148]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("grid=%3d %s N=%3d copies=%d: %.1f cycles/MMA (M=128,K=16)  -> %.0f MAC/clk/SM  [%s]\n", grid, mode ? "SS" : "TS", N, same_b,
                 (double)mx / (reps * 4), 128.0 * N * 16 / ((double)mx / (reps * 4)), cudaGetErrorString(e));
        }
  return 0;
}
