// Microbenchmark: sustained cp.async.bulk (TMA engine) L2 -> shared-memory ingest rate per SM with all SMs streaming
// the same 2.5 MB buffer (the weight image of kernel A), unicast vs cluster multicast.
#include <cstdio>
#include <cstdint>
#include "../nerf-sos_b200/csrc/tc_ptx.cuh"
using namespace nsos::ptx;

__global__ void __launch_bounds__(128, 1) k_bulk(const uint8_t* src, int nchunks_total, int reps, int chunk_bytes, int depth, long long* out, int nw) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = (uint8_t*)(((uintptr_t)raw + 1023) & ~(uintptr_t)1023);
  __shared__ uint64_t bar[32];
  const uint32_t csize = cluster_nctarank(), crank = cluster_ctarank();
  if (threadIdx.x == 0) { for (int i = 0; i < 32; ++i) mbar_init(smem_u32(&bar[i]), 1); fence_mbar_init(); }
  __syncthreads();
  if (csize > 1) cluster_sync();
  const int w = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0 && w < nw) {
    uint64_t* bar_w = bar + w * 8;
    base += (size_t)w * depth * chunk_bytes;
    long long t0 = clock64();
    const uint32_t share = chunk_bytes / csize;
    for (int n = 0; n < reps; ++n) {
      int slot = n % depth;
      if (n >= depth) mbar_wait(smem_u32(&bar_w[slot]), ((n / depth) - 1) & 1, 1);
      mbar_arrive_expect_tx(smem_u32(&bar_w[slot]), chunk_bytes);
      const uint8_t* s = src + (size_t)(n % nchunks_total) * chunk_bytes;
      uint32_t dst = smem_u32(base + (size_t)slot * chunk_bytes);
      if (csize == 1) bulk_g2s(dst, s, chunk_bytes, smem_u32(&bar_w[slot]));
      else bulk_g2s_multicast(dst + crank * share, s + crank * share, share, smem_u32(&bar_w[slot]), (uint16_t)((1u << csize) - 1));
    }
    for (int n = reps - depth; n < reps; ++n) if (n >= 0) mbar_wait(smem_u32(&bar_w[n % depth]), (n / depth) & 1, 2);
    if (w == 0) out[blockIdx.x] = clock64() - t0;
  }
  __syncthreads();
  if (csize > 1) cluster_sync();
}

int main() {
  const int total = 65536 * 40;
  uint8_t* src; cudaMalloc(&src, total); cudaMemset(src, 1, total);
  long long* d; cudaMalloc(&d, 148 * sizeof(long long));
  cudaFuncSetAttribute(k_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, 220 * 1024);
  for (int grid : {148})
    for (int nw : {1, 2, 4})
      for (int chunk : {1024, 8192, 32768, 49152, 65536})
        for (int depth : {2, 3}) {
          if ((size_t)depth * chunk * nw > 200 * 1024) continue;
          int reps = 3000;
          size_t smem = (size_t)depth * chunk * nw + 2048;
          k_bulk<<<grid, 128, smem>>>(src, total / chunk, reps, chunk, depth, d, nw);
          cudaError_t e2 = cudaDeviceSynchronize();
          long long h[148]; cudaMemcpy(h, d, grid * sizeof(long long), cudaMemcpyDeviceToHost);
          long long mx = 0; for (int i = 0; i < grid; ++i) mx = h[i] > mx ? h[i] : mx;
          printf("grid=%3d warps=%d chunk=%5d depth=%d: %.1f B/clk/SM total, %.0f cycles/chunk/warp [%s]\n", grid, nw, chunk, depth,
                 (double)reps * chunk * nw / mx, (double)mx / reps, cudaGetErrorString(e2));
        }
  return 0;
}
