// Microbenchmark for kernel A's stage pipeline: does the in-place epilogue of layer l (tcgen05.ld -> fp32 math -> tcgen05.st,
// 8 warps) overlap with the MMAs of layer l+1 (48 x M128 N256 K16, A operand from the slabs the epilogue has converted)?
//   mode 0: a_ready[0..3] signalled at the END of the epilogue (serial schedule)      mode 1: per 64-column slab (pipelined)
//   asrc 0: A from TMEM (.ts)    asrc 1: A from shared memory (.ss)
//   work 0: bare ld/st           work 1: fma + relu + fp16 hi/lo split per element (kernel A's epilogue arithmetic)
//   tma 0: B operand = one resident 32 KB tile     tma 1: B streamed through a 5-slot ring of 32 KB bulk copies (two producer
//          warps, full/empty mbarriers, one slot per 4 MMAs like kernel A's W_lo plane / fast mode; 2 = one slot per 8 MMAs)
#include <cstdio>
#include <cstdint>
#include <cuda_fp16.h>
#include "../nerf-sos_b200/csrc/tc_ptx.cuh"
using namespace nsos::ptx;

__global__ void __launch_bounds__(352, 1) k_pipe(int mode, int asrc, int work, int nstage, int mma_per_slab, long long* out, int tma, const uint8_t* wsrc) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t acc_full, a_ready[4], full[5], empty[5];
  const int per_slot = tma == 2 ? 8 : 4;
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 65536 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x04000400u;   // tiny fp16 values: no overflow
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&acc_full), 1);
    for (int j = 0; j < 4; ++j) mbar_init(smem_u32(&a_ready[j]), 256);
    for (int j = 0; j < 5; ++j) { mbar_init(smem_u32(&full[j]), 1); mbar_init(smem_u32(&empty[j]), 1); }
    fence_mbar_init();
  }
  if (warp == 8) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tptr;
  if (warp == 8) {
    const uint32_t idesc = make_idesc_f16(256);
    const uint32_t b0 = smem_u32(base), a0 = smem_u32(base) + 32768, ring = smem_u32(base) + 65536;
    uint32_t slot = 0, par = 0, inslot = 0;
    long long t0 = clock64();
    for (int st = 0; st < nstage; ++st) {
      const uint32_t dcol = tm + (st & 1) * 256, acol = tm + ((st & 1) ^ 1) * 256;
      if (elect_one()) {
        for (int j = 0; j < 4; ++j) {
          mbar_wait(smem_u32(&a_ready[j]), st & 1, 10 + j);
          tc_fence_after();
          for (int m = 0; m < mma_per_slab; ++m) {
            if (tma && inslot == 0) { mbar_wait(smem_u32(&full[slot]), par, 30 + slot); tc_fence_after(); }
            const uint64_t bd = make_sw128_desc((tma ? ring + slot * 32768 : b0) + (m & 3) * 32);
            if (asrc == 0) umma_ts(dcol, acol + j * 64 + (m & 3) * 16 + ((m >> 2) & 1) * 8, bd, idesc, (j | m) ? 1u : 0u);
            else umma_ss(dcol, make_sw128_desc(a0 + (m & 3) * 32), bd, idesc, (j | m) ? 1u : 0u);
            if (tma && ++inslot == per_slot) { umma_commit(smem_u32(&empty[slot])); inslot = 0; if (++slot == 5) { slot = 0; par ^= 1; } }
          }
        }
        umma_commit(smem_u32(&acc_full));
      }
      __syncwarp();
    }
    mbar_wait(smem_u32(&acc_full), (nstage - 1) & 1, 1);   // approximately: last stage done
    if (lane == 0) out[0] = clock64() - t0;
  } else if (warp >= 9 && tma) {
    const int pidx = warp - 9;
    const int total = nstage * 4 * mma_per_slab / per_slot;
    for (int c = 0; c < total; ++c) {
      if ((c & 1) != pidx) continue;
      const uint32_t slot = c % 5, par = (c / 5) & 1;
      mbar_wait(smem_u32(&empty[slot]), par ^ 1u, 40 + slot);
      if (elect_one()) {
        mbar_arrive_expect_tx(smem_u32(&full[slot]), 32768);
        bulk_g2s(smem_u32(base) + 65536 + slot * 32768, wsrc + (size_t)(c % 128) * 32768, 32768, smem_u32(&full[slot]));
      }
      __syncwarp();
    }
  } else if (warp < 8) {
    const int q4 = warp & 3, hf = warp >> 2;
    const uint32_t tl = tm + ((uint32_t)(q4 * 32) << 16);
    for (int j = 0; j < 4; ++j) mbar_arrive(smem_u32(&a_ready[j]));     // stage 0 may start
    long long tepi = 0;
    for (int st = 0; st + 1 < nstage; ++st) {
      mbar_wait(smem_u32(&acc_full), st & 1, 2);
      tc_fence_after();
      long long e0 = clock64();
      const uint32_t td = tl + (st & 1) * 256;
      for (int j = 0; j < 4; ++j) {
        for (int cc = 0; cc < 2; ++cc) {
          const int c0 = (4 * j + 2 * hf + cc) * 16;
          uint32_t v[16]; tmem_ld16(td + c0, v); tmem_wait_ld_fence16(v);
          uint32_t h[8], l[8];
          for (int k = 0; k < 16; k += 2) {
            if (work) {
              float x0 = fmaxf(fmaf(__uint_as_float(v[k]), 1e-3f, 0.25f), 0.f), x1 = fmaxf(fmaf(__uint_as_float(v[k + 1]), 1e-3f, 0.25f), 0.f);
              __half2 hh = __floats2half2_rn(x0, x1);
              float2 hf2 = __half22float2(hh);
              __half2 ll = __floats2half2_rn(x0 - hf2.x, x1 - hf2.y);
              h[k >> 1] = *reinterpret_cast<uint32_t*>(&hh); l[k >> 1] = *reinterpret_cast<uint32_t*>(&ll);
            } else { h[k >> 1] = v[k]; l[k >> 1] = v[k + 1]; }
          }
          tmem_st8(td + c0, h); tmem_st8(td + c0 + 8, l);
        }
        if (mode == 1) { tmem_wait_st(); tc_fence_before(); mbar_arrive(smem_u32(&a_ready[j])); }
      }
      if (mode == 0) { tmem_wait_st(); tc_fence_before(); for (int j = 0; j < 4; ++j) mbar_arrive(smem_u32(&a_ready[j])); }
      tepi += clock64() - e0;
    }
    if (threadIdx.x == 0) out[1] = tepi;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  size_t smem = 65536 + 5 * 32768 + 2048;
  uint8_t* wsrc; cudaMalloc(&wsrc, 128 * 32768); cudaMemset(wsrc, 0x04, 128 * 32768);
  cudaFuncSetAttribute(k_pipe, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const int nstage = 200;
  for (int tma : {0, 1, 2})
  for (int mps : {12, 4})
    for (int asrc : {0, 1})
      for (int work : {0, 1})
        for (int mode : {0, 1}) {
          if (tma && asrc) continue;
          cudaMemset(d, 0, 64);
          k_pipe<<<1, 352, smem>>>(mode, asrc, work, nstage, mps, d, tma, wsrc);
          cudaError_t e = cudaDeviceSynchronize();
          long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
          printf("tma=%d mma/slab=%2d A=%s work=%d %s: %.0f cycles per stage (MMA floor %d), epilogue %.0f  [%s]\n", tma, mps, asrc ? "smem" : "tmem", work,
                 mode ? "pipelined" : "serial   ", (double)h[0] / nstage, mps * 4 * 129, (double)h[1] / (nstage - 1), cudaGetErrorString(e));
        }
  return 0;
}
