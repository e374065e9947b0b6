// Microbenchmark: how much do ordinary warps slow down while the tensor pipe runs back-to-back tcgen05.mma (TS form,
// M128 N256 K16, B operand from shared memory) and/or the bulk-copy engine streams 32 KB weight slabs into shared memory?
// 8 worker warps (as in kernel A; warp w shares a scheduler with warps w+-4 and, for w%4==0, with the MMA-issuing warp 8)
// run one of four workloads and report their own cycle counts:
//   0 dependent FMA chain (registers only)   1 sincosf loop (ALU + code)   2 shared-memory LDS.128 sweep   3 L2 pointer chase
#include <cstdio>
#include <cstdint>
#include "../nerf-sos_b200/csrc/tc_ptx.cuh"
using namespace nsos::ptx;

__global__ void __launch_bounds__(352, 1) k_intf(int work, int with_mma, int throttle, int with_bulk, const uint8_t* wsrc, const int* chase,
                                                 long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);   // [0,32K) MMA B tile, [32K,96K) bulk slots, [96K,112K) LDS area
  __shared__ uint64_t bar, thr[2], bfull[2];
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < (112 * 1024) / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) {
    mbar_init(smem_u32(&bar), 1); mbar_init(smem_u32(&thr[0]), 1); mbar_init(smem_u32(&thr[1]), 1);
    mbar_init(smem_u32(&bfull[0]), 1); mbar_init(smem_u32(&bfull[1]), 1);
    stop = 0; fence_mbar_init();
  }
  if (warp == 8) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  const uint32_t tm = tptr;
  if (warp == 8) {
    if (with_mma) {
      const uint32_t idesc = make_idesc_f16(256);
      long long t0 = clock64(); long long n = 0;
      if (elect_one()) {
        uint32_t gi = 0;
        while (!stop) {
          for (int g = 0; g < 4; ++g) {
            if (throttle && gi >= 2) mbar_wait(smem_u32(&thr[gi & 1]), ((gi >> 1) - 1u) & 1u, 7);
            for (int r = 0; r < 4; ++r)
              umma_ts(tm, tm + 256 + ((g * 4 + r) % 16) * 8, make_sw128_desc(smem_u32(base) + r * 32), idesc, 1);
            if (throttle) umma_commit(smem_u32(&thr[gi & 1]));
            ++gi;
          }
          n += 16;
        }
        umma_commit(smem_u32(&bar));
        out[16] = n;
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), 0, 1);
      if (lane == 0) out[17] = clock64() - t0;
    }
  } else if (warp >= 9) {
    if (with_bulk) {
      const int pi = warp - 9;
      uint32_t ph = 0; long long n = 0;
      while (!stop) {
        if (elect_one()) {
          mbar_arrive_expect_tx(smem_u32(&bfull[pi]), 32768);
          bulk_g2s(smem_u32(base) + 32768 + pi * 32768, wsrc + ((n * 2 + pi) % 64) * 32768, 32768, smem_u32(&bfull[pi]));
        }
        __syncwarp();
        mbar_wait(smem_u32(&bfull[pi]), ph, 2);
        ph ^= 1u; ++n;
      }
      if (lane == 0) out[18 + pi] = n;
    }
  } else {
    asm volatile("bar.sync 1, 256;");
    // let the MMA / bulk streams reach steady state
    long long tw = clock64(); while (clock64() - tw < 20000) {}
    long long t0 = clock64();
    float acc = (float)lane * 1e-3f;
    if (work == 0) {
      for (int i = 0; i < 4096; ++i) acc = fmaf(acc, 1.0000001f, 1e-7f);
    } else if (work == 1) {
      for (int i = 0; i < 96; ++i) { float s, c; sincosf(acc * 37.f + (float)i, &s, &c); acc += s * c * 1e-3f; }
    } else if (work == 2) {
      const float4* p = reinterpret_cast<const float4*>(base + 96 * 1024);
      float4 a = make_float4(0, 0, 0, 0);
      for (int i = 0; i < 2048; ++i) { float4 v = p[(i * 32 + lane) & 1023]; a.x += v.x; a.y += v.y; a.z += v.z; a.w += v.w; }
      acc = a.x + a.y + a.z + a.w;
    } else {
      int j = lane;
      for (int i = 0; i < 48; ++i) j = __ldcg(&chase[j]);
      acc = (float)j;
    }
    long long t1 = clock64();
    if (lane == 0) { out[warp] = t1 - t0; out[8 + warp] = __float_as_int(acc); }
    asm volatile("bar.sync 1, 256;");
    if (threadIdx.x == 0) stop = 1;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 256);
  uint8_t* w; cudaMalloc(&w, 64 * 32768); cudaMemset(w, 0, 64 * 32768);
  const int NC = 1 << 20;
  int* hc = new int[NC];
  for (int i = 0; i < NC; ++i) hc[i] = (int)(((long long)i * 40503 + 12345) % NC);
  int* c; cudaMalloc(&c, NC * 4); cudaMemcpy(c, hc, NC * 4, cudaMemcpyHostToDevice);
  size_t smem = 113 * 1024;
  cudaFuncSetAttribute(k_intf, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  const char* names[4] = {"fma-chain", "sincosf", "lds.128", "l2-chase"};
  for (int work = 0; work < 4; ++work)
    for (int cfg = 0; cfg < 5; ++cfg) {
      int with_mma = (cfg == 1 || cfg == 2 || cfg == 4), throttle = (cfg == 2), with_bulk = (cfg == 3 || cfg == 4);
      cudaMemset(d, 0, 256);
      k_intf<<<1, 352, smem>>>(work, with_mma, throttle, with_bulk, w, c, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[20]; cudaMemcpy(h, d, 160, cudaMemcpyDeviceToHost);
      printf("%-10s mma=%d thr=%d bulk=%d | warp cycles:", names[work], with_mma, throttle, with_bulk);
      for (int i = 0; i < 8; ++i) printf(" %6lld", h[i]);
      printf(" | cyc/MMA %.1f  bulk copies %lld+%lld [%s]\n", h[16] ? (double)h[17] / h[16] : 0.0, h[18], h[19], cudaGetErrorString(e));
    }
  return 0;
}
