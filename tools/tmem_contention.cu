// Microbenchmark: cost of tcgen05.ld / tcgen05.st issued by 8 worker warps while the tensor pipe runs MMAs
// (same TMEM column split as kernel A: D cols 0..255, A cols 256..511).
#include <cstdio>
#include <cstdint>
#include "../nerf-sos_b200/csrc/tc_ptx.cuh"
using namespace nsos::ptx;

__device__ __forceinline__ void ld64(uint32_t taddr, uint32_t (&r)[64]) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x64.b32 "
      "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,"
      "%32,%33,%34,%35,%36,%37,%38,%39,%40,%41,%42,%43,%44,%45,%46,%47,%48,%49,%50,%51,%52,%53,%54,%55,%56,%57,%58,%59,%60,%61,%62,%63}, [%64];"
      : "=r"(r[0]),"=r"(r[1]),"=r"(r[2]),"=r"(r[3]),"=r"(r[4]),"=r"(r[5]),"=r"(r[6]),"=r"(r[7]),"=r"(r[8]),"=r"(r[9]),"=r"(r[10]),"=r"(r[11]),"=r"(r[12]),"=r"(r[13]),"=r"(r[14]),"=r"(r[15]),
        "=r"(r[16]),"=r"(r[17]),"=r"(r[18]),"=r"(r[19]),"=r"(r[20]),"=r"(r[21]),"=r"(r[22]),"=r"(r[23]),"=r"(r[24]),"=r"(r[25]),"=r"(r[26]),"=r"(r[27]),"=r"(r[28]),"=r"(r[29]),"=r"(r[30]),"=r"(r[31]),
        "=r"(r[32]),"=r"(r[33]),"=r"(r[34]),"=r"(r[35]),"=r"(r[36]),"=r"(r[37]),"=r"(r[38]),"=r"(r[39]),"=r"(r[40]),"=r"(r[41]),"=r"(r[42]),"=r"(r[43]),"=r"(r[44]),"=r"(r[45]),"=r"(r[46]),"=r"(r[47]),
        "=r"(r[48]),"=r"(r[49]),"=r"(r[50]),"=r"(r[51]),"=r"(r[52]),"=r"(r[53]),"=r"(r[54]),"=r"(r[55]),"=r"(r[56]),"=r"(r[57]),"=r"(r[58]),"=r"(r[59]),"=r"(r[60]),"=r"(r[61]),"=r"(r[62]),"=r"(r[63])
      : "r"(taddr) : "memory");
}
__device__ __forceinline__ void st32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31,%32};"
      :: "r"(taddr), "r"(r[0]),"r"(r[1]),"r"(r[2]),"r"(r[3]),"r"(r[4]),"r"(r[5]),"r"(r[6]),"r"(r[7]),"r"(r[8]),"r"(r[9]),"r"(r[10]),"r"(r[11]),"r"(r[12]),"r"(r[13]),"r"(r[14]),"r"(r[15]),
         "r"(r[16]),"r"(r[17]),"r"(r[18]),"r"(r[19]),"r"(r[20]),"r"(r[21]),"r"(r[22]),"r"(r[23]),"r"(r[24]),"r"(r[25]),"r"(r[26]),"r"(r[27]),"r"(r[28]),"r"(r[29]),"r"(r[30]),"r"(r[31]) : "memory");
}

// variant: 0 = x16 ld + 2 x8 st per 16 columns (kernel A today), 1 = x32 ld + 2 x16 st, 2 = x64 ld + 2 x32 st
// Each worker thread sweeps `cols` D columns per round (like one epilogue half), `rounds` times.
__global__ void __launch_bounds__(352, 1) k_cont(int variant, int with_mma, int N, int rounds, int cols, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* base = raw + ((1024u - (smem_u32(raw) & 1023u)) & 1023u);
  __shared__ uint64_t bar;
  __shared__ uint32_t tptr;
  __shared__ volatile int stop;
  int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  for (int i = threadIdx.x; i < 32768 / 4; i += blockDim.x) ((uint32_t*)base)[i] = 0x3c003c00u;
  if (threadIdx.x == 0) { mbar_init(smem_u32(&bar), 1); stop = 0; fence_mbar_init(); }
  if (warp == 8) { tmem_alloc(smem_u32(&tptr), 512); tmem_relinquish(); }
  fence_proxy_async_smem();
  tc_fence_before(); __syncthreads(); tc_fence_after();
  uint32_t tm = tptr;
  if (warp == 8) {
    if (with_mma) {
      uint32_t idesc = make_idesc_f16(N);
      long long t0 = clock64(); long long n = 0;
      if (elect_one()) {
        while (!stop) {
          for (int r = 0; r < 16; ++r)
            umma_ts(tm + 128, tm + 256 + (r % 16) * 8, make_sw128_desc(smem_u32(base) + (r % 4) * 32), idesc, 1);   // D half 1 (cols 128..)
          n += 16;
        }
        umma_commit(smem_u32(&bar));
        out[2] = n; 
      }
      __syncwarp();
      mbar_wait(smem_u32(&bar), 0, 1);
      if (lane == 0) out[1] = clock64() - t0;
    }
  } else if (warp < 8) {
    const int q4 = warp & 3, hf = warp >> 2;
    const uint32_t tl = tm + ((uint32_t)(q4 * 32) << 16);
    const int c_begin = hf * (cols / 2), c_end = c_begin + cols / 2;     // D half 0: columns [0, cols)
    long long t0 = clock64();
    uint32_t acc = 0;
    for (int r = 0; r < rounds; ++r) {
      if (variant == 0) {
        for (int c = c_begin; c < c_end; c += 16) {
          uint32_t v[16]; tmem_ld16(tl + c, v); tmem_wait_ld_fence16(v);
          uint32_t h[8], l[8];
          for (int j = 0; j < 8; ++j) { h[j] = v[2 * j] + acc; l[j] = v[2 * j + 1] ^ acc; }
          tmem_st8(tl + 256 + c / 2, h); tmem_st8(tl + 384 + c / 2, l);
          acc += h[0];
        }
      } else if (variant == 1) {
        for (int c = c_begin; c < c_end; c += 32) {
          uint32_t v[32]; tmem_ld32(tl + c, v); tmem_wait_ld_fence(v);
          uint32_t h[16], l[16];
          for (int j = 0; j < 16; ++j) { h[j] = v[2 * j] + acc; l[j] = v[2 * j + 1] ^ acc; }
          tmem_st16(tl + 256 + c / 2, h); tmem_st16(tl + 384 + c / 2, l);
          acc += h[0];
        }
      } else {
        for (int c = c_begin; c < c_end; c += 64) {
          uint32_t v[64]; ld64(tl + c, v); tmem_wait_ld();
          uint32_t h[32], l[32];
          for (int j = 0; j < 32; ++j) { h[j] = v[2 * j] + acc; l[j] = v[2 * j + 1] ^ acc; }
          st32(tl + 256 + c / 2, h); st32(tl + 384 + c / 2, l);
          acc += h[0];
        }
      }
      tmem_wait_st();
    }
    long long t1 = clock64();
    if (threadIdx.x == 0) { out[0] = t1 - t0; out[3] = acc; }
    asm volatile("bar.sync 1, 256;");
    if (threadIdx.x == 0) stop = 1;
  }
  tc_fence_before(); __syncthreads();
  if (warp == 8) tmem_dealloc(tm, 512);
}

int main() {
  long long* d; cudaMalloc(&d, 64);
  size_t smem = 32768 + 2048;
  cudaFuncSetAttribute(k_cont, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
  int rounds = 50, cols = 128;
  for (int with_mma : {0, 1})
    for (int variant : {0, 1, 2}) {
      cudaMemset(d, 0, 64);
      k_cont<<<1, 352, smem>>>(variant, with_mma, 128, rounds, cols, d);
      cudaError_t e = cudaDeviceSynchronize();
      long long h[4]; cudaMemcpy(h, d, 32, cudaMemcpyDeviceToHost);
      printf("mma=%d variant=%d (%s): worker sweep of %d columns: %.0f cycles;  MMA: %.1f cycles per N=128 MMA  [%s]\n", with_mma, variant,
             variant == 0 ? "ld.x16 st.x8" : variant == 1 ? "ld.x32 st.x16" : "ld.x64 st.x32", cols, (double)h[0] / rounds,
             h[2] ? (double)h[1] / h[2] : 0.0, cudaGetErrorString(e));
    }
  return 0;
}
