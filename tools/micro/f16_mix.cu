// Micro-test: tcgen05.mma kind::f16 with MIXED operand formats (A fp16 / B bf16 and vice versa), SS and TS form
// (A packed two halves per 32-bit TMEM column), M=128, N=64, K=64 (4 k-steps).  Checks D = A B^T against the host.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include "../../gims_b200/csrc/tc_common.cuh"
using namespace gims::tc;

__device__ __forceinline__ void mma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void mma_f16_ts(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f16(int m, int n, int afmt, int bfmt) {
  return (1u << 4) | ((uint32_t)afmt << 7) | ((uint32_t)bfmt << 10) | ((uint32_t)(n >> 3) << 17) | ((uint32_t)(m >> 4) << 24);
}

// A16 [128][64] and B16 [64][64] raw 16-bit patterns (row-major, K contiguous); out D [128][64] fp32
__global__ void __launch_bounds__(128, 1) k_test(const uint16_t* A16, const uint16_t* B16, int afmt, int bfmt, int ts, float* D) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sa = smem;             // 128 rows x 128 B, SWIZZLE_128B
  uint8_t* sb = smem + 16384;     //  64 rows x 128 B
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  const int t = threadIdx.x;
  for (int i = t; i < 128 * 64; i += 128) {
    int r = i / 64, k = i % 64;
    int off = r * 128 + ((((k * 2) >> 4) ^ (r & 7)) << 4) + ((k * 2) & 15);
    *reinterpret_cast<uint16_t*>(sa + off) = A16[i];
  }
  for (int i = t; i < 64 * 64; i += 128) {
    int r = i / 64, k = i % 64;
    int off = r * 128 + ((((k * 2) >> 4) ^ (r & 7)) << 4) + ((k * 2) & 15);
    *reinterpret_cast<uint16_t*>(sb + off) = B16[i];
  }
  if (t == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (t < 32) tmem_alloc<128>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = slot;
  const uint32_t lane_base = tmem + ((uint32_t)(32 * (t >> 5)) << 16);
  if (ts == 1 || ts == 2) {   // A row t -> TMEM columns 64..95: 32 words, word c = halves (2c, 2c+1), low half = even k
    uint32_t v[32];
    for (int c = 0; c < 32; ++c) v[c] = (uint32_t)A16[t * 64 + 2 * c] | ((uint32_t)A16[t * 64 + 2 * c + 1] << 16);
    tmem_st_32x32(lane_base + 64, v);
    tmem_st_wait();
    tcgen05_fence_before();
  }
  __syncthreads();
  tcgen05_fence_after();
  if (t == 0 && ts >= 2) {        // rate: 2000 back-to-back MMAs, TS (ts = 2) or SS (ts = 3), two accumulators alternating
    const uint32_t idesc = idesc_f16(128, 64, afmt, bfmt);
    const uint64_t da = umma_desc_sw128(smem_u32(sa)), db = umma_desc_sw128(smem_u32(sb));
    long long t0 = clock64();
    for (int i = 0; i < 2000; ++i) {
      const int ks = i & 3;
      const uint64_t off = (ks * 32) >> 4;
      if (ts == 2) mma_f16_ts(tmem + (i & 4 ? 0 : 0), tmem + 64 + ks * 8, db + off, idesc, 1u);
      else         mma_f16_ss(tmem, da + off, db + off, idesc, 1u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t1 = clock64();
    D[0] = (float)(t1 - t0) / 2000.f;
  } else if (t == 0) {
    const uint32_t idesc = idesc_f16(128, 64, afmt, bfmt);
    const uint64_t da = umma_desc_sw128(smem_u32(sa)), db = umma_desc_sw128(smem_u32(sb));
    for (int ks = 0; ks < 4; ++ks) {
      const uint64_t off = (ks * 32) >> 4;
      if (ts) mma_f16_ts(tmem, tmem + 64 + ks * 8, db + off, idesc, ks ? 1u : 0u);
      else    mma_f16_ss(tmem, da + off, db + off, idesc, ks ? 1u : 0u);
    }
    umma_commit(&bar);
    mbar_wait(&bar, 0);
  }
  __syncthreads();
  tcgen05_fence_after();
  if (ts >= 2) { __syncthreads(); if (t < 32) tmem_dealloc<128>(tmem); return; }
  for (int h = 0; h < 2; ++h) {
    uint32_t v[32];
    tmem_ld_32x32(lane_base + h * 32, v);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) D[t * 64 + h * 32 + i] = __uint_as_float(v[i]);
  }
  tcgen05_fence_before();
  __syncthreads();
  if (t < 32) tmem_dealloc<128>(tmem);
}

static float h2f(uint16_t b, int fmt) {
  if (fmt == 0) { __half h; memcpy(&h, &b, 2); return __half2float(h); }
  uint32_t u = (uint32_t)b << 16; float f; memcpy(&f, &u, 4); return f;
}
static uint16_t f2h(float f, int fmt) {
  if (fmt == 0) { __half h = __float2half_rn(f); uint16_t b; memcpy(&b, &h, 2); return b; }
  __nv_bfloat16 h = __float2bfloat16_rn(f); uint16_t b; memcpy(&b, &h, 2); return b;
}

int main(int argc, char** argv) {
  int only_ts = argc > 1 ? atoi(argv[1]) : -1, only_a = argc > 2 ? atoi(argv[2]) : -1, only_b = argc > 3 ? atoi(argv[3]) : -1;
  float bscale = argc > 4 ? atof(argv[4]) : 1.f;
  uint16_t *A, *B, *dA, *dB; float *D, *dD;
  A = (uint16_t*)malloc(128 * 64 * 2); B = (uint16_t*)malloc(64 * 64 * 2); D = (float*)malloc(128 * 64 * 4);
  cudaMalloc(&dA, 128 * 64 * 2); cudaMalloc(&dB, 64 * 64 * 2); cudaMalloc(&dD, 128 * 64 * 4);
  int smem = 16384 + 8192 + 1024;
  cudaFuncSetAttribute(k_test, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  for (int ts = 0; ts < 4; ++ts)
    for (int afmt = 0; afmt < 2; ++afmt)
      for (int bfmt = 0; bfmt < 2; ++bfmt) {
        if ((only_ts >= 0 && ts != only_ts) || (only_a >= 0 && afmt != only_a) || (only_b >= 0 && bfmt != only_b)) continue;
        srand(7);
        for (int i = 0; i < 128 * 64; ++i) A[i] = f2h((rand() % 2001 - 1000) / 997.f, afmt);
        for (int i = 0; i < 64 * 64; ++i) B[i] = f2h((rand() % 2001 - 1000) / 1013.f * ((i % 7 == 0) ? 1e-3f : 1.f) * bscale, bfmt);
        cudaMemcpy(dA, A, 128 * 64 * 2, cudaMemcpyHostToDevice); cudaMemcpy(dB, B, 64 * 64 * 2, cudaMemcpyHostToDevice);
        cudaMemset(dD, 0, 128 * 64 * 4);
        k_test<<<1, 128, smem>>>(dA, dB, afmt, bfmt, ts, dD);
        cudaError_t e = cudaDeviceSynchronize();
        cudaMemcpy(D, dD, 128 * 64 * 4, cudaMemcpyDeviceToHost);
        if (ts >= 2) { printf("%s f16 N=64 K=16: %.1f clk per MMA (%s)\n", ts == 2 ? "TS" : "SS", D[0], cudaGetErrorString(e)); continue; }
        double maxerr = 0, maxref = 0;
        for (int m = 0; m < 128; ++m)
          for (int n = 0; n < 64; ++n) {
            double s = 0;
            for (int k = 0; k < 64; ++k) s += (double)h2f(A[m * 64 + k], afmt) * (double)h2f(B[n * 64 + k], bfmt);
            maxerr = fmax(maxerr, fabs(s - D[m * 64 + n])); maxref = fmax(maxref, fabs(s));
          }
        printf("%s A=%s B=%s : max|err| %.3e (max|ref| %.3e) %s\n", ts ? "TS" : "SS", afmt ? "bf16" : "f16", bfmt ? "bf16" : "f16",
               maxerr, maxref, cudaGetErrorString(e));
      }
  return 0;
}
