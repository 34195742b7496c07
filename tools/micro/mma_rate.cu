// Microbenchmark: issue rate of tcgen05.mma (cta_group::1) for kind::tf32 / kind::f16 at M=128 and several N.
// Operands: whatever is in shared memory (zero-filled); one CTA per SM; the single issuing thread times
// `iters` back-to-back MMAs into one TMEM accumulator with clock64() and a final commit + wait.
#include <cstdio>
#include <cuda_runtime.h>
#include "../../gims_b200/csrc/tc_common.cuh"
using namespace gims::tc;

__device__ __forceinline__ void umma_f16_ss(uint32_t d, uint64_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}"
               ::"r"(d), "l"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}
__device__ __forceinline__ void umma_tf32_ts_(uint32_t d, uint32_t a, uint64_t b, uint32_t idesc, uint32_t acc) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.b32 p, %4, 0;\n\ttcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}"
               ::"r"(d), "r"(a), "l"(b), "r"(idesc), "r"(acc) : "memory");
}

template <int N, int KIND>   // KIND 0: tf32 SS, 1: bf16 SS, 2: tf32 TS
__global__ void __launch_bounds__(128, 1) k_rate(int iters, long long* out) {
  extern __shared__ uint8_t raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(raw) + 1023) & ~uintptr_t(1023));
  __shared__ uint64_t bar;
  __shared__ uint32_t slot;
  for (int i = threadIdx.x; i < (128 * 128 + 256 * 128) / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0;
  if (threadIdx.x == 0) { mbar_init(&bar, 1); fence_barrier_init(); }
  if (threadIdx.x < 32) tmem_alloc<512>(&slot);
  fence_proxy_async_smem();
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  uint32_t tmem = slot;
  if (threadIdx.x == 0) {
    uint64_t da = umma_desc_sw128(smem_u32(smem)), db = umma_desc_sw128(smem_u32(smem + 128 * 128));
    uint32_t idesc = KIND == 1 ? ((1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(N >> 3) << 17) | (8u << 24))
                               : umma_idesc_tf32(128, N);
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
      uint64_t koff = (i & 3) * 2;
      if (KIND == 0) umma_tf32_ss(tmem, da + koff, db + koff, idesc, 1u);
      else if (KIND == 1) umma_f16_ss(tmem, da + koff, db + koff, idesc, 1u);
      else umma_tf32_ts_(tmem, tmem + 256 + (i & 7) * 8, db + koff, idesc, 1u);
    }
    long long t1 = clock64();
    umma_commit(&bar);
    mbar_wait(&bar, 0);
    long long t2 = clock64();
    if (blockIdx.x == 0) { out[0] = t1 - t0; out[1] = t2 - t0; }
  }
  __syncthreads();
  if (threadIdx.x < 32) tmem_dealloc<512>(tmem);
}

template <int N, int KIND>
void run(const char* name, int grid) {
  long long* d; cudaMalloc(&d, 16);
  int smem = 128 * 128 + 256 * 128 + 1024;
  cudaFuncSetAttribute(k_rate<N, KIND>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  int iters = 2000;
  k_rate<N, KIND><<<grid, 128, smem>>>(iters, d);
  k_rate<N, KIND><<<grid, 128, smem>>>(iters, d);
  cudaError_t e = cudaDeviceSynchronize();
  long long h[2]; cudaMemcpy(h, d, 16, cudaMemcpyDeviceToHost);
  printf("%-10s N=%3d grid=%3d  issue %.1f clk/MMA   complete %.1f clk/MMA   (%s)\n", name, N, grid, (double)h[0] / iters,
         (double)h[1] / iters, cudaGetErrorString(e));
  cudaFree(d);
}

int main() {
  for (int grid : {1, 148}) {
    run<64, 0>("tf32 SS", grid); run<128, 0>("tf32 SS", grid); run<192, 0>("tf32 SS", grid); run<256, 0>("tf32 SS", grid);
    run<64, 2>("tf32 TS", grid); run<128, 2>("tf32 TS", grid); run<256, 2>("tf32 TS", grid);
    run<64, 1>("bf16 SS", grid); run<128, 1>("bf16 SS", grid); run<256, 1>("bf16 SS", grid);
  }
  return 0;
}
