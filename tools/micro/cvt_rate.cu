// Throughput of the conversion / special-function instructions the softmax warps use (per SM, 8 warps resident).
#include <cstdio>
#include <cuda_runtime.h>
template <int OP>
__global__ void __launch_bounds__(256) k(float* out, int iters, long long* clk) {
  float x0 = threadIdx.x * 1e-3f + 0.5f, x1 = x0 + 0.25f, x2 = x0 + 0.5f, x3 = x0 + 0.75f;
  unsigned acc = 0;
  long long t0 = clock64();
  for (int i = 0; i < iters; ++i) {
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      if (OP == 0) {          // cvt.rn.f16x2.f32 (F2FP.PACK_AB)
        unsigned a, b;
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(a) : "f"(x0), "f"(x1));
        asm volatile("cvt.rn.f16x2.f32 %0, %1, %2;" : "=r"(b) : "f"(x2), "f"(x3));
        acc ^= a + b; x0 += 1e-3f; x2 += 1e-3f;
      } else if (OP == 1) {   // ex2
        float a, b;
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(a) : "f"(x0));
        asm volatile("ex2.approx.ftz.f32 %0, %1;" : "=f"(b) : "f"(x2));
        acc ^= __float_as_uint(a) + __float_as_uint(b); x0 += 1e-3f; x2 += 1e-3f;
      } else if (OP == 2) {   // cvt.f32.f16 (HADD2.F32)
        float a, b; unsigned w = __float_as_uint(x0) & 0x3fff3fff;
        asm volatile("{.reg .f16 l, h; mov.b32 {l, h}, %2; cvt.f32.f16 %0, l; cvt.f32.f16 %1, h;}" : "=f"(a), "=f"(b) : "r"(w));
        acc ^= __float_as_uint(a) + __float_as_uint(b); x0 += 1e-3f;
      } else if (OP == 3) {   // integer split: mask, sub, shift/rebias/clamp x2, pack
        unsigned b0 = __float_as_uint(x0), b1 = __float_as_uint(x2);
        unsigned h0 = b0 & 0xffffe000u, h1 = b1 & 0xffffe000u;
        float l0 = x0 - __uint_as_float(h0), l1 = x2 - __uint_as_float(h1);
        int e0 = max((int)(h0 >> 13) - 0x1C000, 0), e1 = max((int)(h1 >> 13) - 0x1C000, 0);
        int f0 = max((int)((__float_as_uint(l0) + 0x1000u) >> 13) - 0x1C000, 0), f1 = max((int)((__float_as_uint(l1) + 0x1000u) >> 13) - 0x1C000, 0);
        acc ^= (e0 | (e1 << 16)) + (f0 | (f1 << 16)); x0 += 1e-3f; x2 += 1e-3f;
      }
    }
  }
  long long t1 = clock64();
  out[blockIdx.x * blockDim.x + threadIdx.x] = __uint_as_float(acc);
  if (threadIdx.x == 0 && blockIdx.x == 0) *clk = t1 - t0;
}
int main() {
  float* o; long long* c; cudaMalloc(&o, 148 * 256 * 4); cudaMalloc(&c, 8);
  const char* names[] = {"cvt.rn.f16x2.f32 (2 per step)", "ex2.approx (2 per step)", "cvt.f32.f16 x2 (2 per step)", "integer split of 2 values (hi+lo words)"};
  int iters = 2000;
  for (int op = 0; op < 4; ++op) {
    for (int rep = 0; rep < 2; ++rep) {
      if (op == 0) k<0><<<148, 256>>>(o, iters, c); if (op == 1) k<1><<<148, 256>>>(o, iters, c);
      if (op == 2) k<2><<<148, 256>>>(o, iters, c); if (op == 3) k<3><<<148, 256>>>(o, iters, c);
    }
    cudaDeviceSynchronize(); long long h; cudaMemcpy(&h, c, 8, cudaMemcpyDeviceToHost);
    printf("%-44s %.2f clk per step per SM-quarter (8 warps/SM = 2 warps per scheduler)\n", names[op], (double)h / (iters * 8));
  }
  return 0;
}
