// Microbenchmark: cost of one all-to-all "publish flag / wait for all flags" hop between 148 co-resident CTAs
// (the exchange primitive of k_sinkhorn), with and without a payload write + read.
#include <cstdio>
#include <cuda_runtime.h>
__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v; asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
__device__ __forceinline__ void publish_flag(unsigned* flag, unsigned epoch) {
  __syncthreads();
  if (threadIdx.x == 0) { __threadfence(); asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory"); }
}
__device__ __forceinline__ unsigned ld_relaxed_u32(const unsigned* p) {
  unsigned v; asm volatile("ld.relaxed.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory"); return v;
}
template <int MODE>
__device__ __forceinline__ void wait_flags(const unsigned* flags, int n, unsigned epoch) {
  if (MODE == 0) {           // original: short-circuit chain of acquire loads by one warp
    if (threadIdx.x < 32) {
      while (true) {
        bool ok = true;
        for (int g = threadIdx.x; g < n; g += 32) ok = ok && (ld_acquire_u32(&flags[g]) >= epoch);
        if (__all_sync(0xffffffffu, ok)) break;
      }
    }
  } else if (MODE == 1) {    // one flag per thread (n <= blockDim), relaxed loads, fence afterwards
    if (threadIdx.x < 256) {
      const bool mine = threadIdx.x < n;
      while (true) {
        bool ok = !mine || ld_relaxed_u32(&flags[threadIdx.x]) >= epoch;
        if (__all_sync(0xffffffffu, ok)) break;
      }
      __threadfence();
    }
  } else {                   // single counter instead of flags
    if (threadIdx.x == 0) {
      while (ld_acquire_u32(flags + 512) < epoch * (unsigned)n) {}
    }
  }
  __syncthreads();
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k_hop(unsigned* flags, float* data, int iters, int payload, long long* out) {
  __shared__ float sink[512];
  const int G = gridDim.x, b = blockIdx.x;
  long long t0 = clock64();
  float acc = 0.f;
  for (int it = 1; it <= iters; ++it) {
    if (payload) for (int j = threadIdx.x; j < payload; j += blockDim.x) data[(size_t)b * payload + j] = (float)(it + j);
    if (MODE == 2) {
      __syncthreads();
      if (threadIdx.x == 0) { __threadfence(); atomicAdd(flags + 512, 1u); }
    } else {
      publish_flag(&flags[b], (unsigned)it);
    }
    wait_flags<MODE>(flags, G, (unsigned)it);
    if (payload) {   // read 14 values from every producer, like the Sinkhorn combine step
      for (int g = threadIdx.x >> 4; g < G; g += 32) acc += __ldcg(&data[(size_t)g * payload + (b * 14 + (threadIdx.x & 15)) % payload]);
    }
  }
  sink[threadIdx.x] = acc;
  long long t1 = clock64();
  if (b == 0 && threadIdx.x == 0) out[0] = t1 - t0;
  if (sink[threadIdx.x] == 12345.f) out[1] = 1;
}
int main() {
  unsigned* flags; float* data; long long* out;
  cudaMalloc(&flags, 4096); cudaMalloc(&data, 148 * 2049 * 4); cudaMalloc(&out, 16);
  int iters = 400;
  for (int mode = 0; mode < 3; ++mode)
  for (int payload : {0, 2049}) {
    for (int rep = 0; rep < 2; ++rep) {
      cudaMemset(flags, 0, 4096);
      void* args[] = {&flags, &data, &iters, &payload, &out};
      const void* fn = mode == 0 ? (const void*)k_hop<0> : mode == 1 ? (const void*)k_hop<1> : (const void*)k_hop<2>;
      cudaLaunchCooperativeKernel(fn, dim3(148), dim3(512), args, 0, 0);
      cudaDeviceSynchronize();
    }
    long long h[2]; cudaMemcpy(h, out, 16, cudaMemcpyDeviceToHost);
    printf("mode %d payload %4d floats/CTA: %.0f cycles per hop (%s)\n", mode, payload, (double)h[0] / iters, cudaGetErrorString(cudaGetLastError()));
  }
  return 0;
}
