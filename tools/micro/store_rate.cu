// Microbenchmark: how fast can one CTA (128 threads) put a 128 x BN fp32 tile that sits in shared memory into a
// row-major global matrix?  (The GEMM epilogue spent ~1800 clk per 128x32 chunk in its st.global loop.)
//   mode 0: st.global.v4, one instruction = four complete 128-byte lines (the current epilogue)
//   mode 1: one cp.async.bulk.global.shared::cta per tile row (BN*4 bytes), issued by the row's thread
//   mode 2: st.global.v4, one instruction = one 512-byte row segment (lanes along the row)
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/micro/store_rate tools/micro/store_rate.cu
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>

template <int BN, int MODE>
__global__ void __launch_bounds__(128) k_store(float* Y, int ldy, long long* clk, int reps) {
  extern __shared__ __align__(128) float tile[];      // [128][BN]
  const int t = threadIdx.x, lane = t & 31, warp = t >> 5;
  for (int i = t; i < 128 * BN; i += 128) tile[i] = (float)(i + blockIdx.x);
  __syncthreads();
  float* y0 = Y + (size_t)blockIdx.y * 128 * ldy + blockIdx.x * BN;
  long long t0 = clock64();
  for (int rep = 0; rep < reps; ++rep) {
    if (MODE == 0) {
      for (int cc = 0; cc < BN / 32; ++cc) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          int r = 32 * warp + i * 4 + (lane >> 3), c = cc * 32 + (lane & 7) * 4;
          float4 v = *reinterpret_cast<const float4*>(tile + r * BN + c);
          *reinterpret_cast<float4*>(y0 + (size_t)r * ldy + c) = v;
        }
      }
    } else if (MODE == 2) {
      for (int rr = 0; rr < 32; ++rr) {
        int r = 32 * warp + rr;
        for (int c = lane * 4; c < BN; c += 128) {
          float4 v = *reinterpret_cast<const float4*>(tile + r * BN + c);
          *reinterpret_cast<float4*>(y0 + (size_t)r * ldy + c) = v;
        }
      }
    } else {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
      uint32_t src = (uint32_t)__cvta_generic_to_shared(tile + t * BN);
      asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(y0 + (size_t)t * ldy), "r"(src),
                   "r"(BN * 4)
                   : "memory");
      asm volatile("cp.async.bulk.commit_group;" ::: "memory");
      asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory");
    }
  }
  if (MODE == 1) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
  __syncthreads();
  long long t1 = clock64();
  if (t == 0) clk[blockIdx.y * gridDim.x + blockIdx.x] = t1 - t0;
}

template <int BN, int MODE>
void run(int row_tiles, int N, float* Y, long long* clk) {
  dim3 grid(N / BN, row_tiles);
  int reps = 4;
  int smem = 128 * BN * 4;
  cudaFuncSetAttribute(k_store<BN, MODE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  k_store<BN, MODE><<<grid, 128, smem>>>(Y, N, clk, reps);
  cudaDeviceSynchronize();
  cudaEvent_t e0, e1;
  cudaEventCreate(&e0); cudaEventCreate(&e1);
  cudaEventRecord(e0);
  for (int i = 0; i < 20; ++i) k_store<BN, MODE><<<grid, 128, smem>>>(Y, N, clk, reps);
  cudaEventRecord(e1);
  cudaError_t err = cudaDeviceSynchronize();
  float ms; cudaEventElapsedTime(&ms, e0, e1);
  int n = grid.x * grid.y;
  long long* h = new long long[n];
  cudaMemcpy(h, clk, n * sizeof(long long), cudaMemcpyDeviceToHost);
  double avg = 0; for (int i = 0; i < n; ++i) avg += h[i];
  avg /= n;
  double bytes = 128.0 * BN * 4;
  printf("BN=%3d mode=%d ctas=%4d: %8.0f clk per tile pass (%5.1f B/clk/SM), kernel %.2f us, %s\n", BN, MODE, n, avg / reps,
         bytes / (avg / reps), ms * 1000 / 20, cudaGetErrorString(err));
  delete[] h;
}

int main() {
  const int N = 768, rows = 4096;
  float* Y; long long* clk;
  cudaMalloc(&Y, (size_t)rows * N * 4);
  cudaMalloc(&clk, 4096 * sizeof(long long));
  for (int tiles : {4, 32}) {
    run<64, 0>(tiles, 256, Y, clk);
    run<64, 2>(tiles, 256, Y, clk);
    run<64, 1>(tiles, 256, Y, clk);
    run<128, 0>(tiles, 512, Y, clk);
    run<128, 1>(tiles, 512, Y, clk);
    run<192, 0>(tiles, 768, Y, clk);
    run<192, 2>(tiles, 768, Y, clk);
    run<192, 1>(tiles, 768, Y, clk);
  }
  return 0;
}
