// Fused log-domain Sinkhorn + mutual-nearest-neighbour match extraction.
//
// Replaces log_optimal_transport / log_sinkhorn_iterations (models/gmatcher.py:41-69) and the match
// extraction of GMatcher.forward (gmatcher.py:284-294).
//
// One persistent cooperative kernel, one CTA per SM.  CTA b owns a contiguous slab of rows of the
// couplings matrix Z0 ((N+1) x (M+1), dustbin row/column included) and keeps it in shared memory for
// all iterations when it fits (2049x2049 fp32 = 16.8 MB = 148 x 113.5 KB), so Z0 is read from HBM
// exactly once; otherwise the slab is streamed from L2/HBM every pass.  Per iteration:
//   row pass   u_r = log_mu_r - LSE_j(Z0[r][j] + v_j)            (slab-local, warp per row)
//   col pass   per-CTA partial (max, sum) of LSE_i(Z0[i][j] + u_i) -> global -> grid barrier ->
//              each CTA combines the partials of its column range -> v_j -> grid barrier.
// The last pass evaluates the reference's expression ((Z0 + u) + v) - norm element-wise and takes
// row / column max + first argmax; a tiny follow-up kernel applies the mutual / threshold rule.
#include <math_constants.h>

#include "common.cuh"

namespace gims {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxGrid = 256;
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

struct SinkArgs {
  const float* Z; int ld;
  int n0_max, n1_max;
  const int* n_dev;
  int iters;
  float* u; float* v;              // global potentials (v is also the exchange buffer)
  float2* part;                    // [grid][ldp] per-CTA column partials
  int ldp;
  unsigned* barrier;               // monotonically increasing arrival counter (zeroed before launch)
  int* idx0; int* idx1; float* max0; float* max1;
  int rpc_max;                     // ceil((n0_max+1)/grid): capacity of the per-CTA u buffer
  int slab_rows;                   // rows of the smem slab (0 = stream from global)
  int slab_ld;                     // padded row length of the slab
};

__device__ __forceinline__ void grid_barrier(unsigned* counter, unsigned& epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    epoch += 1;
    unsigned target = epoch * gridDim.x;
    __threadfence();
    atomicAdd(counter, 1u);
    unsigned spins = 0;
    while (true) {
      unsigned cur;
      asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(cur) : "l"(counter) : "memory");
      if (cur >= target) break;
      if (++spins > (1u << 27)) __trap();     // a lost CTA must fail loudly, never hang the GPU
    }
    __threadfence();
  }
  __syncthreads();
}

__device__ __forceinline__ float warp_max(float x) {
#pragma unroll
  for (int o = 16; o; o >>= 1) x = fmaxf(x, __shfl_xor_sync(0xffffffffu, x, o));
  return x;
}
__device__ __forceinline__ float warp_sum(float x) {
#pragma unroll
  for (int o = 16; o; o >>= 1) x += __shfl_xor_sync(0xffffffffu, x, o);
  return x;
}

__global__ void __launch_bounds__(kThreads, 1) k_sinkhorn(SinkArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float2 red[kWarps][32];
  const int n0 = a.n_dev ? min(a.n_dev[0], a.n0_max) : a.n0_max;
  const int n1 = a.n_dev ? min(a.n_dev[1], a.n1_max) : a.n1_max;
  const int R = n0 + 1, C = n1 + 1;
  const int G = gridDim.x, b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  unsigned epoch = 0;

  // marginals exactly as gmatcher.py:62-64 computes them in fp32
  const float ms = (float)n0, ns = (float)n1;
  const float norm = -logf(ms + ns);
  const float log_mu_last = logf(ns) + norm, log_nu_last = logf(ms) + norm;

  const int rpc = (R + G - 1) / G;                 // rows per CTA
  const int r_begin = min(b * rpc, R), r_end = min(r_begin + rpc, R);
  const int nrows = r_end - r_begin;
  const int Ga = (R + rpc - 1) / rpc;              // CTAs that own at least one row
  const int cpc = (C + G - 1) / G;                 // columns per CTA in the combine step
  const int c_begin = min(b * cpc, C), c_end = min(c_begin + cpc, C);

  float* v_s = smem;                                // [C]
  float* u_s = v_s + ((a.n1_max + 1 + 3) & ~3);     // [rpc_max]
  float* slab = u_s + ((a.rpc_max + 3) & ~3);
  const bool resident = a.slab_rows > 0 && nrows <= a.slab_rows && C <= a.slab_ld;
  const int zs = resident ? a.slab_ld : a.ld;       // row stride used by the passes
  const float* zbase = resident ? slab : a.Z + (size_t)r_begin * a.ld;

  if (resident) {
    for (int r = warp; r < nrows; r += kWarps) {
      const float* src = a.Z + (size_t)(r_begin + r) * a.ld;
      float* dst = slab + (size_t)r * a.slab_ld;
      for (int j = lane; j < C; j += 32) dst[j] = src[j];
    }
  }
  for (int j = tid; j < C; j += kThreads) v_s[j] = 0.f;
  for (int j = c_begin + tid; j < c_end; j += kThreads) a.v[j] = 0.f;   // stays if iters == 0
  __syncthreads();

  for (int it = 0; it < a.iters; ++it) {
    // ---- row pass -------------------------------------------------------------------------
    for (int r = warp; r < nrows; r += kWarps) {
      const float* z = zbase + (size_t)r * zs;
      float m = -CUDART_INF_F;
      for (int j = lane; j < C; j += 32) m = fmaxf(m, z[j] + v_s[j]);
      m = warp_max(m);
      float s = 0.f;
      for (int j = lane; j < C; j += 32) s += exp2f((z[j] + v_s[j] - m) * kLog2e);
      s = warp_sum(s);
      if (lane == 0) {
        float lmu = (r_begin + r == n0) ? log_mu_last : norm;
        u_s[r] = lmu - (m + logf(s));
      }
    }
    __syncthreads();
    // ---- column pass: per-CTA partial LSE ----------------------------------------------------
    if (b < Ga) {
      for (int j = tid; j < C; j += kThreads) {
        float m = -CUDART_INF_F;
        for (int r = 0; r < nrows; ++r) m = fmaxf(m, zbase[(size_t)r * zs + j] + u_s[r]);
        float s = 0.f;
        for (int r = 0; r < nrows; ++r) s += exp2f((zbase[(size_t)r * zs + j] + u_s[r] - m) * kLog2e);
        a.part[(size_t)b * a.ldp + j] = make_float2(m, s);
      }
    }
    grid_barrier(a.barrier, epoch);
    // ---- combine the partials of my column range -> v ---------------------------------------
    for (int jc = c_begin; jc < c_end; jc += 32) {
      int j = jc + lane;
      float m = -CUDART_INF_F, s = 0.f;
      if (j < c_end) {
        for (int g = warp; g < Ga; g += kWarps) {
          float2 p = __ldcg(&a.part[(size_t)g * a.ldp + j]);
          float mn = fmaxf(m, p.x);
          s = s * exp2f((m - mn) * kLog2e) + p.y * exp2f((p.x - mn) * kLog2e);
          m = mn;
        }
      }
      red[warp][lane] = make_float2(m, s);
      __syncthreads();
      if (warp == 0 && j < c_end) {
        float mm = -CUDART_INF_F;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) mm = fmaxf(mm, red[w][lane].x);
        float ss = 0.f;
#pragma unroll
        for (int w = 0; w < kWarps; ++w) {
          float2 p = red[w][lane];
          if (p.y > 0.f) ss += p.y * exp2f((p.x - mm) * kLog2e);
        }
        float lnu = (j == n1) ? log_nu_last : norm;
        a.v[j] = lnu - (mm + logf(ss));
      }
      __syncthreads();
    }
    grid_barrier(a.barrier, epoch);
    for (int j = tid; j < C; j += kThreads) v_s[j] = __ldcg(&a.v[j]);
    __syncthreads();
  }

  // ---- final pass: Z = ((Z0 + u) + v) - norm, row / column max + first argmax -------------------
  for (int r = warp; r < nrows; r += kWarps) {
    int gr = r_begin + r;
    float ur = u_s[r];
    if (lane == 0) a.u[gr] = (a.iters > 0) ? ur : 0.f;
    if (gr >= n0) continue;
    if (a.iters == 0) ur = 0.f;
    const float* z = zbase + (size_t)r * zs;
    float best = -CUDART_INF_F;
    int bj = 0x7fffffff;
    for (int j = lane; j < n1; j += 32) {
      float t = ((z[j] + ur) + v_s[j]) - norm;
      if (t > best) { best = t; bj = j; }
    }
#pragma unroll
    for (int o = 16; o; o >>= 1) {
      float ob = __shfl_xor_sync(0xffffffffu, best, o);
      int oj = __shfl_xor_sync(0xffffffffu, bj, o);
      if (ob > best || (ob == best && oj < bj)) { best = ob; bj = oj; }
    }
    if (lane == 0) { a.idx0[gr] = bj; a.max0[gr] = best; }
  }
  if (b < Ga) {
    int live = min(r_end, n0) - r_begin;           // rows of mine that are real keypoints
    for (int j = tid; j < n1; j += kThreads) {
      float best = -CUDART_INF_F;
      int bi = 0x7fffffff;
      float vj = v_s[j];
      for (int r = 0; r < live; ++r) {
        float ur = (a.iters > 0) ? u_s[r] : 0.f;
        float t = ((zbase[(size_t)r * zs + j] + ur) + vj) - norm;
        if (t > best) { best = t; bi = r_begin + r; }
      }
      a.part[(size_t)b * a.ldp + j] = make_float2(best, __int_as_float(bi));
    }
  }
  grid_barrier(a.barrier, epoch);
  for (int jc = c_begin; jc < c_end; jc += 32) {
    int j = jc + lane;
    float best = -CUDART_INF_F;
    int bi = 0x7fffffff;
    if (j < c_end && j < n1) {
      for (int g = warp; g < Ga; g += kWarps) {
        float2 p = __ldcg(&a.part[(size_t)g * a.ldp + j]);
        int pi = __float_as_int(p.y);
        if (p.x > best || (p.x == best && pi < bi)) { best = p.x; bi = pi; }
      }
    }
    red[warp][lane] = make_float2(best, __int_as_float(bi));
    __syncthreads();
    if (warp == 0 && j < c_end && j < n1) {
      float bb = -CUDART_INF_F;
      int ii = 0x7fffffff;
#pragma unroll
      for (int w = 0; w < kWarps; ++w) {
        float2 p = red[w][lane];
        int pi = __float_as_int(p.y);
        if (p.x > bb || (p.x == bb && pi < ii)) { bb = p.x; ii = pi; }
      }
      a.idx1[j] = ii;
      a.max1[j] = bb;
    }
    __syncthreads();
  }
}

// gmatcher.py:286-294
__global__ void k_match_finalize(int n0_max, int n1_max, const int* __restrict__ n_dev, const int* __restrict__ idx0,
                                 const int* __restrict__ idx1, const float* __restrict__ max0, float thr,
                                 int64_t* __restrict__ matches0, int64_t* __restrict__ matches1,
                                 float* __restrict__ ms0, float* __restrict__ ms1) {
  int n0 = n_dev ? min(n_dev[0], n0_max) : n0_max;
  int n1 = n_dev ? min(n_dev[1], n1_max) : n1_max;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t < n0) {
    int j = idx0[t];
    bool mutual = idx1[j] == t;
    float s = mutual ? expf(max0[t]) : 0.f;
    ms0[t] = s;
    matches0[t] = (mutual && s > thr) ? (int64_t)j : (int64_t)-1;
  }
  if (t < n1) {
    int i = idx1[t];
    bool mutual = idx0[i] == t;
    float s = mutual ? expf(max0[i]) : 0.f;      // mutual1 implies mutual0[i]
    ms1[t] = s;
    matches1[t] = (mutual && s > thr) ? (int64_t)i : (int64_t)-1;
  }
}

struct SinkWs {
  float2* part;
  unsigned* barrier;
  float* max0;
  float* max1;
};

size_t carve(SinkWs& w, void* base, size_t cap, int n0_max, int n1_max) {
  Arena a(base, cap);
  w.barrier = a.take<unsigned>(64);
  w.part = a.take<float2>((size_t)kMaxGrid * (n1_max + 1));
  w.max0 = a.take<float>(n0_max + 1);
  w.max1 = a.take<float>(n1_max + 1);
  return align_up(a.off, 256);
}

}  // namespace

}  // namespace gims

using namespace gims;

extern "C" size_t gims_sinkhorn_workspace_bytes(int n0_max, int n1_max) {
  SinkWs w;
  return carve(w, nullptr, 0, n0_max, n1_max);
}

extern "C" int gims_sinkhorn_match(const float* couplings, int n0_max, int n1_max, const int* n_dev, int iters,
                                   float match_threshold, void* workspace, size_t workspace_bytes, float* u, float* v,
                                   int* indices0, int* indices1, int64_t* matches0, int64_t* matches1, float* mscores0,
                                   float* mscores1, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n0_max < 1 || n1_max < 1 || iters < 0) { set_error("gims_sinkhorn_match: bad sizes"); return GIMS_ERR_ARG; }
  SinkWs w;
  size_t need = carve(w, workspace, workspace_bytes, n0_max, n1_max);
  if (need > workspace_bytes) { set_error("gims_sinkhorn_match: workspace %zu < %zu", workspace_bytes, need); return GIMS_ERR_WORKSPACE; }
  int dev = 0, sms = 0, smem_optin = 0;
  GIMS_CUDA_OK(cudaGetDevice(&dev));
  GIMS_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GIMS_CUDA_OK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int G = sms < kMaxGrid ? sms : kMaxGrid;
  int R = n0_max + 1, C = n1_max + 1;
  int rpc = (R + G - 1) / G;
  int slab_ld = (C + 3) & ~3;
  size_t fixed = (size_t)(((C + 3) & ~3) + ((rpc + 3) & ~3)) * sizeof(float);
  size_t slab_bytes = (size_t)rpc * slab_ld * sizeof(float);
  size_t budget = (size_t)smem_optin - 4608;     // static smem (red[]) + margin
  SinkArgs a;
  a.rpc_max = rpc;
  a.slab_rows = (fixed + slab_bytes <= budget) ? rpc : 0;
  a.slab_ld = slab_ld;
  size_t dyn = fixed + (a.slab_rows ? slab_bytes : 0);
  if (dyn > budget) { set_error("gims_sinkhorn_match: n1_max=%d needs %zu B of shared memory", n1_max, dyn); return GIMS_ERR_ARG; }
  a.Z = couplings; a.ld = C; a.n0_max = n0_max; a.n1_max = n1_max; a.n_dev = n_dev; a.iters = iters;
  a.u = u; a.v = v; a.part = w.part; a.ldp = C; a.barrier = w.barrier;
  a.idx0 = indices0; a.idx1 = indices1; a.max0 = w.max0; a.max1 = w.max1;
  GIMS_CUDA_OK(cudaFuncSetAttribute(k_sinkhorn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  int per_sm = 0;
  GIMS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sinkhorn, kThreads, dyn));
  if (per_sm < 1) { set_error("gims_sinkhorn_match: kernel does not fit an SM (dyn smem %zu)", dyn); return GIMS_ERR_ARG; }
  GIMS_CUDA_OK(cudaMemsetAsync(w.barrier, 0, 64 * sizeof(unsigned), st));
  void* params[] = {&a};
  GIMS_TRY(coop_chain_wait(st));
  {
    ProfScope prof(GIMS_PROF_SINKHORN, st);
    GIMS_CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_sinkhorn, dim3(G), dim3(kThreads), params, dyn, st));
  }
  GIMS_TRY(coop_chain_record(st));
  count_launch();
  int m = n0_max > n1_max ? n0_max : n1_max;
  k_match_finalize<<<cdiv(m, 256), 256, 0, st>>>(n0_max, n1_max, n_dev, indices0, indices1, w.max0, match_threshold,
                                                 matches0, matches1, mscores0, mscores1);
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}
