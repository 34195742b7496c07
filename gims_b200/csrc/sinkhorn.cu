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
//   col pass   per-CTA partial (max, sum) of LSE_i(Z0[i][j] + u_i) for every column
//   exchange   there is NO grid-wide barrier: partials and the new v_j travel through L2 as single
//              64-bit words that carry their own tag (iteration parity in the sign bit of the always
//              non-negative partial sum; iteration number next to v_j), and consumers simply poll the
//              words they need.  Data-flow dependencies make slot reuse safe (a producer can only
//              overwrite a slot after it has received every v of the iteration that consumed it).
// The last pass evaluates the reference's expression ((Z0 + u) + v) - norm element-wise and takes
// row / column max + first argmax; a tiny follow-up kernel applies the mutual / threshold rule.
#include <math_constants.h>

#include "common.cuh"

namespace gims {

namespace {

constexpr int kThreads = 512;
constexpr int kWarps = kThreads / 32;
constexpr int kMaxGrid = 256;
constexpr int kRMax = 16;            // rows per CTA handled by the unrolled (register) column pass
constexpr float kLog2e = 1.4426950408889634f;

typedef unsigned long long u64;

struct SinkArgs {
  const float* Z; int ld;
  int n0_max, n1_max;
  const int* n_dev;
  int iters;
  float* u; float* v;              // outputs: final potentials
  u64* part;                       // [grid][ldp]  (m, +-s): per-CTA column partials, sign(s) = iteration tag
  u64* vx;                         // [ldp]        (v_j, iteration+1)
  u64* cpart;                      // [grid][ldp]  (best value, row+1): column argmax partials of the last pass
  int ldp;
  unsigned* err;                   // set if a poll timed out (never expected; the results are then poisoned)
  int* idx0; int* idx1; float* max0; float* max1;
  int rpc_max;                     // ceil((n0_max+1)/grid): capacity of the per-CTA u buffer
  int slab_rows;                   // rows of the smem slab (0 = stream from global)
  int slab_ld;                     // padded row length of the slab (multiple of 4)
};

__device__ __forceinline__ u64 pack2(float a, float b) {
  return (u64)__float_as_uint(a) | ((u64)__float_as_uint(b) << 32);
}
__device__ __forceinline__ void st_relaxed(u64* p, u64 v) {
  asm volatile("st.relaxed.gpu.global.b64 [%0], %1;" ::"l"(p), "l"(v) : "memory");
}
__device__ __forceinline__ u64 ld_relaxed(const u64* p) {
  u64 v;
  asm volatile("ld.relaxed.gpu.global.b64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// poll one tagged word until `ready(word)`; bounded so that a lost producer cannot hang the GPU
template <typename Pred>
__device__ __forceinline__ u64 poll(const u64* p, unsigned* err, Pred ready) {
  u64 w = ld_relaxed(p);
  if (ready(w)) return w;
  long long t0 = clock64();
  unsigned spins = 0;
  while (true) {
    w = ld_relaxed(p);
    if (ready(w)) return w;
    if ((++spins & 1023u) == 0) {
      if (*(volatile unsigned*)err) return w;
      if (clock64() - t0 > 400000000LL) { atomicExch(err, 1u); return w; }
    }
  }
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
// (m, s) <- combine of two (max, sum-of-exp) partials
__device__ __forceinline__ void lse_merge(float& m, float& s, float pm, float ps) {
  float mn = fmaxf(m, pm);
  s = s * exp2f((m - mn) * kLog2e) + ps * exp2f((pm - mn) * kLog2e);
  m = mn;
}

__global__ void __launch_bounds__(kThreads, 1) k_sinkhorn(SinkArgs a) {
  extern __shared__ __align__(16) float smem[];
  __shared__ float red_m[32][17], red_s[32][17];
  const int n0 = a.n_dev ? min(a.n_dev[0], a.n0_max) : a.n0_max;
  const int n1 = a.n_dev ? min(a.n_dev[1], a.n1_max) : a.n1_max;
  const int R = n0 + 1, C = n1 + 1;
  const int G = gridDim.x, b = blockIdx.x;
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;

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

  const int vlen = (a.n1_max + 1 + 3) & ~3;
  float* v_s = smem;                                // [vlen], entries >= C stay 0
  float* u_s = v_s + vlen;                          // [rpc_max]
  float* slab = u_s + ((a.rpc_max + 3) & ~3);
  const bool resident = a.slab_rows > 0 && nrows <= a.slab_rows && C <= a.slab_ld;
  const int zs = resident ? a.slab_ld : a.ld;       // row stride used by the passes
  const float* zbase = resident ? slab : a.Z + (size_t)r_begin * a.ld;
  const int C4 = resident ? ((C + 3) & ~3) : C;     // resident rows are padded with -inf up to a multiple of 4

  if (resident) {
    for (int r = warp; r < nrows; r += kWarps) {
      const float* src = a.Z + (size_t)(r_begin + r) * a.ld;
      float* dst = slab + (size_t)r * a.slab_ld;
      for (int j = lane; j < C4; j += 32) dst[j] = (j < C) ? src[j] : -CUDART_INF_F;
    }
  }
  for (int j = tid; j < vlen; j += kThreads) v_s[j] = 0.f;
  if (a.iters == 0) for (int j = c_begin + tid; j < c_end; j += kThreads) a.v[j] = 0.f;
  __syncthreads();

  for (int it = 0; it < a.iters; ++it) {
    // ---- row pass -------------------------------------------------------------------------
    for (int r = warp; r < nrows; r += kWarps) {
      const float* z = zbase + (size_t)r * zs;
      float m = -CUDART_INF_F, s = 0.f;
      if (resident) {
        const float4* z4 = reinterpret_cast<const float4*>(z);
        const float4* v4 = reinterpret_cast<const float4*>(v_s);
        const int n4 = C4 >> 2;
#pragma unroll 4
        for (int i = lane; i < n4; i += 32) {
          float4 zz = z4[i], vv = v4[i];
          m = fmaxf(m, fmaxf(fmaxf(zz.x + vv.x, zz.y + vv.y), fmaxf(zz.z + vv.z, zz.w + vv.w)));
        }
        m = warp_max(m);
#pragma unroll 4
        for (int i = lane; i < n4; i += 32) {
          float4 zz = z4[i], vv = v4[i];
          s += exp2f((zz.x + vv.x - m) * kLog2e) + exp2f((zz.y + vv.y - m) * kLog2e) +
               exp2f((zz.z + vv.z - m) * kLog2e) + exp2f((zz.w + vv.w - m) * kLog2e);
        }
      } else {
#pragma unroll 4
        for (int j = lane; j < C; j += 32) m = fmaxf(m, z[j] + v_s[j]);
        m = warp_max(m);
#pragma unroll 4
        for (int j = lane; j < C; j += 32) s += exp2f((z[j] + v_s[j] - m) * kLog2e);
      }
      s = warp_sum(s);
      if (lane == 0) {
        float lmu = (r_begin + r == n0) ? log_mu_last : norm;
        u_s[r] = lmu - (m + logf(s));
      }
    }
    __syncthreads();
    // ---- column pass: per-CTA partial LSE, published with the iteration tag in sign(s) --------------
    const unsigned tagbit = ((unsigned)(it + 1) & 1u) << 31;
    if (b < Ga) {
      if (nrows <= kRMax) {
        for (int j = tid; j < C; j += kThreads) {
          float x[kRMax];
          float m = -CUDART_INF_F;
#pragma unroll
          for (int r = 0; r < kRMax; ++r) {
            x[r] = (r < nrows) ? zbase[(size_t)r * zs + j] + u_s[r] : -CUDART_INF_F;
            m = fmaxf(m, x[r]);
          }
          float s = 0.f;
#pragma unroll
          for (int r = 0; r < kRMax; ++r) s += exp2f((x[r] - m) * kLog2e);
          st_relaxed(&a.part[(size_t)b * a.ldp + j], pack2(m, __uint_as_float(__float_as_uint(s) | tagbit)));
        }
      } else {
        for (int j = tid; j < C; j += kThreads) {
          float m = -CUDART_INF_F;
          for (int r = 0; r < nrows; ++r) m = fmaxf(m, zbase[(size_t)r * zs + j] + u_s[r]);
          float s = 0.f;
          for (int r = 0; r < nrows; ++r) s += exp2f((zbase[(size_t)r * zs + j] + u_s[r] - m) * kLog2e);
          st_relaxed(&a.part[(size_t)b * a.ldp + j], pack2(m, __uint_as_float(__float_as_uint(s) | tagbit)));
        }
      }
    }
    // ---- combine the partials of my column range -> v_j, published with the iteration number --------
    const unsigned epoch = (unsigned)(it + 1);
    for (int jc = c_begin; jc < c_end; jc += 16) {
      const int jj = tid & 15, gs = tid >> 4, j = jc + jj;
      float m = -CUDART_INF_F, s = 0.f;
      if (j < c_end) {
        for (int g = gs; g < Ga; g += 32) {
          u64 w = poll(&a.part[(size_t)g * a.ldp + j], a.err,
                       [&](u64 x) { return (((unsigned)(x >> 32)) & 0x80000000u) == tagbit; });
          float pm = __uint_as_float((unsigned)w), ps = __uint_as_float(((unsigned)(w >> 32)) & 0x7fffffffu);
          lse_merge(m, s, pm, ps);
        }
      }
      red_m[gs][jj] = m; red_s[gs][jj] = s;
      __syncthreads();
      const int jw = jc + warp;                      // warp w finishes column jc + w
      if (jw < c_end) {
        float mm = red_m[lane][warp], ss = red_s[lane][warp];
        float M = warp_max(mm);
        ss = (ss > 0.f) ? ss * exp2f((mm - M) * kLog2e) : 0.f;
        ss = warp_sum(ss);
        if (lane == 0) {
          float lnu = (jw == n1) ? log_nu_last : norm;
          float vj = lnu - (M + logf(ss));
          st_relaxed(&a.vx[jw], pack2(vj, __uint_as_float(epoch)));
          if (it == a.iters - 1) a.v[jw] = vj;
        }
      }
      __syncthreads();
    }
    // ---- gather the new v (poll the tagged words) -----------------------------------------------------
    for (int j = tid; j < C; j += kThreads) {
      u64 w = poll(&a.vx[j], a.err, [&](u64 x) { return (unsigned)(x >> 32) == epoch; });
      v_s[j] = __uint_as_float((unsigned)w);
    }
    __syncthreads();
  }

  // ---- final pass: Z = ((Z0 + u) + v) - norm, row / column max + first argmax -------------------
  for (int r = warp; r < nrows; r += kWarps) {
    int gr = r_begin + r;
    float ur = (a.iters > 0) ? u_s[r] : 0.f;
    if (lane == 0) a.u[gr] = ur;
    if (gr >= n0) continue;
    const float* z = zbase + (size_t)r * zs;
    float best = -CUDART_INF_F;
    int bj = 0x7fffffff;
#pragma unroll 4
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
    int live = min(r_end, n0) - r_begin;           // rows of mine that are real keypoints (may be <= 0)
    for (int j = tid; j < n1; j += kThreads) {
      float best = -CUDART_INF_F;
      int bi = 0x7ffffffe;
      float vj = v_s[j];
      for (int r = 0; r < live; ++r) {
        float ur = (a.iters > 0) ? u_s[r] : 0.f;
        float t = ((zbase[(size_t)r * zs + j] + ur) + vj) - norm;
        if (t > best) { best = t; bi = r_begin + r; }
      }
      st_relaxed(&a.cpart[(size_t)b * a.ldp + j], pack2(best, __int_as_float(bi + 1)));   // row+1 != 0: "written"
    }
  }
  for (int jc = c_begin; jc < c_end; jc += 16) {
    const int jj = tid & 15, gs = tid >> 4, j = jc + jj;
    float best = -CUDART_INF_F;
    int bi = 0x7fffffff;
    if (j < c_end && j < n1) {
      for (int g = gs; g < Ga; g += 32) {
        u64 w = poll(&a.cpart[(size_t)g * a.ldp + j], a.err, [&](u64 x) { return (unsigned)(x >> 32) != 0u; });
        float pv = __uint_as_float((unsigned)w);
        int pi = (int)(unsigned)(w >> 32) - 1;
        if (pv > best || (pv == best && pi < bi)) { best = pv; bi = pi; }
      }
    }
    red_m[gs][jj] = best; red_s[gs][jj] = __int_as_float(bi);
    __syncthreads();
    const int jw = jc + warp;
    if (jw < c_end && jw < n1) {
      float bb = red_m[lane][warp];
      int ii = __float_as_int(red_s[lane][warp]);
#pragma unroll
      for (int o = 16; o; o >>= 1) {
        float ob = __shfl_xor_sync(0xffffffffu, bb, o);
        int oi = __shfl_xor_sync(0xffffffffu, ii, o);
        if (ob > bb || (ob == bb && oi < ii)) { bb = ob; ii = oi; }
      }
      if (lane == 0) { a.idx1[jw] = ii; a.max1[jw] = bb; }
    }
    __syncthreads();
  }
}

// gmatcher.py:286-294
__global__ void k_match_finalize(int n0_max, int n1_max, const int* __restrict__ n_dev, const int* __restrict__ idx0,
                                 const int* __restrict__ idx1, const float* __restrict__ max0, float thr,
                                 int64_t* __restrict__ matches0, int64_t* __restrict__ matches1,
                                 float* __restrict__ ms0, float* __restrict__ ms1, const unsigned* __restrict__ err) {
  int n0 = n_dev ? min(n_dev[0], n0_max) : n0_max;
  int n1 = n_dev ? min(n_dev[1], n1_max) : n1_max;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (*err) {                                     // a poll timed out inside k_sinkhorn: fail loudly, not silently
    if (t < n0) { ms0[t] = CUDART_NAN_F; matches0[t] = -2; }
    if (t < n1) { ms1[t] = CUDART_NAN_F; matches1[t] = -2; }
    return;
  }
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
  // zeroed before every launch (tags)
  unsigned* err;
  u64* vx;
  u64* part;
  u64* cpart;
  size_t zero_bytes;
  float* max0;
  float* max1;
};

size_t carve(SinkWs& w, void* base, size_t cap, int n0_max, int n1_max, int grid) {
  Arena a(base, cap);
  size_t ldp = (size_t)n1_max + 1;
  w.err = a.take<unsigned>(64);
  w.vx = a.take<u64>(ldp);
  w.part = a.take<u64>((size_t)grid * ldp);
  w.cpart = a.take<u64>((size_t)grid * ldp);
  w.zero_bytes = align_up(a.off, 256);
  w.max0 = a.take<float>(n0_max + 1);
  w.max1 = a.take<float>(n1_max + 1);
  return align_up(a.off, 256);
}

}  // namespace

}  // namespace gims

using namespace gims;

extern "C" size_t gims_sinkhorn_workspace_bytes(int n0_max, int n1_max) {
  SinkWs w;
  return carve(w, nullptr, 0, n0_max, n1_max, kMaxGrid);
}

extern "C" int gims_sinkhorn_match(const float* couplings, int n0_max, int n1_max, const int* n_dev, int iters,
                                   float match_threshold, void* workspace, size_t workspace_bytes, float* u, float* v,
                                   int* indices0, int* indices1, int64_t* matches0, int64_t* matches1, float* mscores0,
                                   float* mscores1, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (n0_max < 1 || n1_max < 1 || iters < 0) { set_error("gims_sinkhorn_match: bad sizes"); return GIMS_ERR_ARG; }
  int dev = 0, sms = 0, smem_optin = 0;
  GIMS_CUDA_OK(cudaGetDevice(&dev));
  GIMS_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  GIMS_CUDA_OK(cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
  int G = sms < kMaxGrid ? sms : kMaxGrid;
  SinkWs w;
  size_t need = carve(w, workspace, workspace_bytes, n0_max, n1_max, G);
  if (need > workspace_bytes) { set_error("gims_sinkhorn_match: workspace %zu < %zu", workspace_bytes, need); return GIMS_ERR_WORKSPACE; }
  int R = n0_max + 1, C = n1_max + 1;
  int rpc = (R + G - 1) / G;
  int slab_ld = (C + 3) & ~3;
  size_t fixed = (size_t)(((C + 3) & ~3) + ((rpc + 3) & ~3)) * sizeof(float);
  size_t slab_bytes = (size_t)rpc * slab_ld * sizeof(float);
  size_t budget = (size_t)smem_optin - 5120;     // static smem (red_m / red_s) + margin
  SinkArgs a;
  a.rpc_max = rpc;
  a.slab_rows = (fixed + slab_bytes <= budget) ? rpc : 0;
  a.slab_ld = slab_ld;
  size_t dyn = fixed + (a.slab_rows ? slab_bytes : 0);
  if (dyn > budget) { set_error("gims_sinkhorn_match: n1_max=%d needs %zu B of shared memory", n1_max, dyn); return GIMS_ERR_ARG; }
  a.Z = couplings; a.ld = C; a.n0_max = n0_max; a.n1_max = n1_max; a.n_dev = n_dev; a.iters = iters;
  a.u = u; a.v = v; a.part = w.part; a.vx = w.vx; a.cpart = w.cpart; a.ldp = C; a.err = w.err;
  a.idx0 = indices0; a.idx1 = indices1; a.max0 = w.max0; a.max1 = w.max1;
  GIMS_CUDA_OK(cudaFuncSetAttribute(k_sinkhorn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)dyn));
  int per_sm = 0;
  GIMS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sinkhorn, kThreads, dyn));
  if (per_sm < 1) { set_error("gims_sinkhorn_match: kernel does not fit an SM (dyn smem %zu)", dyn); return GIMS_ERR_ARG; }
  GIMS_CUDA_OK(cudaMemsetAsync(w.err, 0, w.zero_bytes, st));    // clears every tag (err, vx, part, cpart are contiguous)
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
                                                 matches0, matches1, mscores0, mscores1, w.err);
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}
