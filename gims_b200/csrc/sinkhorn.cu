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
constexpr int kRowChunks = 17;       // float4 chunks per lane of the register row pass (rows up to 2176 columns)
constexpr float kLog2e = 1.4426950408889634f;
// Bound of every poll in clock64 cycles (~2 s at 1.9 GHz): far beyond any legitimate wait, also under ncu replay, MPS
// time-slicing or a debugger; on expiry the launch poisons its outputs and raises GIMS_STATUS_SINKHORN_TIMEOUT.
constexpr long long kPollTimeoutClk = 4000000000LL;
constexpr int kWays = 2;             // copies of the column-sum buffer (CTA b adds into copy b % kWays): the red.adds of
                                     // 148 CTAs on one 128-byte line serialize in L2, four copies cut that chain by four

typedef unsigned long long u64;

struct SinkArgs {
  const float* Z; int ld;
  int n0_max, n1_max;
  const int* n_dev;
  int iters;
  float* u; float* v;              // outputs: final potentials
  float2* part;                    // [grid][ldp]  (m, s): per-CTA column partials; last pass: (best value, row index bits)
  float* vg;                       // [ldp + 4]    v_j exchange buffer
  unsigned* pflag;                 // [grid]       iteration number of the partials CTA g has published
  unsigned* vflag;                 // [grid]       iteration number of the v_j CTA g has published
  float* colsum;                   // [3][kWays][ldp] scaled-kernel path: column sums accumulated with red.add, 3 rotating buffers
  unsigned* counter;               // grid-wide arrival counter of the scaled-kernel path
  u64* mm;                         // [2*grid]     (zmin | 1<<32), (zmax | 1<<32) of every CTA's slab
  int ldp;
  unsigned* err;                   // set if a poll timed out (never expected; the results are then poisoned)
  int* idx0; int* idx1; float* max0; float* max1;
  int rpc_max;                     // ceil((n0_max+1)/grid): capacity of the per-CTA u buffer
  int slab_rows;                   // rows of the smem slab (0 = stream from global)
  int slab_ld;                     // padded row length of the slab (multiple of 4)
};

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
      if (clock64() - t0 > kPollTimeoutClk) { atomicExch(err, 1u); return w; }
    }
  }
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
// Publish: every thread of the CTA has written its data with plain stores; one fence + one flag store makes them
// visible (the block barrier orders the other threads' stores before thread 0's fence — cumulativity).
__device__ __forceinline__ void publish_flag(unsigned* flag, unsigned epoch) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(flag), "r"(epoch) : "memory");
  }
}
// Warp 0 waits until flags[0..n) have all reached `epoch`, then the whole CTA proceeds (data is then read with
// ld.global.cg, which cannot hit a stale L1 line).  Bounded like every other wait in this file.
__device__ __forceinline__ void wait_flags(const unsigned* flags, int n, unsigned epoch, unsigned* err) {
  if (threadIdx.x < 32) {
    const int lane = threadIdx.x;
    long long t0 = clock64();
    unsigned spins = 0;
    while (true) {
      bool ok = true;
      for (int g = lane; g < n; g += 32) ok = ok && (ld_acquire_u32(&flags[g]) >= epoch);
      if (__all_sync(0xffffffffu, ok)) break;
      if ((++spins & 63u) == 0) {
        bool bail = *(volatile unsigned*)err != 0u;
        if (clock64() - t0 > kPollTimeoutClk) { atomicExch(err, 1u); bail = true; }
        if (__any_sync(0xffffffffu, bail)) break;
      }
    }
  }
  __syncthreads();
}

// Grid-wide arrive + wait on one monotonically increasing counter (every CTA is co-resident: cooperative launch).
// Measured on B200 (tools/micro/hop_latency.cu): 2.5 k cycles per hop, about half of a per-CTA flag scheme.
__device__ __forceinline__ void grid_hop(unsigned* counter, unsigned target, unsigned* err) {
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(counter, 1u);
    long long t0 = clock64();
    unsigned spins = 0;
    while (ld_acquire_u32(counter) < target) {
      if ((++spins & 255u) == 0) {
        if (*(volatile unsigned*)err) break;
        if (clock64() - t0 > kPollTimeoutClk) { atomicExch(err, 1u); break; }
      }
    }
  }
  __syncthreads();
}

// Up to N tagged words: issue every load first (one L2 round trip for all of them), then poll only the stragglers.
template <int N, typename Addr, typename Pred>
__device__ __forceinline__ void poll_batch(u64 (&w)[N], int count, Addr addr, unsigned* err, Pred ready) {
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (k < count) w[k] = ld_relaxed(addr(k));
#pragma unroll
  for (int k = 0; k < N; ++k)
    if (k < count && !ready(w[k])) w[k] = poll(addr(k), err, ready);
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
// 2^x for x <= 0: one MUFU (results below 2^-126 flush to zero, which is what a log-sum-exp term wants)
__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
// (m, s) <- combine of two (max, sum-of-exp) partials
__device__ __forceinline__ void lse_merge(float& m, float& s, float pm, float ps) {
  float mn = fmaxf(m, pm);
  s = s * ex2((m - mn) * kLog2e) + ps * ex2((pm - mn) * kLog2e);
  m = mn;
}

// Optional timeline (profiling): CTA 0 / thread 0 stores clock64() at 6 points of each of the first 16 iterations.
__device__ long long* g_sink_trace = nullptr;
#define SINK_TRACE(slot)                                                        \
  do {                                                                          \
    if (trace && it < 16) trace[it * 8 + (slot)] = clock64();                   \
    if (gtrace && it == 5) {                                                    \
      unsigned long long gt_;                                                   \
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt_));                   \
      gtrace[128 + blockIdx.x * 8 + (slot)] = (long long)gt_;                   \
    }                                                                           \
  } while (0)

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
  const int rlen = (a.rpc_max + 3) & ~3;
  float* v_s = smem;                                // [vlen], entries >= C stay 0
  float* w_s = v_s + vlen;                          // [vlen]  exp(v_j - vmax) (scaled-kernel path), pads 0
  float* u_s = w_s + vlen;                          // [rlen]  u_r
  float* ut_s = u_s + rlen;                         // [rlen]  u_r + rowmax_r
  float* rmax_s = ut_s + rlen;                      // [rlen]  row maxima of my slab
  float* e_s = rmax_s + rlen;                       // [rlen]  exp(ut_r - max ut)
  float* slab = e_s + rlen;
  __shared__ float red_x[kWarps], red_y[kWarps];
  const bool resident = a.slab_rows > 0 && nrows <= a.slab_rows && C <= a.slab_ld;
  const int zs = resident ? a.slab_ld : a.ld;       // row stride used by the passes
  const float* zbase = resident ? slab : a.Z + (size_t)r_begin * a.ld;
  const int C4 = resident ? ((C + 3) & ~3) : C;     // resident rows are padded with -inf up to a multiple of 4

  float zmin_w = CUDART_INF_F, zmax_w = -CUDART_INF_F;
  if (resident) {
    for (int r = warp; r < nrows; r += kWarps) {
      const float* src = a.Z + (size_t)(r_begin + r) * a.ld;
      float* dst = slab + (size_t)r * a.slab_ld;
      float mx = -CUDART_INF_F, mn = CUDART_INF_F;
      for (int j = lane; j < C4; j += 32) {
        float z = (j < C) ? src[j] : -CUDART_INF_F;
        dst[j] = z;
        if (j < C) { mx = fmaxf(mx, z); mn = fminf(mn, z); }
      }
      mx = warp_max(mx);
      mn = -warp_max(-mn);
      if (lane == 0) rmax_s[r] = mx;
      zmax_w = fmaxf(zmax_w, mx); zmin_w = fminf(zmin_w, mn);
    }
  }
  for (int j = tid; j < vlen; j += kThreads) { v_s[j] = 0.f; w_s[j] = (j < C) ? 1.f : 0.f; }
  if (a.iters == 0) for (int j = c_begin + tid; j < c_end; j += kThreads) a.v[j] = 0.f;
  if (lane == 0) { red_x[warp] = zmax_w; red_y[warp] = zmin_w; }
  __syncthreads();
  // ---- scaled-kernel path: decided once, identically by every CTA, from the global range of the couplings.
  // With R = max Z0 - min Z0 every potential stays within R + 2 log(N+M) of any other (u_r differs between rows by at
  // most R plus the log-marginals, likewise v_j), also from one iteration to the next.  For R <= 20 the weights
  // exp(pot - reference) therefore stay inside e^+-70 and every row / column sum contains a term >= e^-70:
  // nothing overflows, and anything that underflows is < 1e-7 of its sum.  The iteration then needs no exp per
  // matrix element at all: slab = exp(z - rowmax) once, row pass = sum_j E_rj w_j, column pass = sum_r E_rj e_r.
  bool fast = false;
  if (resident && nrows <= kRMax && (C4 >> 2) <= kRowChunks * 32 && a.iters > 0) {
    if (tid == 0 && b < Ga) {
      float zx = red_x[0], zn = red_y[0];
      for (int w = 1; w < kWarps; ++w) { zx = fmaxf(zx, red_x[w]); zn = fminf(zn, red_y[w]); }
      st_relaxed(&a.mm[2 * b], (u64)__float_as_uint(zn) | (1ull << 32));
      st_relaxed(&a.mm[2 * b + 1], (u64)__float_as_uint(zx) | (1ull << 32));
    }
    float gx = -CUDART_INF_F, gn = CUDART_INF_F;
    if (warp == 0) {
      for (int g = lane; g < Ga; g += 32) {
        u64 w0 = poll(&a.mm[2 * g], a.err, [](u64 x) { return (x >> 32) != 0ull; });
        u64 w1 = poll(&a.mm[2 * g + 1], a.err, [](u64 x) { return (x >> 32) != 0ull; });
        gn = fminf(gn, __uint_as_float((unsigned)w0));
        gx = fmaxf(gx, __uint_as_float((unsigned)w1));
      }
      gx = warp_max(gx);
      gn = -warp_max(-gn);
      if (lane == 0) red_x[0] = ((gx - gn) <= 20.f) ? 1.f : 0.f;
    }
    __syncthreads();
    fast = red_x[0] != 0.f;
    __syncthreads();
    if (fast) {                                     // slab <- E = exp(z - rowmax), pads 0
      for (int r = warp; r < nrows; r += kWarps) {
        float* row = slab + (size_t)r * a.slab_ld;
        const float rm = rmax_s[r];
        for (int j = lane; j < C4; j += 32) row[j] = (j < C) ? ex2((row[j] - rm) * kLog2e) : 0.f;
      }
      __syncthreads();
    }
  }

  if (b == 0 && tid == 0) a.err[1] = fast ? GIMS_STATUS_SINKHORN_FAST : GIMS_STATUS_SINKHORN_EXACT;
  // the trace pointer is read once: a load of the global per stamp would sit on every iteration's critical path
  long long* const gtrace = (tid == 0) ? g_sink_trace : nullptr;
  long long* trace = (b == 0) ? gtrace : nullptr;
  const bool reg_rows = resident && (C4 >> 2) <= kRowChunks * 32;   // a row's (z + v) fits the lanes' registers
  const bool reg_cols = resident && nrows <= kRMax;
  if (fast) {
    // ===== scaled-kernel iteration: ONE grid-wide exchange per iteration =================================================
    //   row pass   sr_r = sum_j E_rj w_j           -> u_r, and the column weight e_r = exp(ut_r - Rref)
    //   col pass   s_j  = sum_{my rows} E_rj e_r   -> red.add into colsum[it%3][j]   (all CTAs share the reference Rref,
    //                                                 so partial sums add up directly; fp32 atomics: order-dependent
    //                                                 rounding, ~1e-7 relative)
    //   hop        arrive/wait on the counter
    //   gather     v_j = log_nu_j - (Rref + log colsum_j) computed by every CTA for every column; w_j = exp(v_j - v_0)
    const int n4 = C4 >> 2;
    const int Gf = C >> 2;
    const size_t cs_ld = ((size_t)a.ldp + 3) & ~(size_t)3;
    float vref = 0.f;                               // reference the current w_s was scaled with (initially v = 0, w = 1)
    float v0_prev = 0.f;                            // v_0 after the previous iteration (initially v = 0)
    unsigned target = 0;
    if (Gf <= kThreads) {
      // ----- register-resident variant (C <= 4 * kThreads + 3, i.e. up to 2048 keypoints in image 1): thread t owns the
      // float4 column group t of every slab row in registers (<= 16 rows x 4 floats) together with the group's weights,
      // so neither pass reads the slab from shared memory: the row pass is 64 FMAs + a block reduction of <= 16 row
      // sums, the column pass 64 FMAs whose result is already this CTA's complete partial for those columns.  The
      // (<= 3) columns past the last full group, among them the dustbin column, still come from the shared-memory slab.
      const bool own = tid < Gf;
      float4 ereg[kRMax];
#pragma unroll
      for (int r = 0; r < kRMax; ++r)
        ereg[r] = (own && r < nrows) ? *reinterpret_cast<const float4*>(slab + (size_t)r * a.slab_ld + 4 * tid)
                                     : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 wreg = make_float4(1.f, 1.f, 1.f, 1.f);
      for (int it = 0; it < a.iters; ++it) {
        SINK_TRACE(0);
        float* cs = a.colsum + (size_t)(it % 3) * kWays * cs_ld;   // [kWays][cs_ld]
        float* my = cs + (size_t)(b % kWays) * cs_ld;
        const float Rref = norm - vref;
        float part[kRMax];
#pragma unroll
        for (int r = 0; r < kRMax; ++r)
          part[r] = fmaf(ereg[r].x, wreg.x, ereg[r].y * wreg.y) + fmaf(ereg[r].z, wreg.z, ereg[r].w * wreg.w);
        // 16 sums over 32 lanes with 16 shuffles (a butterfly that halves the number of live values per step) instead
        // of 16 x 5: shuffles issue at one warp instruction per clock per SM and were the longest part of this pass
        static_assert(kRMax == 16, "the butterfly below is written for 16 row sums");
#pragma unroll
        for (int h = 8, bit = 16; h >= 1; h >>= 1, bit >>= 1) {
          const bool up = (lane & bit) != 0;
#pragma unroll
          for (int i = 0; i < h; ++i) {
            const float send = up ? part[i] : part[i + h];
            const float keep = up ? part[i + h] : part[i];
            part[i] = keep + __shfl_xor_sync(0xffffffffu, send, bit);
          }
        }
        part[0] += __shfl_xor_sync(0xffffffffu, part[0], 1);
        if ((lane & 1) == 0) red_m[warp][lane >> 1] = part[0];      // lane 2r holds the warp's sum of row r
        __syncthreads();
        SINK_TRACE(7);
        if (tid < nrows) {                              // thread r finishes row r
          float sr = 0.f;
#pragma unroll
          for (int w = 0; w < kWarps; ++w) sr += red_m[w][tid];
          for (int j = 4 * Gf; j < C; ++j) sr = fmaf(slab[(size_t)tid * a.slab_ld + j], w_s[j], sr);
          const float lmu = (r_begin + tid == n0) ? log_mu_last : norm;
          const float lse_rel = vref + logf(sr);               // LSE_j(z + v) - rowmax
          e_s[tid] = ex2(((lmu - lse_rel) - Rref) * kLog2e);    // exp(u_r + rowmax_r - Rref)
          u_s[tid] = lmu - (rmax_s[tid] + lse_rel);
        }
        __syncthreads();
        SINK_TRACE(1);
        if (b < Ga) {
          if (own) {
            float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
            for (int r = 0; r < kRMax; ++r) {
              const float er = (r < nrows) ? e_s[r] : 0.f;
              sm.x = fmaf(ereg[r].x, er, sm.x); sm.y = fmaf(ereg[r].y, er, sm.y);
              sm.z = fmaf(ereg[r].z, er, sm.z); sm.w = fmaf(ereg[r].w, er, sm.w);
            }
            // (packed FFMA2 for these passes and a 16-lanes-per-row finish were tried: both slower on B200)
            SINK_TRACE(6);
            asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(my + 4 * tid), "f"(sm.x), "f"(sm.y),
                         "f"(sm.z), "f"(sm.w)
                         : "memory");
          }
          const int jl = 4 * Gf + warp;                 // the (at most 3) left-over columns
          if (jl < C) {
            const float xx = (lane < nrows) ? slab[(size_t)lane * a.slab_ld + jl] * e_s[lane] : 0.f;
            const float ss = warp_sum(xx);
            if (lane == 0) atomicAdd(my + jl, ss);
          }
        }
        SINK_TRACE(2);
        target += (unsigned)G;
        grid_hop(a.counter, target, a.err);
        SINK_TRACE(3);
        {                                               // recycle the buffer that was read one iteration ago
          float* old = a.colsum + (size_t)((it + 2) % 3) * kWays * cs_ld;
          for (int w = 0; w < kWays; ++w)
            for (int j = c_begin + tid; j < c_end; j += kThreads) old[(size_t)w * cs_ld + j] = 0.f;
        }
        // Reference of the new weights: v_0 of the PREVIOUS iteration — every thread already has it, and any reference
        // inside the (bounded) range of v works.  (Deriving it from this iteration's colsum[0] made every warp of the
        // grid load the same address right after the hop: 2400 serialized requests on one L2 sector per iteration.)
        const float v0 = v0_prev;
        // n4 <= kThreads + 1 here: a thread has its own group and at most one more (the left-over columns); both loads
        // are issued before either is used — one L2 round trip, not two, on the thread every other thread waits for
        const int g4b = tid + kThreads;
        auto load_cs = [&](int g4) {
          float4 c4 = __ldcg(reinterpret_cast<const float4*>(cs) + g4);
#pragma unroll
          for (int w = 1; w < kWays; ++w) {
            const float4 cw = __ldcg(reinterpret_cast<const float4*>(cs + (size_t)w * cs_ld) + g4);
            c4.x += cw.x; c4.y += cw.y; c4.z += cw.z; c4.w += cw.w;
          }
          return c4;
        };
        auto finish = [&](int g4, const float4 c4) {
          const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
          float wn[4];
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int j = 4 * g4 + q;
            wn[q] = 0.f;
            if (j < C) {
              const float lnu = (j == n1) ? log_nu_last : norm;
              const float vj = lnu - (Rref + logf(cc[q]));
              wn[q] = ex2((vj - v0) * kLog2e);
              v_s[j] = vj;
              w_s[j] = wn[q];
              if (it == a.iters - 1 && j >= c_begin && j < c_end) a.v[j] = vj;
            }
          }
          if (g4 == tid) wreg = make_float4(wn[0], wn[1], wn[2], wn[3]);
        };
        float4 ca = make_float4(1.f, 1.f, 1.f, 1.f), cb = ca;
        if (tid < n4) ca = load_cs(tid);
        if (g4b < n4) cb = load_cs(g4b);
        if (tid < n4) finish(tid, ca);
        if (g4b < n4) finish(g4b, cb);
        SINK_TRACE(4);
        vref = v0;
        __syncthreads();
        v0_prev = v_s[0];
        SINK_TRACE(5);
      }
    } else
    for (int it = 0; it < a.iters; ++it) {
      SINK_TRACE(0);
      float* cs = a.colsum + (size_t)(it % 3) * kWays * cs_ld;   // [kWays][cs_ld]
        float* my = cs + (size_t)(b % kWays) * cs_ld;
      const float Rref = norm - vref;
      for (int r = warp; r < nrows; r += kWarps) {
        const float4* e4 = reinterpret_cast<const float4*>(slab + (size_t)r * a.slab_ld);
        const float4* w4 = reinterpret_cast<const float4*>(w_s);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int k = 0; k < kRowChunks; ++k) {
          const int i = lane + 32 * k;
          if (i < n4) {
            const float4 ee = e4[i], ww = w4[i];
            s0 = fmaf(ee.x, ww.x, s0); s1 = fmaf(ee.y, ww.y, s1); s2 = fmaf(ee.z, ww.z, s2); s3 = fmaf(ee.w, ww.w, s3);
          }
        }
        const float sr = warp_sum((s0 + s1) + (s2 + s3));
        if (lane == 0) {
          const float lmu = (r_begin + r == n0) ? log_mu_last : norm;
          const float lse_rel = vref + logf(sr);               // LSE_j(z + v) - rowmax
          e_s[r] = ex2(((lmu - lse_rel) - Rref) * kLog2e);      // exp(u_r + rowmax_r - Rref)
          u_s[r] = lmu - (rmax_s[r] + lse_rel);
        }
      }
      __syncthreads();
      SINK_TRACE(1);
      if (b < Ga) {
        for (int g = tid; g < Gf; g += kThreads) {
          float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int r = 0; r < kRMax; ++r) {
            if (r < nrows) {
              const float4 ee = *reinterpret_cast<const float4*>(slab + (size_t)r * a.slab_ld + 4 * g);
              const float er = e_s[r];
              sm.x = fmaf(ee.x, er, sm.x); sm.y = fmaf(ee.y, er, sm.y); sm.z = fmaf(ee.z, er, sm.z); sm.w = fmaf(ee.w, er, sm.w);
            }
          }
          asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(my + 4 * g), "f"(sm.x), "f"(sm.y), "f"(sm.z),
                       "f"(sm.w)
                       : "memory");
        }
        const int jl = 4 * Gf + warp;                 // the (at most 3) left-over columns, among them the dustbin column
        if (jl < C) {
          const float xx = (lane < nrows) ? slab[(size_t)lane * a.slab_ld + jl] * e_s[lane] : 0.f;
          const float ss = warp_sum(xx);
          if (lane == 0) atomicAdd(my + jl, ss);
        }
      }
      SINK_TRACE(2);
      target += (unsigned)G;
      grid_hop(a.counter, target, a.err);
      SINK_TRACE(3);
      {                                               // recycle the buffer that was read one iteration ago
        float* old = a.colsum + (size_t)((it + 2) % 3) * kWays * cs_ld;
        for (int w = 0; w < kWays; ++w)
            for (int j = c_begin + tid; j < c_end; j += kThreads) old[(size_t)w * cs_ld + j] = 0.f;
      }
      const float v0 = v0_prev;                                   // last iteration's v_0 (see the register variant)
      for (int g4 = tid; g4 < n4; g4 += kThreads) {               // 128-bit loads: 4x fewer requests on these hot lines
        float4 c4 = __ldcg(reinterpret_cast<const float4*>(cs) + g4);
#pragma unroll
          for (int w = 1; w < kWays; ++w) {
            const float4 cw = __ldcg(reinterpret_cast<const float4*>(cs + (size_t)w * cs_ld) + g4);
            c4.x += cw.x; c4.y += cw.y; c4.z += cw.z; c4.w += cw.w;
          }
        const float cc[4] = {c4.x, c4.y, c4.z, c4.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          const int j = 4 * g4 + q;
          if (j < C) {
            const float lnu = (j == n1) ? log_nu_last : norm;
            const float vj = lnu - (Rref + logf(cc[q]));
            v_s[j] = vj;
            w_s[j] = ex2((vj - v0) * kLog2e);
            if (it == a.iters - 1 && j >= c_begin && j < c_end) a.v[j] = vj;
          }
        }
      }
      vref = v0;
      __syncthreads();
      v0_prev = v_s[0];
      SINK_TRACE(5);
    }
  } else
  for (int it = 0; it < a.iters; ++it) {
    SINK_TRACE(0);
    // ---- row pass -------------------------------------------------------------------------
    if (reg_rows) {
      const int n4 = C4 >> 2;
      for (int r = warp; r < nrows; r += kWarps) {
        const float4* z4 = reinterpret_cast<const float4*>(slab + (size_t)r * a.slab_ld);
        const float4* v4 = reinterpret_cast<const float4*>(v_s);
        float4 x[kRowChunks];
        float m = -CUDART_INF_F;
#pragma unroll
        for (int k = 0; k < kRowChunks; ++k) {
          const int i = lane + 32 * k;
          if (i < n4) {
            const float4 zz = z4[i], vv = v4[i];
            x[k] = make_float4(zz.x + vv.x, zz.y + vv.y, zz.z + vv.z, zz.w + vv.w);
            m = fmaxf(m, fmaxf(fmaxf(x[k].x, x[k].y), fmaxf(x[k].z, x[k].w)));
          } else {
            x[k] = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
          }
        }
        m = warp_max(m);
        float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
#pragma unroll
        for (int k = 0; k < kRowChunks; ++k) {
          s0 += ex2((x[k].x - m) * kLog2e); s1 += ex2((x[k].y - m) * kLog2e);
          s2 += ex2((x[k].z - m) * kLog2e); s3 += ex2((x[k].w - m) * kLog2e);
        }
        const float s = warp_sum((s0 + s1) + (s2 + s3));
        if (lane == 0) {
          float lmu = (r_begin + r == n0) ? log_mu_last : norm;
          u_s[r] = lmu - (m + logf(s));
        }
      }
    } else {
      for (int r = warp; r < nrows; r += kWarps) {
        const float* z = zbase + (size_t)r * zs;
        float m = -CUDART_INF_F, s = 0.f;
#pragma unroll 4
        for (int j = lane; j < C; j += 32) m = fmaxf(m, z[j] + v_s[j]);
        m = warp_max(m);
#pragma unroll 4
        for (int j = lane; j < C; j += 32) s += ex2((z[j] + v_s[j] - m) * kLog2e);
        s = warp_sum(s);
        if (lane == 0) {
          float lmu = (r_begin + r == n0) ? log_mu_last : norm;
          u_s[r] = lmu - (m + logf(s));
        }
      }
    }
    __syncthreads();
    SINK_TRACE(1);
    // ---- column pass: per-CTA partial LSE, published with the iteration tag in sign(s) --------------
    const unsigned epoch = (unsigned)(it + 1);
    if (b < Ga) {
      if (reg_cols) {
        const int Gf = C >> 2;                               // groups of 4 columns, one 128-bit LDS per row
        for (int g = tid; g < Gf; g += kThreads) {
          float4 x[kRMax];
          float4 m = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
#pragma unroll
          for (int r = 0; r < kRMax; ++r) {
            if (r < nrows) {
              const float4 zz = *reinterpret_cast<const float4*>(slab + (size_t)r * a.slab_ld + 4 * g);
              const float ur = u_s[r];
              x[r] = make_float4(zz.x + ur, zz.y + ur, zz.z + ur, zz.w + ur);
              m.x = fmaxf(m.x, x[r].x); m.y = fmaxf(m.y, x[r].y); m.z = fmaxf(m.z, x[r].z); m.w = fmaxf(m.w, x[r].w);
            } else {
              x[r] = make_float4(-CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F, -CUDART_INF_F);
            }
          }
          float4 sm = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
          for (int r = 0; r < kRMax; ++r) {
            sm.x += ex2((x[r].x - m.x) * kLog2e); sm.y += ex2((x[r].y - m.y) * kLog2e);
            sm.z += ex2((x[r].z - m.z) * kLog2e); sm.w += ex2((x[r].w - m.w) * kLog2e);
          }
          float2* dst = &a.part[(size_t)b * a.ldp + 4 * g];
          dst[0] = make_float2(m.x, sm.x); dst[1] = make_float2(m.y, sm.y);
          dst[2] = make_float2(m.z, sm.z); dst[3] = make_float2(m.w, sm.w);
        }
        // the (at most 3) columns left over — among them the dustbin column — are reduced across the lanes of a warp
        const int jl = 4 * Gf + warp;
        if (jl < C) {
          const float xx = (lane < nrows) ? slab[(size_t)lane * a.slab_ld + jl] + u_s[lane] : -CUDART_INF_F;
          const float mm = warp_max(xx);
          const float ss = warp_sum(ex2((xx - mm) * kLog2e));
          if (lane == 0) a.part[(size_t)b * a.ldp + jl] = make_float2(mm, ss);
        }
      } else {
        for (int j = tid; j < C; j += kThreads) {
          float m = -CUDART_INF_F;
          for (int r = 0; r < nrows; ++r) m = fmaxf(m, zbase[(size_t)r * zs + j] + u_s[r]);
          float s = 0.f;
          for (int r = 0; r < nrows; ++r) s += ex2((zbase[(size_t)r * zs + j] + u_s[r] - m) * kLog2e);
          a.part[(size_t)b * a.ldp + j] = make_float2(m, s);
        }
      }
    }
    publish_flag(&a.pflag[b], epoch);
    SINK_TRACE(2);
    // ---- combine the partials of my column range -> v_j ------------------------------------------------
    wait_flags(a.pflag, Ga, epoch, a.err);
    for (int jc = c_begin; jc < c_end; jc += 16) {
      const int jj = tid & 15, gs = tid >> 4, j = jc + jj;
      float m = -CUDART_INF_F, s = 0.f;
      if (j < c_end) {
        float2 w[8];                                   // grid <= 256 -> at most 8 producers per (column, slot)
        const int cnt = (Ga - gs + 31) >> 5;
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < cnt) w[k] = __ldcg(&a.part[(size_t)(gs + 32 * k) * a.ldp + j]);
#pragma unroll
        for (int k = 0; k < 8; ++k)
          if (k < cnt) lse_merge(m, s, w[k].x, w[k].y);
      }
      red_m[gs][jj] = m; red_s[gs][jj] = s;
      __syncthreads();
      SINK_TRACE(3);
      const int jw = jc + warp;                      // warp w finishes column jc + w
      if (jw < c_end) {
        float mm = red_m[lane][warp], ss = red_s[lane][warp];
        float M = warp_max(mm);
        ss = (ss > 0.f) ? ss * ex2((mm - M) * kLog2e) : 0.f;
        ss = warp_sum(ss);
        if (lane == 0) {
          float lnu = (jw == n1) ? log_nu_last : norm;
          float vj = lnu - (M + logf(ss));
          a.vg[jw] = vj;
          if (it == a.iters - 1) a.v[jw] = vj;
        }
      }
      __syncthreads();
    }
    publish_flag(&a.vflag[b], epoch);
    SINK_TRACE(4);
    // ---- gather the new v ----------------------------------------------------------------------------------
    wait_flags(a.vflag, G, epoch, a.err);
    for (int j = tid; j < C; j += kThreads) v_s[j] = __ldcg(&a.vg[j]);
    __syncthreads();
    SINK_TRACE(5);
  }

  // ---- final pass: Z = ((Z0 + u) + v) - norm, row / column max + first argmax -------------------
  if (fast) { zbase = a.Z + (size_t)r_begin * a.ld; }            // the slab holds exp(z - rowmax): re-read Z0 (L2)
  const int zsf = fast ? a.ld : zs;
  for (int r = warp; r < nrows; r += kWarps) {
    int gr = r_begin + r;
    float ur = (a.iters > 0) ? u_s[r] : 0.f;
    if (lane == 0) a.u[gr] = ur;
    if (gr >= n0) continue;
    const float* z = zbase + (size_t)r * zsf;
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
        float t = ((zbase[(size_t)r * zsf + j] + ur) + vj) - norm;
        if (t > best) { best = t; bi = r_begin + r; }
      }
      a.part[(size_t)b * a.ldp + j] = make_float2(best, __int_as_float(bi));
    }
  }
  publish_flag(&a.pflag[b], (unsigned)(a.iters + 1));
  wait_flags(a.pflag, Ga, (unsigned)(a.iters + 1), a.err);
  for (int jc = c_begin; jc < c_end; jc += 16) {
    const int jj = tid & 15, gs = tid >> 4, j = jc + jj;
    float best = -CUDART_INF_F;
    int bi = 0x7fffffff;
    if (j < c_end && j < n1) {
      const int cnt = (Ga - gs + 31) >> 5;
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        if (k < cnt) {
          const float2 w = __ldcg(&a.part[(size_t)(gs + 32 * k) * a.ldp + j]);
          const float pv = w.x;
          const int pi = __float_as_int(w.y);
          if (pv > best || (pv == best && pi < bi)) { best = pv; bi = pi; }
        }
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
                                 float* __restrict__ ms0, float* __restrict__ ms1, const unsigned* __restrict__ err,
                                 unsigned* __restrict__ status) {
  int n0 = n_dev ? min(n_dev[0], n0_max) : n0_max;
  int n1 = n_dev ? min(n_dev[1], n1_max) : n1_max;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  // err[0]: a poll timed out; err[1]: which iteration the launch ran (GIMS_STATUS_SINKHORN_FAST / _EXACT)
  if (t == 0 && status) atomicOr(status, (err[0] ? GIMS_STATUS_SINKHORN_TIMEOUT : 0u) | err[1]);
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
  unsigned* pflag;
  unsigned* vflag;
  u64* mm;
  float* colsum;
  size_t zero_bytes;
  float2* part;
  float* vg;
  float* max0;
  float* max1;
};

size_t carve(SinkWs& w, void* base, size_t cap, int n0_max, int n1_max, int grid) {
  Arena a(base, cap);
  size_t ldp = (size_t)n1_max + 1;
  w.err = a.take<unsigned>(64);
  w.pflag = a.take<unsigned>(kMaxGrid);
  w.vflag = a.take<unsigned>(kMaxGrid);
  w.mm = a.take<u64>(2 * (size_t)kMaxGrid);
  w.colsum = a.take<float>(3 * kWays * ((ldp + 3) & ~(size_t)3));
  w.zero_bytes = align_up(a.off, 256);
  w.part = a.take<float2>((size_t)grid * ldp);
  w.vg = a.take<float>(ldp + 4);
  w.max0 = a.take<float>(n0_max + 1);
  w.max1 = a.take<float>(n1_max + 1);
  return align_up(a.off, 256);
}

}  // namespace

}  // namespace gims

using namespace gims;

extern "C" int gims_debug_sinkhorn_trace(long long* dev_buf) {
  GIMS_CUDA_OK(cudaMemcpyToSymbol(g_sink_trace, &dev_buf, sizeof(dev_buf)));
  return GIMS_OK;
}

extern "C" size_t gims_sinkhorn_workspace_bytes(int n0_max, int n1_max) {
  SinkWs w;
  return carve(w, nullptr, 0, n0_max, n1_max, kMaxGrid);
}

// largest n1_max (image-1 keypoints) whose potentials fit the kernel's shared memory, for n0_max <= GIMS_MAX_KPTS
extern "C" int gims_sinkhorn_max_columns(void) {
  int dev = 0, sms = 0, smem_optin = 0;
  if (cudaGetDevice(&dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess ||
      cudaDeviceGetAttribute(&smem_optin, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess || sms < 1)
    return 0;
  int G = sms < kMaxGrid ? sms : kMaxGrid;
  long long budget = (long long)smem_optin - 5120;
  long long rlen = ((GIMS_MAX_KPTS + 1 + G - 1) / G + 3) & ~3;
  long long c = (budget / 4 - 4 * rlen) / 2 - 4;         // 2 * round4(C) + 4 * rlen floats <= budget
  return (int)(c < 1 ? 0 : c - 1);
}

extern "C" int gims_sinkhorn_match(const float* couplings, int n0_max, int n1_max, const int* n_dev, int iters,
                                   float match_threshold, void* workspace, size_t workspace_bytes, float* u, float* v,
                                   int* indices0, int* indices1, int64_t* matches0, int64_t* matches1, float* mscores0,
                                   float* mscores1, unsigned* status_dev, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!couplings || !workspace || !u || !v || !indices0 || !indices1 || !matches0 || !matches1 || !mscores0 || !mscores1) {
    set_error("gims_sinkhorn_match: null pointer argument");
    return GIMS_ERR_ARG;
  }
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
  size_t fixed = (size_t)(2 * ((C + 3) & ~3) + 4 * ((rpc + 3) & ~3)) * sizeof(float);   // v, w | u, ut, rmax, e
  size_t slab_bytes = (size_t)rpc * slab_ld * sizeof(float);
  size_t budget = (size_t)smem_optin - 5120;     // static smem (red_m / red_s) + margin
  SinkArgs a;
  a.rpc_max = rpc;
  a.slab_rows = (fixed + slab_bytes <= budget) ? rpc : 0;
  a.slab_ld = slab_ld;
  size_t dyn = fixed + (a.slab_rows ? slab_bytes : 0);
  if (dyn > budget) { set_error("gims_sinkhorn_match: n1_max=%d needs %zu B of shared memory", n1_max, dyn); return GIMS_ERR_ARG; }
  a.Z = couplings; a.ld = C; a.n0_max = n0_max; a.n1_max = n1_max; a.n_dev = n_dev; a.iters = iters;
  a.u = u; a.v = v; a.part = w.part; a.vg = w.vg; a.pflag = w.pflag; a.vflag = w.vflag; a.mm = w.mm; a.colsum = w.colsum; a.counter = w.err + 32; a.ldp = C; a.err = w.err;
  a.idx0 = indices0; a.idx1 = indices1; a.max0 = w.max0; a.max1 = w.max1;
  // the attribute is per function, not per launch: always the full budget, so that concurrent callers with different
  // problem sizes cannot lower it under one another's launch (found by test_concurrent_callers_match_sequential)
  GIMS_CUDA_OK(cudaFuncSetAttribute(k_sinkhorn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)budget));
  int per_sm = 0;
  GIMS_CUDA_OK(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_sinkhorn, kThreads, dyn));
  if (per_sm < 1) { set_error("gims_sinkhorn_match: kernel does not fit an SM (dyn smem %zu)", dyn); return GIMS_ERR_ARG; }
  GIMS_CUDA_OK(cudaMemsetAsync(w.err, 0, w.zero_bytes, st));    // clears err and every flag / tag word (contiguous)
  void* params[] = {&a};
  {
    CoopChainScope chain(st);
    GIMS_TRY(chain.rc);
    ProfScope prof(GIMS_PROF_SINKHORN, st);
    GIMS_CUDA_OK(cudaLaunchCooperativeKernel((const void*)k_sinkhorn, dim3(G), dim3(kThreads), params, dyn, st));
  }
  count_launch();
  int m = n0_max > n1_max ? n0_max : n1_max;
  k_match_finalize<<<cdiv(m, 256), 256, 0, st>>>(n0_max, n1_max, n_dev, indices0, indices1, w.max0, match_threshold,
                                                 matches0, matches1, mscores0, mscores1, w.err, status_dev);
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}
