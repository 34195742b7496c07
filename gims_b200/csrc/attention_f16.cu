// Dense multi-head attention on tcgen05 tensor cores with 16-bit operand planes (kind::f16), flash style.
//
// Replaces `attention` + the einsums of MultiHeadedAttention (models/gmatcher.py:35-39, 108-113) for head_dim 64:
// out = softmax(q^T k / 8) v per head, without materialising the (4, N, M) probabilities.
//
// Two precisions, one kernel template:
//   PLANES = 2, fp16   fp32-class accuracy.  Every operand is x = hi + lo with hi = fp16(x), lo = fp16(x - hi)
//                      (|x - hi - lo| <= max(2^-22 |x|, 2^-25)); a product is hi*hi + hi*lo + lo*hi = three
//                      kind::f16 MMAs, which run at TWICE the kind::tf32 rate: half the tensor-pipe time of the 3xTF32
//                      kernel (attention_tc.cu) at the same error class.  fp16 overflows at 65504: the producers
//                      flag |x| >= 32768 in the status word and the host reruns the pair on the tf32 kernels.
//   PLANES = 1, bf16   the "bf16 variant" of BASELINE.json (one MMA per product, 8-bit mantissa) — reported separately.
//
// Operands (written by the QKV projection epilogue, gemm_tc.cu MODE 3):
//   Qp [rows][256]            fp32 q * log2(e)/8, channels h*64+d; converted here, once per CTA
//   Kp [PLANES][rows][256]    16-bit planes of k
//   Vt [PLANES][256][ldv]     16-bit planes of v, TRANSPOSED (channel-major) so that PV's B operand is K-major
//
// One CTA = 2 x 128 queries of one head of one image; 384 threads:
//   warps 0..3 / 4..7  softmax warps of query tile 0 / 1: thread = query row = TMEM lane.  Per 64-key tile:
//                      tcgen05.ld S -> online max / ex2 / sum in fp32 -> P split into 16-bit planes -> tcgen05.st
//                      -> fold the finished PV tile into register accumulators (O = O * alpha + PV, fp32 RN)
//   warp 8             one thread: S = Q K^T of query tile 0, and the TMA loads of the K ring
//   warp 9             one thread: PV = P V of query tile 0, and the TMA loads of the Vt ring
//   warps 10, 11       the same two issuers for query tile 1 (no TMA)
// Per 64-key tile and query tile the softmax chain (TMEM load, max, 64 ex2, convert, TMEM store, fold) is ~3200 clk of
// mostly latency; two tiles per CTA overlap two such chains on one tensor pipe (2 x 24 MMAs = 2300 clk per 64 keys).
// Splitting a row over two threads (profiles/r02_attn16_trace_v5*.txt) did not shorten the chain: the exponentials are
// MUFU-bound (16 per clock and SM) and the extra exchange of the row maxima costs a barrier.
// Four independent issuers: with one QK and one PV thread serving both tiles in order, each tile's PV waited for the
// other tile's softmax.  Both query tiles share every K / Vt tile in shared memory; they run in anti-phase (tile 1 starts
// half an iteration late), so that one tile's exponentials and tensor-pipe bursts fall into the other tile's load /
// convert / fold phases.
//
// TMEM columns of query tile t (base 256 t): S 64 | P_hi 32 | P_lo 32 | PV 64 | Q_hi 32 | Q_lo 32.  A operands are
// packed two 16-bit values per column (even k in the low half).  Within a tile the small correction products go
// first and hi*hi last: the tensor core truncates the accumulator at every step, and this order keeps those
// roundings at the magnitude of the small terms for as long as possible.
#include <math_constants.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gims {

namespace {

using namespace tc;

constexpr int TQ = 128;        // queries per tile (UMMA M)
constexpr int NQT = 2;         // query tiles per CTA
constexpr int TKV = 64;        // keys per tile (UMMA N of S, K of PV)
constexpr int HD = 64;         // head dim
constexpr int kStages = 4;
constexpr int kPlaneBytes = TKV * HD * 2;       // 8 KB: 64 rows x 128 bytes
constexpr int kThreads16 = 384;                 // 8 softmax warps + 4 MMA issuer warps at 168 registers per thread (no
                                                // setmaxnreg: a CTA can only redistribute the registers it was LAUNCHED with,
                                                // and 384 x 168 already is the whole register file)

template <int PLANES> struct Cfg16 {
  static constexpr int kStageBytes = PLANES * kPlaneBytes;
  static constexpr int kSmem = 2 * kStages * kStageBytes + 1024 + 512;
};

constexpr int cS = 0, cPh = 64, cPl = 96, cO = 128, cQh = 192, cQl = 224, cTile = 256;

// Optional pipeline trace (bring-up / profiling): CTA (0,0,0) stores clock64() stamps, 16 slots per 64-key tile j < 40:
//   [0] QK(j) of tile 0: operands ready  [1] issued   [2] PV(j) of tile 0: operands ready  [3] issued
//   [4] softmax tile 0: S(j) observed  [5] P(j) handed over  [6] PV(j-1) folded   [7] softmax tile 1: S(j) observed
//   [8] S(j) in registers  [9] row max known  [10] exponentials and row sum done  [11] PV(j-1) complete (o_full)
__device__ long long* g_attn16_trace = nullptr;

struct Attn16Args {
  const float* qp;
  float* out;                 // [rows][256]
  Segs segs;
  int cross;
  int rows_total;             // plane stride of Kp in rows
  int vbase[kMaxSegs];        // first Vt key column of each segment
  unsigned* status;           // GIMS_STATUS_FP16_RANGE is OR-ed in on overflow (fp16 only); may be null
};

__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// a 64-bit value every lane of the (converged) warp holds, made provably warp-uniform for the compiler
__device__ __forceinline__ uint64_t uniform64(uint64_t x) {
  const uint32_t lo = __shfl_sync(0xffffffffu, (uint32_t)x, 0), hi = __shfl_sync(0xffffffffu, (uint32_t)(x >> 32), 0);
  return ((uint64_t)hi << 32) | lo;
}

template <int PLANES, int FMT>
__global__ void __launch_bounds__(kThreads16, 1)
k_attention_f16(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapVt, Attn16Args a) {
  extern __shared__ uint8_t smem_raw[];
  using Cfg = Cfg16<PLANES>;
  const int img = blockIdx.z, head = blockIdx.y;
  const int src = a.cross ? (img ^ 1) : img;
  // counts come from device memory; the shuffle makes them provably warp-uniform for the compiler, so that the
  // single-thread TMA / MMA loops below compile to uniform-datapath code (no per-operand R2UR moves)
  const int nq = __shfl_sync(0xffffffffu, seg_count(a.segs, img), 0);
  const int nk = __shfl_sync(0xffffffffu, seg_count(a.segs, src), 0);
  const int q0 = blockIdx.x * (NQT * TQ);
  if (q0 >= nq || nk <= 0) return;
  const int nqt = (q0 + TQ < nq) ? 2 : 1;          // live query tiles of this CTA
  const int qrow0 = a.segs.base[img] + q0;         // global row of the first query
  const int krow0 = a.segs.base[src];              // global row of the first key
  const int ntiles = (nk + TKV - 1) / TKV;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* k_smem = smem;
  uint8_t* v_smem = k_smem + kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + kStages * Cfg::kStageBytes);
  uint64_t* k_full = bars;                       // [kStages] TMA -> MMA
  uint64_t* k_empty = bars + kStages;            // [kStages] QK MMAs of both tiles retired (tcgen05.commit)
  uint64_t* v_full = bars + 2 * kStages;
  uint64_t* v_empty = bars + 3 * kStages;
  uint64_t* s_full = bars + 4 * kStages;         // [2] S of tile t written (commit)
  uint64_t* s_free = s_full + 2;                 // [2] S of tile t read by its softmax warps (4 arrivals)
  uint64_t* p_full = s_free + 2;                 // [2] P of tile t stored (4 arrivals)
  uint64_t* o_full = p_full + 2;                 // [2] PV of tile t written (commit)
  uint64_t* o_free = o_full + 2;                 // [2] PV of tile t folded into registers (4 arrivals)
  uint64_t* q_ready = o_free + 2;                // [2] Q planes of tile t stored into TMEM (4 arrivals)
  uint64_t* skew = q_ready + 2;                  // [1] tile 0 has finished the exponentials of its first key tile (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(skew + 1);


  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpQk = 8, kWarpPv = 9;                      // tile 1: kWarpQk + 2, kWarpPv + 2
  if (warp == kWarpQk && lane == 0) {
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapVt);
    for (int s = 0; s < kStages; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], nqt);   // one commit per live query tile
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], nqt);
    }
    for (int t = 0; t < 2; ++t) {
      mbar_init(&s_full[t], 1); mbar_init(&s_free[t], 4);     // 4 = the softmax warps of a query tile
      mbar_init(&p_full[t], 4); mbar_init(&o_full[t], 1); mbar_init(&o_free[t], 4);
      mbar_init(&q_ready[t], 4);
    }
    mbar_init(skew, 4);
    fence_barrier_init();
  }
  // This thread's Q row is requested before the CTA-wide sync: the global round trip overlaps the TMEM allocation and
  // the barrier hand-shake instead of following them.
  const int qt = warp >> 2;                                    // query tile of a softmax warp
  float4 qreg[HD / 4];
  if (warp < 8) {
    const int qr = q0 + TQ * qt + 32 * (warp & 3) + lane;
    const float4* p4 = reinterpret_cast<const float4*>(a.qp + (size_t)(qrow0 + TQ * qt + 32 * (warp & 3) + lane) * kD + head * HD);
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) qreg[i] = (qr < nq) ? p4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (warp == kWarpQk) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // warp-uniform for the compiler

  const int wt = __shfl_sync(0xffffffffu, (warp - kWarpQk) >> 1, 0);  // query tile of an issuer warp (warp-uniform)
  if (warp >= 8) {
    if ((threadIdx.x == kWarpQk * 32 || threadIdx.x == (kWarpQk + 2) * 32) && wt < nqt) {
      // ===== QK issuer of query tile t: ONE thread; tile 0's issuer also feeds the K ring =====
      const int t = wt;
      constexpr uint32_t idesc = umma_idesc_f16(TQ, TKV, FMT);
      const uint64_t dk0 = umma_desc_sw128(smem_u32(k_smem));
      const uint32_t sacc = tmem + cTile * t + cS, q_hi = tmem + cTile * t + cQh, q_lo = tmem + cTile * t + cQl;
      auto load_k = [&](int j, int s) {                       // K box: 64 keys x 64 channels per plane
        mbar_arrive_expect_tx(&k_full[s], Cfg::kStageBytes);
        for (int pl = 0; pl < PLANES; ++pl)
          tma_load_2d(k_smem + s * Cfg::kStageBytes + pl * kPlaneBytes, &mapK, &k_full[s], head * HD, pl * a.rows_total + krow0 + j * TKV);
      };
      if (t == 0)
        for (int j = 0; j < kStages && j < ntiles; ++j) load_k(j, j);
      mbar_wait(&q_ready[t], 0);
      long long* trace = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && t == 0 ? g_attn16_trace : nullptr;
      int s = 0; uint32_t ph = 0;
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(&k_full[s], ph);
        if (j > 0) mbar_wait(&s_free[t], (uint32_t)(j - 1) & 1u);      // S_t(j-1) has been read out
        tcgen05_fence_after();
        if (trace && j < 40) trace[j * 16 + 0] = clock64();
        const uint64_t kh = dk0 + (uint64_t)(s * (Cfg::kStageBytes >> 4)), kl = kh + (kPlaneBytes >> 4);
        // K-dim = 64 channels = 4 k-steps of 16 (8 packed TMEM columns of Q, 32 bytes of the swizzle row of K)
        if (PLANES == 2) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_f16_ts(sacc, q_lo + ks * 8, kh + ks * 2, idesc, ks ? 1u : 0u);
            umma_f16_ts(sacc, q_hi + ks * 8, kl + ks * 2, idesc, 1u);
          }
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_f16_ts(sacc, q_hi + ks * 8, kh + ks * 2, idesc, (PLANES == 2 || ks) ? 1u : 0u);
        umma_commit(&s_full[t]);
        umma_commit(&k_empty[s]);
        if (trace && j < 40) trace[j * 16 + 1] = clock64();
        if (t == 0 && j >= 1 && j - 1 + kStages < ntiles) {     // refill the stage of tile j-1 (its MMAs were issued a whole
          const int sp = s == 0 ? kStages - 1 : s - 1;          // iteration ago) with tile j-1+kStages
          mbar_wait(&k_empty[sp], (uint32_t)((j - 1) / kStages) & 1u);
          load_k(j - 1 + kStages, sp);
        }
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    } else if ((threadIdx.x == kWarpPv * 32 || threadIdx.x == (kWarpPv + 2) * 32) && wt < nqt) {
      // ===== PV issuer of query tile t: ONE thread; tile 0's issuer also feeds the Vt ring =====
      const int t = wt;
      constexpr uint32_t idesc = umma_idesc_f16(TQ, HD, FMT);
      const uint64_t dv0 = umma_desc_sw128(smem_u32(v_smem));
      const uint32_t oacc = tmem + cTile * t + cO, p_hi = tmem + cTile * t + cPh, p_lo = tmem + cTile * t + cPl;
      auto load_v = [&](int j, int s) {                       // Vt box: 64 channels x 64 keys per plane
        mbar_arrive_expect_tx(&v_full[s], Cfg::kStageBytes);
        for (int pl = 0; pl < PLANES; ++pl)
          tma_load_2d(v_smem + s * Cfg::kStageBytes + pl * kPlaneBytes, &mapVt, &v_full[s], a.vbase[src] + j * TKV, pl * kD + head * HD);
      };
      if (t == 0)
        for (int j = 0; j < kStages && j < ntiles; ++j) load_v(j, j);
      long long* trace = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && t == 0 ? g_attn16_trace : nullptr;
      int s = 0; uint32_t ph = 0;
      for (int j = 0; j < ntiles; ++j) {
        mbar_wait(&v_full[s], ph);
        mbar_wait(&p_full[t], (uint32_t)j & 1u);                       // P_t(j) is in TMEM
        if (j > 0) mbar_wait(&o_free[t], (uint32_t)(j - 1) & 1u);      // PV_t(j-1) has been folded
        tcgen05_fence_after();
        if (trace && j < 40) trace[j * 16 + 2] = clock64();
        const uint64_t vh = dv0 + (uint64_t)(s * (Cfg::kStageBytes >> 4)), vl = vh + (kPlaneBytes >> 4);
        if (PLANES == 2) {
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            umma_f16_ts(oacc, p_lo + ks * 8, vh + ks * 2, idesc, ks ? 1u : 0u);
            umma_f16_ts(oacc, p_hi + ks * 8, vl + ks * 2, idesc, 1u);
          }
        }
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) umma_f16_ts(oacc, p_hi + ks * 8, vh + ks * 2, idesc, (PLANES == 2 || ks) ? 1u : 0u);
        umma_commit(&o_full[t]);
        umma_commit(&v_empty[s]);
        if (trace && j < 40) trace[j * 16 + 3] = clock64();
        if (t == 0 && j >= 1 && j - 1 + kStages < ntiles) {
          const int sp = s == 0 ? kStages - 1 : s - 1;
          mbar_wait(&v_empty[sp], (uint32_t)((j - 1) / kStages) & 1u);
          load_v(j - 1 + kStages, sp);
        }
        if (++s == kStages) { s = 0; ph ^= 1; }
      }
    }
  } else {
    // ===== softmax / accumulate warps: thread <-> query row of tile qt =====
    if (qt < nqt) {
      const uint32_t lane_base = tmem + ((uint32_t)(32 * (warp & 3)) << 16) + (uint32_t)(cTile * qt);
      const int row = q0 + TQ * qt + 32 * (warp & 3) + lane;
      bool ovf = false;
      {   // Q row -> packed 16-bit planes in TMEM (zeros for rows past the live count: nothing of them is ever stored)
        uint32_t vh[32], vl[32];
#pragma unroll
        for (int i = 0; i < HD / 4; ++i) {
          const float4 x = qreg[i];
          if (FMT == 0) ovf = ovf || fmaxf(fmaxf(fabsf(x.x), fabsf(x.y)), fmaxf(fabsf(x.z), fabsf(x.w))) >= 32768.f;
          if (PLANES == 2) {
            split16x2<FMT>(x.x, x.y, vh[2 * i], vl[2 * i]);
            split16x2<FMT>(x.z, x.w, vh[2 * i + 1], vl[2 * i + 1]);
          } else {
            vh[2 * i] = pack16<FMT>(x.x, x.y);
            vh[2 * i + 1] = pack16<FMT>(x.z, x.w);
          }
        }
        tmem_st_32x32(lane_base + cQh, vh);
        if (PLANES == 2) tmem_st_32x32(lane_base + cQl, vl);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&q_ready[qt]);
      }
      if (__any_sync(0xffffffffu, ovf) && lane == 0 && a.status) atomicOr(a.status, GIMS_STATUS_FP16_RANGE);
      float o[HD];
#pragma unroll
      for (int d = 0; d < HD; ++d) o[d] = 0.f;
      float m_run = -CUDART_INF_F, l_run = 0.f, alpha_prev = 0.f;
      long long* trace = ((blockIdx.x | blockIdx.y | blockIdx.z) == 0 && (threadIdx.x & 127) == 0) ? g_attn16_trace : nullptr;
      for (int j = 0; j <= ntiles; ++j) {
        float s[TKV];                                      // scores in the log2 domain (Q was scaled by log2(e)/8)
        float alpha = 0.f;
        uint32_t ph_[32], pl_[32];                         // P(j) as packed planes
        if (j < ntiles) {
          // The two query tiles run in ANTI-PHASE: tile 1 starts half an iteration late, so that one tile's exponentials
          // (MUFU) and tensor-pipe bursts fall into the other tile's load / convert / fold phases.  Started together
          // they stay in lockstep: 2 x 520 clk of MUFU back to back, then both wait for the same PV burst.
          if (j == 0 && qt == 1) mbar_wait(skew, 0);
          mbar_wait(&s_full[qt], (uint32_t)j & 1u);
          tcgen05_fence_after();
          if (trace && j < 40) trace[j * 16 + (qt ? 7 : 4)] = clock64();
          {   // both halves are requested before the one wait: one TMEM round trip instead of two
            uint32_t v0[32], v1[32];
            tmem_ld_32x32(lane_base + cS, v0);
            tmem_ld_32x32(lane_base + cS + 32, v1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) { s[i] = __uint_as_float(v0[i]); s[32 + i] = __uint_as_float(v1[i]); }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&s_free[qt]);         // the tensor core may overwrite S of this tile
          if (trace && j < 40 && qt == 0) trace[j * 16 + 8] = clock64();
          const int valid = nk - j * TKV;                  // keys of this tile that exist
          if (valid < TKV) {
#pragma unroll
            for (int i = 0; i < TKV; ++i)
              if (i >= valid) s[i] = -CUDART_INF_F;
          }
          float mx8[8];                                    // eight independent chains, then a tree
#pragma unroll
          for (int i = 0; i < 8; ++i) mx8[i] = s[i];
#pragma unroll
          for (int i = 8; i < TKV; ++i) mx8[i & 7] = fmaxf(mx8[i & 7], s[i]);
          const float mx = fmaxf(fmaxf(fmaxf(mx8[0], mx8[1]), fmaxf(mx8[2], mx8[3])), fmaxf(fmaxf(mx8[4], mx8[5]), fmaxf(mx8[6], mx8[7])));
          const float m_new = fmaxf(m_run, mx);
          if (trace && j < 40 && qt == 0) trace[j * 16 + 9] = clock64() + (long long)(__float_as_int(m_new) & 0);
          alpha = ex2_approx(m_run - m_new);
          float rs[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
          for (int i = 0; i < TKV; i += 4) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              s[i + q] = ex2_approx(s[i + q] - m_new);
              rs[q] += s[i + q];
            }
          }
          l_run = l_run * alpha + ((rs[0] + rs[1]) + (rs[2] + rs[3]));
          m_run = m_new;
          if (trace && j < 40 && qt == 0) trace[j * 16 + 10] = clock64() + (long long)(__float_as_int(l_run) & 0);
          if (j == 0 && qt == 0) { __syncwarp(); if (lane == 0) mbar_arrive(skew); }
#pragma unroll
          for (int i = 0; i < 32; ++i) {                   // converted BEFORE waiting for PV(j-1)
            if (PLANES == 2) split16x2<FMT>(s[2 * i], s[2 * i + 1], ph_[i], pl_[i]);
            else ph_[i] = pack16<FMT>(s[2 * i], s[2 * i + 1]);
          }
        }
        if (j > 0) {                                       // PV(j-1) is complete: P may be overwritten, O is final
          mbar_wait(&o_full[qt], (uint32_t)(j - 1) & 1u);
          tcgen05_fence_after();
          if (trace && j < 40 && qt == 0) trace[j * 16 + 11] = clock64();
        }
        if (j < ntiles) {                                  // P(j) -> TMEM
          tmem_st_32x32(lane_base + cPh, ph_);
          if (PLANES == 2) tmem_st_32x32(lane_base + cPl, pl_);
          tmem_st_wait();
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&p_full[qt]);
          if (trace && j < 40 && qt == 0) trace[j * 16 + 5] = clock64();
        }
        if (j > 0) {                                       // fold PV(j-1): O = O * alpha(j-1) + PV
          {
            uint32_t v0[32], v1[32];
            tmem_ld_32x32(lane_base + cO, v0);
            tmem_ld_32x32(lane_base + cO + 32, v1);
            tmem_ld_wait();
#pragma unroll
            for (int i = 0; i < 32; ++i) {
              o[i] = fmaf(o[i], alpha_prev, __uint_as_float(v0[i]));
              o[32 + i] = fmaf(o[32 + i], alpha_prev, __uint_as_float(v1[i]));
            }
          }
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive(&o_free[qt]);         // the tensor core may overwrite PV of this tile
          if (trace && j < 40 && qt == 0) trace[j * 16 + 6] = clock64();
        }
        alpha_prev = alpha;
      }
      if (row < nq) {
        const float inv = 1.f / l_run;
        float4* dst = reinterpret_cast<float4*>(a.out + (size_t)(a.segs.base[img] + row) * kD + head * HD);
#pragma unroll
        for (int d = 0; d < HD; d += 4) dst[d / 4] = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
      }
      tcgen05_fence_before();
    }
  }
  __syncthreads();
  if (warp == kWarpQk) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

template <int PLANES, int FMT>
int launch16(const CUtensorMap& mK, const CUtensorMap& mV, const Attn16Args& a, dim3 grid, cudaStream_t st) {
  using Cfg = Cfg16<PLANES>;
  GIMS_CUDA_OK(cudaFuncSetAttribute(k_attention_f16<PLANES, FMT>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmem));
  k_attention_f16<PLANES, FMT><<<grid, kThreads16, Cfg::kSmem, st>>>(mK, mV, a);
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}

}  // namespace

int set_attention16_trace(long long* dev_buf) {
  GIMS_CUDA_OK(cudaMemcpyToSymbol(g_attn16_trace, &dev_buf, sizeof(dev_buf)));
  return GIMS_OK;
}

// planes as written by the 16-bit qkv-mode GEMM epilogue (QkvPlanes with fmt >= 0, common.cuh)
int launch_attention_f16(const QkvPlanes& pl, float* out, const Segs& segs, int cross, cudaStream_t st) {
  const int rows = segs_rows(segs);
  if (pl.fmt < 0 || (pl.planes != 1 && pl.planes != 2) || pl.ldv % 8) {
    set_error("launch_attention_f16: bad plane description (fmt %d, planes %d, ldv %d)", pl.fmt, pl.planes, pl.ldv);
    return GIMS_ERR_ARG;
  }
  CUtensorMap mK, mV;
  GIMS_TRY(tc::make_tmap_16_k64(&mK, pl.kp, pl.fmt, (uint64_t)pl.planes * rows, kD, kD, TKV));
  GIMS_TRY(tc::make_tmap_16_k64(&mV, pl.vt, pl.fmt, (uint64_t)pl.planes * kD, pl.ldv, pl.ldv, HD));
  Attn16Args a;
  a.qp = pl.qp;
  a.out = out;
  a.segs = segs;
  a.cross = cross;
  a.rows_total = rows;
  for (int i = 0; i < kMaxSegs; ++i) a.vbase[i] = pl.vbase[i];
  a.status = pl.status;
  dim3 grid(cdiv(segs_nmax(segs), NQT * TQ), kHeads, segs.nseg);
  ProfScope prof(GIMS_PROF_ATTENTION, st);
  if (pl.planes == 2 && pl.fmt == 0) return launch16<2, 0>(mK, mV, a, grid, st);
  if (pl.planes == 1 && pl.fmt == 1) return launch16<1, 1>(mK, mV, a, grid, st);
  set_error("launch_attention_f16: unsupported combination planes=%d fmt=%d", pl.planes, pl.fmt);
  return GIMS_ERR_ARG;
}

}  // namespace gims
