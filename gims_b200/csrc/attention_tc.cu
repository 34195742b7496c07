// Dense multi-head attention on tcgen05 tensor cores, fp32-accurate (3xTF32), flash style.
//
// Replaces `attention` + the einsums of MultiHeadedAttention (models/gmatcher.py:35-39, 108-113) for
// head_dim 64:  out = softmax(q^T k / 8) v  per head, without materialising the (4, N, M) probabilities.
//
// Operands are written by the QKV projection epilogue (gemm_tc.cu, MODE 2):
//   Qp [rows][256]      fp32 q * log2(e)/8 (scores land in the log2 domain), channels h*64+d; split here
//   Kp [2][rows][256]   tf32 hi / lo planes of k
//   Vt [2][256][ldv]    tf32 hi / lo planes of v, TRANSPOSED (channel-major) so that PV's B operand is K-major
// One CTA = 128 queries of one head of one image; 224 threads:
//   warps 0..3  one thread per query row: load the Q row (requested before the CTA-wide sync), split it and put the
//               hi / lo planes into TMEM once; per 64-key tile tcgen05.ld S ->
//               online max / ex2 / sum in fp32 -> fold the previous PV tile into register accumulators
//               (O = O * alpha + PV, fp32 RN) -> P split into tf32 hi / lo -> tcgen05.st into TMEM
//   warp 4      TMA: K and Vt tiles of 64 keys through two 3-stage rings (K is released right after S = Q K^T)
//   warp 5      one thread issues the S = Q K^T MMAs, warp 6 one thread the PV = P V MMAs (A operand in TMEM, TS form);
//               two issuers because every barrier wait / commit of a single issuer drained the tensor pipe
//               (~1000 of 3500 cycles per tile): now one of them is issuing while the other one synchronises
//
// Why Q lives in TMEM: at N = 64 an SS-form MMA already saturates the 128 B/clk shared-memory port with its own
// operand reads (4 KB of A + 2 KB of B per 48 clk, tools/micro/mma_rate.cu); with the TMA writes of the K/V ring on
// top, the QK phase ran at 67-85 clk per MMA (profiles/r01_attention_pipeline_trace.txt).  Reading A from TMEM
// leaves shared memory to the K / Vt tiles alone.
//
// TMEM columns: S[2] 64 each | P_hi 64 | P_lo 64 | PV[0] 64 | Q_hi 64 | Q_lo 64 | PV[1] 64 (all 512).  Each S / PV accumulator receives the
// 16 small correction products of its tile FIRST and the 8 hi*hi products last: the tensor core rounds the accumulator
// toward zero at every step, and this order keeps those roundings at the magnitude of the small terms for as long as
// possible (same error as a separate correction accumulator, without its TMEM columns and the extra add).
#include <math_constants.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gims {

namespace {

using namespace tc;

constexpr int TQ = 128;        // queries per CTA (UMMA M)
constexpr int TKV = 64;        // keys per tile (UMMA N of S, K of PV)
constexpr int HD = 64;         // head dim
constexpr int kStagesKV = 3;                    // Vt ring (64 keys per stage); K ring at KT = 64
constexpr int kBoxBytesKV = TKV * 32 * 4;       // 8 KB: 64 rows x 32 floats
constexpr int kKStageBytes = 4 * kBoxBytesKV;   // hi(2 boxes) lo(2 boxes); Vt stage, and K stage at KT = 64
constexpr int kAttnThreads = 224;
// KT = keys per S = Q K^T tile (UMMA N).  At 128 the QK product costs 64 clk per MMA for 128 keys instead of 2 x 45 for
// 2 x 64 (tools/micro/mma_rate.cu), the PV product and the softmax still work on 64-key halves of that tile.  The
// default is 64 (see launch_attention_tc).
template <int KT> struct AttnCfg {
  static constexpr int kKBox = KT * 32 * 4;                 // K box: KT keys x 32 channels
  static constexpr int kKStage = 4 * kKBox;
  static constexpr int kStagesK = KT == 64 ? 3 : 2;
  static constexpr int kSmem = kStagesK * kKStage + kStagesKV * kKStageBytes + 1024 + 256;
};

constexpr int cS0 = 0, cPh = 128, cPl = 192, cO0 = 256, cQh = 320, cQl = 384, cO1 = 448;

// Optional pipeline trace (bring-up / profiling): CTA (0,0,0) stores clock64() stamps per tile.
//   [j*8+0] MMA: QK(j+1) issue start   [j*8+1] MMA: PV(j) operands ready   [j*8+2] MMA: PV(j) issued
//   [j*8+3] softmax warp 0: P(j) handed over   [j*8+4] softmax warp 0: S(j) observed
__device__ long long* g_attn_trace = nullptr;

struct AttnTcArgs {
  const float* qp;            // Q planes (read directly by the softmax warps)
  float* out;                 // [rows][256]
  Segs segs;
  int cross;
  int rows_total;             // plane stride of Qp/Kp in rows
  int vbase[kMaxSegs];        // first Vt key column of each segment
};

// 2^x for x <= 0 (softmax arguments): one MUFU, flushes results below 2^-126 to zero
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Pipeline: the QK issuer runs ahead into the free S buffer while the softmax warps work on the current tile and the PV
// issuer multiplies the previous one.  S and PV are double-buffered, P is single-buffered: the softmax warps store
// P(j) as soon as o_full says that PV(j-1) has consumed P(j-1), and fold PV(j-1) out of its buffer afterwards, while
// PV(j) already accumulates into the other one.
template <int KT>
__global__ void __launch_bounds__(kAttnThreads, 1)
k_attention_tc(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapVt, AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  if (g_attn_trace && (blockIdx.x | blockIdx.y | blockIdx.z | threadIdx.x) == 0) {
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    g_attn_trace[40 * 8 + 0] = clock64();
    g_attn_trace[40 * 8 + 2] = (long long)gt;
  }
  const int img = blockIdx.z, head = blockIdx.y;
  const int src = a.cross ? (img ^ 1) : img;
  // counts come from device memory; the shuffle makes them provably warp-uniform for the compiler, so that the
  // single-thread TMA / MMA loops below compile to uniform-datapath code (no per-operand R2UR moves)
  const int nq = __shfl_sync(0xffffffffu, seg_count(a.segs, img), 0);
  const int nk = __shfl_sync(0xffffffffu, seg_count(a.segs, src), 0);
  const int q0 = blockIdx.x * TQ;
  if (q0 >= nq || nk <= 0) return;
  const int qrow0 = a.segs.base[img] + q0;       // global row of the first query
  const int krow0 = a.segs.base[src];            // global row of the first key
  const int ntiles = (nk + TKV - 1) / TKV;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* k_smem = smem;
  using Cfg = AttnCfg<KT>;
  uint8_t* v_smem = k_smem + Cfg::kStagesK * Cfg::kKStage;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + kStagesKV * kKStageBytes);
  uint64_t* k_full = bars;                       // [3] TMA -> MMA
  uint64_t* k_empty = bars + 3;                  // [3] QK MMAs retired (tcgen05.commit)
  uint64_t* v_full = bars + 6;                   // [3]
  uint64_t* v_empty = bars + 9;                  // [3] PV MMAs retired
  uint64_t* s_full = bars + 12;                  // [2] S buffer written (commit)
  uint64_t* p_full = bars + 14;                  // P written, previous PV folded (4 arrivals)
  uint64_t* o_full = bars + 15;                  // PV written (commit)
  uint64_t* q_ready = bars + 16;                 // Q planes stored into TMEM (4 arrivals)
  uint64_t* s_free = bars + 17;                  // [2] S buffer read by the softmax warps (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  // Warp roles: 0..3 = data warps (TMEM lane quarter = warp), 4 = TMA producer, 5 = QK issuer, 6 = PV issuer.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpTma = 4, kWarpMma = 5, kWarpPv = 6;
  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapVt);
    for (int s = 0; s < kStagesKV; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
    }
    mbar_init(&s_full[0], 1); mbar_init(&s_full[1], 1);
    mbar_init(&s_free[0], 4); mbar_init(&s_free[1], 4);
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    mbar_init(q_ready, 4);
    fence_barrier_init();
  }
  // This thread's Q row is requested before the CTA-wide sync: the global round trip overlaps the TMEM allocation and
  // the barrier hand-shake instead of following them.  (Not the first TMA stages: the issuing thread would stall on
  // the tensor-map fetch and hold up the sync.)
  auto tma_produce = [&]() {
    int sk = 0, sv = 0; uint32_t phk = 0, phv = 0;
    for (int j = 0; j < ntiles; ++j) {
      if (j % (KT / TKV) == 0) {                              // K tile of KT keys
        const int key0 = krow0 + j * TKV;                     // row in the K planes
        mbar_wait(&k_empty[sk], phk ^ 1);
        mbar_arrive_expect_tx(&k_full[sk], Cfg::kKStage);
        for (int pl = 0; pl < 2; ++pl)
          for (int hf = 0; hf < 2; ++hf)    // K box: KT keys x 32 channels
            tma_load_2d(k_smem + sk * Cfg::kKStage + (pl * 2 + hf) * Cfg::kKBox, &mapK, &k_full[sk], head * HD + hf * 32,
                        pl * a.rows_total + key0);
        if (++sk == Cfg::kStagesK) { sk = 0; phk ^= 1; }
      }
      const int vcol0 = a.vbase[src] + j * TKV;       // key column in the Vt planes (multiple of 64)
      mbar_wait(&v_empty[sv], phv ^ 1);
      mbar_arrive_expect_tx(&v_full[sv], kKStageBytes);
      for (int pl = 0; pl < 2; ++pl)
        for (int hf = 0; hf < 2; ++hf)      // Vt box: 64 channels x 32 keys
          tma_load_2d(v_smem + sv * kKStageBytes + (pl * 2 + hf) * kBoxBytesKV, &mapVt, &v_full[sv], vcol0 + hf * 32,
                      pl * kD + head * HD);
      if (++sv == kStagesKV) { sv = 0; phv ^= 1; }
    }
  };
  float4 qreg[HD / 4];                                         // warps 0..3: fp32 Q row (scaled by log2(e)/8)
  if (warp < 4) {
    const int qr = q0 + 32 * warp + lane;
    const float4* p4 = reinterpret_cast<const float4*>(a.qp + (size_t)(qrow0 + 32 * warp + lane) * kD + head * HD);
#pragma unroll
    for (int i = 0; i < HD / 4; ++i) qreg[i] = (qr < nq) ? p4[i] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
  if (warp == kWarpMma) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // warp-uniform for the compiler

  if (threadIdx.x == kWarpTma * 32) {
    // ===== TMA producer =====
    tma_produce();
  } else if (threadIdx.x == kWarpMma * 32) {
    // ===== QK issuer: ONE thread, all operand arithmetic in the uniform datapath (descriptor = base + constant) =====
    constexpr uint32_t idesc = umma_idesc_tf32(TQ, KT);       // M=128, N=KT
    const uint64_t dk0 = umma_desc_sw128(smem_u32(k_smem));
    const uint32_t q_hi = tmem + cQh, q_lo = tmem + cQl;
    long long* trace = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 ? g_attn_trace : nullptr;
    mbar_wait(q_ready, 0);
    int sk = 0; uint32_t phk = 0;
    for (int jq = 0; jq < ntiles; jq += KT / TKV) {
      if (trace) trace[jq * 8 + 0] = clock64();
      mbar_wait(&k_full[sk], phk);
      // the 64-key halves of S this tile overwrites have been read out: S(jq-2) at KT = 64, both halves of the
      // previous 128-key tile at KT = 128 (half h of every tile lives in columns 64 h and signals s_free[h])
      if (KT == 64) {
        if (jq >= 2) mbar_wait(&s_free[jq & 1], (uint32_t)((jq >> 1) - 1) & 1u);
      } else if (jq >= 2) {
        mbar_wait(&s_free[0], (uint32_t)((jq >> 1) - 1) & 1u);
        mbar_wait(&s_free[1], (uint32_t)((jq >> 1) - 1) & 1u);
      }
      tcgen05_fence_after();
      if (trace) trace[jq * 8 + 7] = clock64();
      const uint64_t kh = dk0 + (uint64_t)(sk * (Cfg::kKStage >> 4)), kl = kh + ((2 * Cfg::kKBox) >> 4);
      const uint32_t sacc = tmem + cS0 + (KT == 64 ? (jq & 1) * 64 : 0);
      // K-dim = 64 channels = 2 boxes x 4 k-steps (8 TMEM columns of Q each); corrections first, hi*hi last
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t koff = ((ks >> 2) * Cfg::kKBox + (ks & 3) * 32) >> 4;
        umma_tf32_ts(sacc, q_lo + ks * 8, kh + koff, idesc, ks ? 1u : 0u);
        umma_tf32_ts(sacc, q_hi + ks * 8, kl + koff, idesc, 1u);
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t koff = ((ks >> 2) * Cfg::kKBox + (ks & 3) * 32) >> 4;
        umma_tf32_ts(sacc, q_hi + ks * 8, kh + koff, idesc, 1u);
      }
      if (trace) trace[jq * 8 + 6] = clock64();
      umma_commit(&k_empty[sk]);
      umma_commit(&s_full[KT == 64 ? (jq & 1) : 0]);
      if (++sk == Cfg::kStagesK) { sk = 0; phk ^= 1; }
    }
  } else if (threadIdx.x == kWarpPv * 32) {
    // ===== PV issuer =====
    constexpr uint32_t idesc = umma_idesc_tf32(TQ, TKV);      // M=128, N=64
    const uint64_t dv0 = umma_desc_sw128(smem_u32(v_smem));
    const uint32_t p_hi = tmem + cPh, p_lo = tmem + cPl;
    long long* trace = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 ? g_attn_trace : nullptr;
    int sv = 0; uint32_t phv = 0;
    for (int j = 0; j < ntiles; ++j) {
      mbar_wait(&v_full[sv], phv);
      mbar_wait(p_full, j & 1);                                // P(j) is in TMEM (and PV(j-2) was folded before that)
      tcgen05_fence_after();
      if (trace) trace[j * 8 + 1] = clock64();
      const uint64_t vh = dv0 + (uint64_t)(sv * (kKStageBytes >> 4)), vl = vh + ((2 * kBoxBytesKV) >> 4);
      const uint32_t oacc = tmem + ((j & 1) ? cO1 : cO0);      // PV is double-buffered: the fold of PV(j-1) overlaps
      // K-dim = 64 keys = 2 Vt boxes x 4 k-steps; A (P planes) from TMEM, 8 columns per k-step
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t voff = ((ks >> 2) * kBoxBytesKV + (ks & 3) * 32) >> 4;
        umma_tf32_ts(oacc, p_lo + ks * 8, vh + voff, idesc, ks ? 1u : 0u);
        umma_tf32_ts(oacc, p_hi + ks * 8, vl + voff, idesc, 1u);
      }
#pragma unroll
      for (int ks = 0; ks < 8; ++ks) {
        const uint64_t voff = ((ks >> 2) * kBoxBytesKV + (ks & 3) * 32) >> 4;
        umma_tf32_ts(oacc, p_hi + ks * 8, vh + voff, idesc, 1u);
      }
      umma_commit(&v_empty[sv]);
      umma_commit(o_full);
      if (trace) trace[j * 8 + 2] = clock64();
      if (++sv == kStagesKV) { sv = 0; phv ^= 1; }
    }
  } else if (warp < 4) {
    // ===== softmax / accumulate warps: thread <-> query row =====
    const uint32_t lane_base = tmem + ((uint32_t)(32 * warp) << 16);
    const int row = q0 + 32 * warp + lane;
    {   // Q row -> tf32 hi / lo planes in TMEM (zeros for rows past the live count: nothing of them is ever stored)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t vh[32], vl[32];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float4 x = qreg[h * 8 + i];
          float hi, lo;
          split_tf32(x.x, hi, lo); vh[4 * i] = __float_as_uint(hi); vl[4 * i] = __float_as_uint(lo);
          split_tf32(x.y, hi, lo); vh[4 * i + 1] = __float_as_uint(hi); vl[4 * i + 1] = __float_as_uint(lo);
          split_tf32(x.z, hi, lo); vh[4 * i + 2] = __float_as_uint(hi); vl[4 * i + 2] = __float_as_uint(lo);
          split_tf32(x.w, hi, lo); vh[4 * i + 3] = __float_as_uint(hi); vl[4 * i + 3] = __float_as_uint(lo);
        }
        tmem_st_32x32(lane_base + cQh + h * 32, vh);
        tmem_st_32x32(lane_base + cQl + h * 32, vl);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(q_ready);
    }
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    float m_run = -CUDART_INF_F, l_run = 0.f, alpha_prev = 0.f;
    long long* trace = ((blockIdx.x | blockIdx.y | blockIdx.z) == 0 && threadIdx.x == 0) ? g_attn_trace : nullptr;
    for (int j = 0; j <= ntiles; ++j) {
      float s[TKV];                                      // scores in the log2 domain (Q was scaled by log2(e)/8)
      float alpha = 0.f;
      if (j < ntiles) {
        // S(j) sits in columns 64 (j & 1): one of two 64-key buffers (KT = 64) or one half of the 128-key tile of
        // this key pair (KT = 128, one s_full barrier per pair)
        if (KT == 64 || (j & 1) == 0) mbar_wait(&s_full[KT == 64 ? (j & 1) : 0], (uint32_t)(j >> 1) & 1u);
        tcgen05_fence_after();
        if (trace) trace[j * 8 + 4] = clock64();
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_32x32(lane_base + cS0 + (j & 1) * 64 + h * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) s[h * 32 + i] = __uint_as_float(v[i]);
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&s_free[j & 1]);      // the tensor core may overwrite these 64 columns of S
        const int valid = nk - j * TKV;                  // keys of this tile that exist
        if (valid < TKV) {
#pragma unroll
          for (int i = 0; i < TKV; ++i)
            if (i >= valid) s[i] = -CUDART_INF_F;
        }
        float mx = s[0];
#pragma unroll
        for (int i = 1; i < TKV; ++i) mx = fmaxf(mx, s[i]);
        const float m_new = fmaxf(m_run, mx);
        alpha = ex2_approx(m_run - m_new);
        float rs0 = 0.f, rs1 = 0.f;
#pragma unroll
        for (int i = 0; i < TKV; i += 2) {
          s[i] = ex2_approx(s[i] - m_new);
          s[i + 1] = ex2_approx(s[i + 1] - m_new);
          rs0 += s[i];
          rs1 += s[i + 1];
        }
        l_run = l_run * alpha + (rs0 + rs1);
        m_run = m_new;
      }
      if (j > 0) {                                       // PV(j-1) is complete: P may be overwritten, O[(j-1)&1] is final
        mbar_wait(o_full, (uint32_t)(j - 1) & 1u);
        tcgen05_fence_after();
      }
      // P(j) goes out BEFORE PV(j-1) is folded: the PV -> P -> PV chain through these warps is what paces the loop
      // once the tensor pipe has slack, and the fold (two TMEM loads + 64 FMAs) need not be on it — PV is
      // double-buffered, PV(j) accumulates into the other buffer meanwhile.
      if (j < ntiles) {
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t ph_[32], pl_[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            float hi, lo;
            split_tf32(s[h * 32 + i], hi, lo);
            ph_[i] = __float_as_uint(hi);
            pl_[i] = __float_as_uint(lo);
          }
          tmem_st_32x32(lane_base + cPh + h * 32, ph_);
          tmem_st_32x32(lane_base + cPl + h * 32, pl_);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(p_full);
        if (trace) trace[j * 8 + 3] = clock64();
      }
      if (j > 0) {                                       // fold PV(j-1): O = O * alpha(j-1) + PV
        const uint32_t ocol = ((j - 1) & 1) ? cO1 : cO0;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t v[32];
          tmem_ld_32x32(lane_base + ocol + h * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) o[h * 32 + i] = fmaf(o[h * 32 + i], alpha_prev, __uint_as_float(v[i]));
        }
        tcgen05_fence_before();
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
  __syncthreads();
  if (warp == kWarpMma) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
    if (g_attn_trace && (blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0) {
      unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      g_attn_trace[40 * 8 + 1] = clock64();
      g_attn_trace[40 * 8 + 3] = (long long)gt;
    }
  }
}

}  // namespace

int set_attention16_trace(long long* dev_buf);
int set_attention_trace(long long* dev_buf) {
  GIMS_CUDA_OK(cudaMemcpyToSymbol(g_attn_trace, &dev_buf, sizeof(dev_buf)));
  return set_attention16_trace(dev_buf);
}

// planes as written by the qkv-mode GEMM epilogue (QkvPlanes, common.cuh)
int launch_attention_tc(const QkvPlanes& pl, float* out, const Segs& segs, int cross, cudaStream_t st) {
  int rows = segs_rows(segs);
  CUtensorMap mK, mV;
  // KT = 128 (GIMS_ATTN_KT=128) is correct but measured slower on B200 (56 vs 51 us per launch): the softmax warps need
  // ~1700 clk per 64-key tile, which is under the tensor pipe's 2170 clk at KT = 64 but not under its 1850 clk at 128.
  static const int kt = [] { const char* e = getenv("GIMS_ATTN_KT"); return (e && atoi(e) == 128) ? 128 : 64; }();
  GIMS_TRY(tc::make_tmap_f32_k32(&mK, static_cast<const float*>(pl.kp), 2 * (uint64_t)rows, kD, kD, kt));
  GIMS_TRY(tc::make_tmap_f32_k32(&mV, static_cast<const float*>(pl.vt), 2 * (uint64_t)kD, pl.ldv, pl.ldv, HD));
  AttnTcArgs a;
  a.qp = pl.qp;
  a.out = out;
  a.segs = segs;
  a.cross = cross;
  a.rows_total = rows;
  for (int i = 0; i < kMaxSegs; ++i) a.vbase[i] = pl.vbase[i];
  dim3 grid(cdiv(segs_nmax(segs), TQ), kHeads, segs.nseg);
  ProfScope prof(GIMS_PROF_ATTENTION, st);
  if (kt == 64) {
    GIMS_CUDA_OK(cudaFuncSetAttribute(k_attention_tc<64>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<64>::kSmem));
    k_attention_tc<64><<<grid, kAttnThreads, AttnCfg<64>::kSmem, st>>>(mK, mV, a);
  } else {
    GIMS_CUDA_OK(cudaFuncSetAttribute(k_attention_tc<128>, cudaFuncAttributeMaxDynamicSharedMemorySize, AttnCfg<128>::kSmem));
    k_attention_tc<128><<<grid, kAttnThreads, AttnCfg<128>::kSmem, st>>>(mK, mV, a);
  }
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}

}  // namespace gims
