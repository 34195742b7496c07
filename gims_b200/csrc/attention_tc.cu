// Dense multi-head attention on tcgen05 tensor cores, fp32-accurate (3xTF32), flash style.
//
// Replaces `attention` + the einsums of MultiHeadedAttention (models/gmatcher.py:35-39, 108-113) for
// head_dim 64:  out = softmax(q^T k / 8) v  per head, without materialising the (4, N, M) probabilities.
//
// Operands arrive as tf32 planes written by the QKV projection epilogue (gemm_tc.cu, qkv mode):
//   Qp [2][rows][256]   hi / lo planes of q * log2(e)/8 (scores land in the log2 domain), channels h*64+d
//   Kp [2][rows][256]   hi / lo planes of k
//   Vt [2][256][ldv]    hi / lo planes of v, TRANSPOSED (channel-major) so that PV's B operand is K-major
// One CTA = 128 queries of one head of one image; 192 threads:
//   warp 0      TMA: Q tile once, then K / Vt tiles of 64 keys through a 2-stage ring
//   warp 1      one thread issues tcgen05.mma:  S = Q K^T (SS, 3 products),  PV = P V (TS: P read from TMEM)
//   warps 2..5  softmax: tcgen05.ld S -> online max / exp2 / sum in fp32 -> P split into tf32 hi / lo ->
//               tcgen05.st into TMEM -> after the PV MMAs: O = O * alpha + PV in registers (fp32, RN)
#include <math_constants.h>

#include "common.cuh"
#include "tc_common.cuh"

namespace gims {

namespace {

using namespace tc;

constexpr int TQ = 128;        // queries per CTA (UMMA M)
constexpr int TKV = 64;        // keys per tile (UMMA N of S, K of PV)
constexpr int HD = 64;         // head dim
constexpr int kStagesKV = 2;   // K and Vt travel through separate 2-stage rings (K is released right after S = QK^T)
constexpr int kBoxBytesQ = TQ * 32 * 4;       // 16 KB: 128 rows x 32 floats
constexpr int kBoxBytesKV = TKV * 32 * 4;     // 8 KB: 64 rows x 32 floats
constexpr int kQBytes = 4 * kBoxBytesQ;       // hi{d0-31,d32-63}, lo{...}
constexpr int kKStageBytes = 4 * kBoxBytesKV;   // K hi(2) lo(2)   (same size for Vt)
constexpr int kAttnSmem = kQBytes + 2 * kStagesKV * kKStageBytes + 1024 + 256;
constexpr int kAttnThreads = 192;

// TMEM columns (all 512 in use): S[2] 64 each | P[2] = (hi 64 | lo 64) each | PV[2] 64 each.
// Each S / PV accumulator receives the 16 small correction products of its tile FIRST and the 8 hi*hi products
// last: the tensor core rounds the accumulator toward zero at every step, and this order keeps those roundings at
// the magnitude of the small terms for as long as possible (same error as a separate correction accumulator,
// without its TMEM columns and the extra add).
__device__ __forceinline__ constexpr int cS(int b) { return b * 64; }
__device__ __forceinline__ constexpr int cPh(int b) { return 128 + b * 128; }
__device__ __forceinline__ constexpr int cPl(int b) { return 192 + b * 128; }
__device__ __forceinline__ constexpr int cO(int b) { return 384 + b * 64; }

// Optional pipeline trace (bring-up / profiling): CTA (0,0,0) stores clock64() stamps per tile.
//   [j*8+0] MMA: QK(j+1) issue start   [j*8+1] MMA: PV(j) operands ready   [j*8+2] MMA: PV(j) issued
//   [j*8+4] softmax warp 2: S(j) observed   [j*8+5] P(j) handed over   [j*8+6] fold of PV(j-1) done
__device__ long long* g_attn_trace = nullptr;
#define ATTN_TRACE(slot)                                                                       \
  do {                                                                                         \
    if (trace) trace[j * 8 + (slot)] = clock64();                                              \
  } while (0)

struct AttnTcArgs {
  float* out;                 // [rows][256]
  Segs segs;
  int cross;
  int rows_total;             // plane stride of Qp/Kp in rows
  int vbase1;                 // first Vt key column of image 1
};

__device__ __forceinline__ void tmem_st_32x32(uint32_t taddr, const uint32_t (&r)[32]) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// D[tmem] (+)= A[tmem] * B[smem]^T
__device__ __forceinline__ void umma_tf32_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], [%1], %2, %3, p;\n\t}" ::"r"(tmem_d),
      "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 2^x for x <= 0 (softmax arguments): one MUFU, flushes results below 2^-126 to zero
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}

// Pipeline (per 64-key tile j):  QK(j+1) is issued before PV(j), so the tensor core computes the next score tile
// while the softmax warps work on tile j; the softmax warps fold PV(j-1) into their register accumulators after
// they have handed P(j) to the tensor core.  S, P and PV are all double-buffered in TMEM.
__global__ void __launch_bounds__(kAttnThreads, 1)
k_attention_tc(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
               const __grid_constant__ CUtensorMap mapVt, AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const int img = blockIdx.z, head = blockIdx.y;
  const int src = a.cross ? 1 - img : img;
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
  uint8_t* q_smem = smem;
  uint8_t* k_smem = smem + kQBytes;
  uint8_t* v_smem = k_smem + kStagesKV * kKStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(v_smem + kStagesKV * kKStageBytes);
  uint64_t* q_full = bars;                 // 1
  uint64_t* k_full = bars + 1;             // [2] TMA -> MMA
  uint64_t* k_empty = bars + 3;            // [2] QK MMAs retired (tcgen05.commit)
  uint64_t* v_full = bars + 5;             // [2]
  uint64_t* v_empty = bars + 7;            // [2] PV MMAs retired
  uint64_t* s_full = bars + 9;             // [2] S buffer written (commit)
  uint64_t* s_free = bars + 11;            // [2] S buffer read by the softmax warps (4 arrivals)
  uint64_t* p_full = bars + 13;            // [2] P buffer written by the softmax warps (4 arrivals)
  uint64_t* o_full = bars + 15;            // [2] PV buffer written (commit); also: P buffer consumed
  uint64_t* o_free = bars + 17;            // [2] PV buffer folded by the softmax warps (4 arrivals)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 19);

  // Warp roles: 0..3 = data warps (TMEM lane quarter = warp), 4 = TMA producer, 5 = MMA issuer.  The issuer gets
  // the highest warp id on its scheduler: the arbiter favours high warp ids, and a starved issuer starves the tensor pipe.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpTma = 4, kWarpMma = 5;
  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapVt);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&s_free[s], 4);
      mbar_init(&p_full[s], 4); mbar_init(&o_full[s], 1); mbar_init(&o_free[s], 4);
    }
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = __shfl_sync(0xffffffffu, *tmem_slot, 0);      // warp-uniform for the compiler

  if (threadIdx.x == kWarpTma * 32) {
    {
      // Q planes: rows [qrow0, +128) of plane 0 (hi) and plane 1 (lo, row offset rows_total), channels head*64..+64
      mbar_arrive_expect_tx(q_full, kQBytes);
      for (int pl = 0; pl < 2; ++pl)
        for (int hf = 0; hf < 2; ++hf)
          tma_load_2d(q_smem + (pl * 2 + hf) * kBoxBytesQ, &mapQ, q_full, head * HD + hf * 32, pl * a.rows_total + qrow0);
      for (int j = 0; j < ntiles; ++j) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        const int key0 = krow0 + j * TKV;                       // row in the K planes
        const int vcol0 = (src ? a.vbase1 : 0) + j * TKV;       // key column in the Vt planes (multiple of 64)
        mbar_wait(&k_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[s], kKStageBytes);
        for (int pl = 0; pl < 2; ++pl)
          for (int hf = 0; hf < 2; ++hf)      // K box: 64 keys x 32 channels
            tma_load_2d(k_smem + s * kKStageBytes + (pl * 2 + hf) * kBoxBytesKV, &mapK, &k_full[s], head * HD + hf * 32,
                        pl * a.rows_total + key0);
        mbar_wait(&v_empty[s], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[s], kKStageBytes);
        for (int pl = 0; pl < 2; ++pl)
          for (int hf = 0; hf < 2; ++hf)      // Vt box: 64 channels x 32 keys
            tma_load_2d(v_smem + s * kKStageBytes + (pl * 2 + hf) * kBoxBytesKV, &mapVt, &v_full[s], vcol0 + hf * 32,
                        pl * kD + head * HD);
      }
    }
  } else if (threadIdx.x == kWarpMma * 32) {
    // ===== MMA issuer: ONE thread, written so that all operand arithmetic stays in the uniform datapath
    // (descriptor = base + constant; no arrays, no lambdas) — an issuer that needs R2UR moves per operand
    // cannot keep the tensor pipe fed.
    constexpr uint32_t idesc = umma_idesc_tf32(TQ, TKV);      // M=128, N=64 for both S and PV
    const uint64_t dq_hi = umma_desc_sw128(smem_u32(q_smem)), dq_lo = dq_hi + ((2 * kBoxBytesQ) >> 4);
    const uint64_t dk0 = umma_desc_sw128(smem_u32(k_smem)), dv0 = umma_desc_sw128(smem_u32(v_smem));
    long long* trace = (blockIdx.x | blockIdx.y | blockIdx.z) == 0 ? g_attn_trace : nullptr;
    mbar_wait(q_full, 0);
    // software pipeline: iteration jq issues QK(jq) and then PV(jq-1)
    for (int jq = 0; jq <= ntiles; ++jq) {
      const int j = jq - 1;                                    // tile whose PV is issued in this iteration
      if (trace && j >= 0) trace[j * 8 + 0] = clock64();
      if (jq < ntiles) {
        const int s = jq & 1;
        const uint32_t ph = (uint32_t)(jq >> 1) & 1u;
        mbar_wait(&k_full[s], ph);
        if (jq >= 2) mbar_wait(&s_free[s], ph ^ 1);            // softmax warps have read S(jq-2) out of this buffer
        tcgen05_fence_after();
        const uint64_t kh = dk0 + (uint64_t)(s * (kKStageBytes >> 4)), kl = kh + ((2 * kBoxBytesKV) >> 4);
        const uint32_t sacc = tmem + s * 64;
        // K-dim = 64 channels = 2 boxes x 4 k-steps; corrections first, hi*hi last
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t qoff = ((ks >> 2) * kBoxBytesQ + (ks & 3) * 32) >> 4, koff = ((ks >> 2) * kBoxBytesKV + (ks & 3) * 32) >> 4;
          umma_tf32_ss(sacc, dq_lo + qoff, kh + koff, idesc, ks ? 1u : 0u);
          umma_tf32_ss(sacc, dq_hi + qoff, kl + koff, idesc, 1u);
        }
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {
          const uint64_t qoff = ((ks >> 2) * kBoxBytesQ + (ks & 3) * 32) >> 4, koff = ((ks >> 2) * kBoxBytesKV + (ks & 3) * 32) >> 4;
          umma_tf32_ss(sacc, dq_hi + qoff, kh + koff, idesc, 1u);
        }
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[s]);
      }
      if (j >= 0) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        mbar_wait(&v_full[s], ph);
        mbar_wait(&p_full[s], ph);                             // P(j) is in TMEM
        if (j >= 2) mbar_wait(&o_free[s], ph ^ 1);             // PV(j-2) has been folded out of this buffer
        tcgen05_fence_after();
        if (trace) trace[j * 8 + 1] = clock64();
        const uint64_t vh = dv0 + (uint64_t)(s * (kKStageBytes >> 4)), vl = vh + ((2 * kBoxBytesKV) >> 4);
        const uint32_t oacc = tmem + 384 + s * 64, p_hi = tmem + 128 + s * 128, p_lo = p_hi + 64;
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
        umma_commit(&v_empty[s]);
        umma_commit(&o_full[s]);
        if (trace) trace[j * 8 + 2] = clock64();
      }
    }
  } else if (warp < 4) {
    // ===== softmax / accumulate warps: thread <-> query row =====
    const int q = warp & 3;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);
    const int row = q0 + 32 * q + lane;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    float m_run = -CUDART_INF_F, l_run = 0.f, alpha_prev = 0.f;
    auto fold_pv = [&](int jj, float alpha) {                  // O = O * alpha + PV(jj)
      const int b = jj & 1;
      mbar_wait(&o_full[b], (uint32_t)(jj >> 1) & 1u);
      tcgen05_fence_after();
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_32x32(lane_base + cO(b) + h * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) o[h * 32 + i] = fmaf(o[h * 32 + i], alpha, __uint_as_float(v[i]));
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&o_free[b]);
    };
    long long* trace = ((blockIdx.x | blockIdx.y | blockIdx.z) == 0 && lane == 0) ? g_attn_trace : nullptr;
    const int tw = warp;   // trace: slots 4 (S seen, warp 0), 3/5/6/7 (P handed over by warps 0..3)
    for (int j = 0; j < ntiles; ++j) {
      const int sb = j & 1;
      const uint32_t ph = (uint32_t)(j >> 1) & 1u;
      mbar_wait(&s_full[sb], ph);
      tcgen05_fence_after();
      if (tw == 0) ATTN_TRACE(4);
      float s[TKV];                                      // scores in the log2 domain (Q was scaled by log2(e)/8)
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32];
        tmem_ld_32x32(lane_base + cS(sb) + h * 32, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[h * 32 + i] = __uint_as_float(v[i]);
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[sb]);           // the tensor core may overwrite this S buffer (tile j+2)
      const int valid = nk - j * TKV;                    // keys of this tile that exist
      if (valid < TKV) {
#pragma unroll
        for (int i = 0; i < TKV; ++i)
          if (i >= valid) s[i] = -CUDART_INF_F;
      }
      float mx = s[0];
#pragma unroll
      for (int i = 1; i < TKV; ++i) mx = fmaxf(mx, s[i]);
      const float m_new = fmaxf(m_run, mx);
      const float alpha = ex2_approx(m_run - m_new);
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
      // P(j) -> TMEM buffer sb; that buffer was last read by PV(j-2), whose completion we observed when folding it
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
        tmem_st_32x32(lane_base + cPh(sb) + h * 32, ph_);
        tmem_st_32x32(lane_base + cPl(sb) + h * 32, pl_);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&p_full[sb]);
      if (tw == 0) ATTN_TRACE(3);
      if (j > 0) fold_pv(j - 1, alpha_prev);             // overlaps PV(j) / QK(j+1) on the tensor core
      alpha_prev = alpha;
    }
    fold_pv(ntiles - 1, alpha_prev);
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
  }
}

}  // namespace

// planes as written by the qkv-mode GEMM epilogue (QkvPlanes, common.cuh)
int set_attention_trace(long long* dev_buf) {
  GIMS_CUDA_OK(cudaMemcpyToSymbol(g_attn_trace, &dev_buf, sizeof(dev_buf)));
  return GIMS_OK;
}

int launch_attention_tc(const QkvPlanes& pl, float* out, int n0_max, int n1_max, const int* n_dev, int cross,
                        cudaStream_t st) {
  int rows = n0_max + n1_max;
  CUtensorMap mQ, mK, mV;
  GIMS_TRY(tc::make_tmap_f32_k32(&mQ, pl.qp, 2 * (uint64_t)rows, kD, kD, TQ));
  GIMS_TRY(tc::make_tmap_f32_k32(&mK, pl.kp, 2 * (uint64_t)rows, kD, kD, TKV));
  GIMS_TRY(tc::make_tmap_f32_k32(&mV, pl.vt, 2 * (uint64_t)kD, pl.ldv, pl.ldv, HD));
  AttnTcArgs a;
  a.out = out;
  a.segs.base[0] = 0; a.segs.base[1] = n0_max; a.segs.nmax[0] = n0_max; a.segs.nmax[1] = n1_max; a.segs.n_dev = n_dev;
  a.segs.nseg = 2;
  a.cross = cross;
  a.rows_total = rows;
  a.vbase1 = pl.vbase1;
  GIMS_CUDA_OK(cudaFuncSetAttribute(k_attention_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem));
  int nmax = n0_max > n1_max ? n0_max : n1_max;
  ProfScope prof(GIMS_PROF_ATTENTION, st);
  k_attention_tc<<<dim3(cdiv(nmax, TQ), kHeads, 2), kAttnThreads, kAttnSmem, st>>>(mQ, mK, mV, a);
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}

}  // namespace gims
