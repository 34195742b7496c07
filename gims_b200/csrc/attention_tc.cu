// Dense multi-head attention on tcgen05 tensor cores, fp32-accurate (3xTF32), flash style.
//
// Replaces `attention` + the einsums of MultiHeadedAttention (models/gmatcher.py:35-39, 108-113) for
// head_dim 64:  out = softmax(q^T k / 8) v  per head, without materialising the (4, N, M) probabilities.
//
// Operands arrive as tf32 planes written by the QKV projection epilogue (gemm_tc.cu, qkv mode):
//   Qp [2][rows][256]   hi / lo planes of q * 1/8 (exact scaling), head-major channels h*64+d
//   Kp [2][rows][256]   hi / lo planes of k
//   Vt [2][256][ldv]    hi / lo planes of v, TRANSPOSED (channel-major) so that PV's B operand is K-major
// One CTA = 128 queries of one head of one image; 192 threads:
//   warp 0      TMA: Q tile once, then K / Vt tiles of 64 keys through a 2-stage ring
//   warp 1      one thread issues tcgen05.mma:  S = Q K^T (SS, 3 products),  PV = P V (TS: P read from TMEM)
//   warps 2..5  softmax: tcgen05.ld S -> online max / exp2 / sum in fp32 -> P split into tf32 hi / lo ->
//               tcgen05.st into TMEM -> after the PV MMAs: O = O * alpha + PV in registers (fp32, RN)
// TMEM columns: S_main 64 | S_corr 64 | P_hi 64 | P_lo 64 | PV_main 64 | PV_corr 64  (corr = the two small
// 3xTF32 products, kept apart because the tensor core rounds its accumulator toward zero).
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

// TMEM column offsets: two S buffers (main + corr each), P planes, PV (main + corr)
constexpr int cS0 = 0, cPh = 256, cPl = 320, cO = 384, cOc = 448;
__device__ __forceinline__ constexpr int cS(int buf) { return cS0 + buf * 128; }       // main; corr at +64

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

// Pipeline (per 64-key tile j):  QK(j+1) is issued before PV(j), so the tensor core computes the next score tile
// while the softmax warps work on tile j; the softmax warps fold PV(j-1) into their register accumulators while
// PV(j) / QK(j+1) run.  S is double-buffered in TMEM, P and PV are single-buffered (their reuse is ordered by
// p_full / o_full).
__global__ void __launch_bounds__(kAttnThreads, 1)
k_attention_tc(const __grid_constant__ CUtensorMap mapQ, const __grid_constant__ CUtensorMap mapK,
               const __grid_constant__ CUtensorMap mapVt, AttnTcArgs a) {
  extern __shared__ uint8_t smem_raw[];
  const int img = blockIdx.z, head = blockIdx.y;
  const int src = a.cross ? 1 - img : img;
  const int nq = seg_count(a.segs, img), nk = seg_count(a.segs, src);
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
  uint64_t* p_full = bars + 13;            // P written, PV(j-1) folded (4 arrivals)
  uint64_t* o_full = bars + 14;            // PV written (commit)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&mapQ);
    tma_prefetch_desc(&mapK);
    tma_prefetch_desc(&mapVt);
    mbar_init(q_full, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(&k_full[s], 1); mbar_init(&k_empty[s], 1);
      mbar_init(&v_full[s], 1); mbar_init(&v_empty[s], 1);
      mbar_init(&s_full[s], 1); mbar_init(&s_free[s], 4);
    }
    mbar_init(p_full, 4);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 1) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem = *tmem_slot;

  if (warp == 0) {
    if (lane == 0) {
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
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_tf32(TQ, TKV);      // M=128, N=64 for both S and PV
      const uint32_t qh = smem_u32(q_smem), ql = qh + 2 * kBoxBytesQ;
      auto issue_qk = [&](int j) {
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        mbar_wait(&k_full[s], ph);
        if (j >= 2) mbar_wait(&s_free[s], ph ^ 1);             // softmax warps have read S(j-2) out of this buffer
        tcgen05_fence_after();
        const uint32_t kh = smem_u32(k_smem + s * kKStageBytes), kl = kh + 2 * kBoxBytesKV;
        const uint32_t sm = tmem + cS(s), sc = sm + 64;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {                       // K-dim = 64 channels = 2 boxes x 4 k-steps
          uint32_t qoff = (ks >> 2) * kBoxBytesQ + (ks & 3) * 32;
          uint32_t koff = (ks >> 2) * kBoxBytesKV + (ks & 3) * 32;
          uint64_t dqh = umma_desc_sw128(qh + qoff), dql = umma_desc_sw128(ql + qoff);
          uint64_t dkh = umma_desc_sw128(kh + koff), dkl = umma_desc_sw128(kl + koff);
          umma_tf32_ss(sc, dql, dkh, idesc, ks ? 1u : 0u);
          umma_tf32_ss(sc, dqh, dkl, idesc, 1u);
          umma_tf32_ss(sm, dqh, dkh, idesc, ks ? 1u : 0u);
        }
        umma_commit(&k_empty[s]);
        umma_commit(&s_full[s]);
      };
      mbar_wait(q_full, 0);
      issue_qk(0);
      for (int j = 0; j < ntiles; ++j) {
        if (j + 1 < ntiles) issue_qk(j + 1);
        const int s = j & 1;
        const uint32_t ph = (uint32_t)(j >> 1) & 1u;
        mbar_wait(&v_full[s], ph);
        mbar_wait(p_full, j & 1);                              // P(j) in TMEM, PV(j-1) already folded away
        tcgen05_fence_after();
        const uint32_t vh = smem_u32(v_smem + s * kKStageBytes), vl = vh + 2 * kBoxBytesKV;
#pragma unroll
        for (int ks = 0; ks < 8; ++ks) {                       // K-dim = 64 keys = 2 Vt boxes x 4 k-steps; A from TMEM
          uint32_t voff = (ks >> 2) * kBoxBytesKV + (ks & 3) * 32;
          uint64_t dvh = umma_desc_sw128(vh + voff), dvl = umma_desc_sw128(vl + voff);
          umma_tf32_ts(tmem + cOc, tmem + cPl + ks * 8, dvh, idesc, ks ? 1u : 0u);
          umma_tf32_ts(tmem + cOc, tmem + cPh + ks * 8, dvl, idesc, 1u);
          umma_tf32_ts(tmem + cO, tmem + cPh + ks * 8, dvh, idesc, ks ? 1u : 0u);
        }
        umma_commit(&v_empty[s]);
        umma_commit(o_full);
      }
    }
  } else {
    // ===== softmax / accumulate warps: thread <-> query row =====
    const int q = warp & 3;
    const uint32_t lane_base = tmem + ((uint32_t)(32 * q) << 16);
    const int row = q0 + 32 * q + lane;
    float o[HD];
#pragma unroll
    for (int d = 0; d < HD; ++d) o[d] = 0.f;
    float m_run = -CUDART_INF_F, l_run = 0.f, alpha_prev = 0.f;
    const float kLog2e = 1.4426950408889634f;
    auto fold_pv = [&](float alpha) {                          // O = O * alpha + PV
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32], c[32];
        tmem_ld_32x32(lane_base + cO + h * 32, v);
        tmem_ld_32x32(lane_base + cOc + h * 32, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i)
          o[h * 32 + i] = fmaf(o[h * 32 + i], alpha, __uint_as_float(v[i]) + __uint_as_float(c[i]));
      }
    };
    for (int j = 0; j < ntiles; ++j) {
      const int sb = j & 1;
      mbar_wait(&s_full[sb], (uint32_t)(j >> 1) & 1u);
      tcgen05_fence_after();
      float s[TKV];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        uint32_t v[32], c[32];
        tmem_ld_32x32(lane_base + cS(sb) + h * 32, v);
        tmem_ld_32x32(lane_base + cS(sb) + 64 + h * 32, c);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) s[h * 32 + i] = (__uint_as_float(v[i]) + __uint_as_float(c[i])) * kLog2e;
      }
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(&s_free[sb]);           // the tensor core may overwrite this S buffer (tile j+2)
      const int valid = nk - j * TKV;                    // keys of this tile that exist
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int i = 0; i < TKV; ++i) {
        if (i >= valid) s[i] = -CUDART_INF_F;
        mx = fmaxf(mx, s[i]);
      }
      const float m_new = fmaxf(m_run, mx);
      const float alpha = exp2f(m_run - m_new);
      float rs = 0.f;
#pragma unroll
      for (int i = 0; i < TKV; ++i) { s[i] = exp2f(s[i] - m_new); rs += s[i]; }
      l_run = l_run * alpha + rs;
      m_run = m_new;
      if (j > 0) {                                       // PV(j-1) has to be out of TMEM before P(j) goes in
        mbar_wait(o_full, (j - 1) & 1);
        tcgen05_fence_after();
        fold_pv(alpha_prev);
      }
      alpha_prev = alpha;
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
    }
    mbar_wait(o_full, (ntiles - 1) & 1);
    tcgen05_fence_after();
    fold_pv(alpha_prev);
    if (row < nq) {
      const float inv = 1.f / l_run;
      float4* dst = reinterpret_cast<float4*>(a.out + (size_t)(a.segs.base[img] + row) * kD + head * HD);
#pragma unroll
      for (int d = 0; d < HD; d += 4) dst[d / 4] = make_float4(o[d] * inv, o[d + 1] * inv, o[d + 2] * inv, o[d + 3] * inv);
    }
    tcgen05_fence_before();
  }
  __syncthreads();
  if (warp == 1) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem);
  }
}

}  // namespace

// planes as written by the qkv-mode GEMM epilogue (QkvPlanes, common.cuh)
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
