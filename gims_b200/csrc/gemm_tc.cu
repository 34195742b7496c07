// fp32-accurate GEMM on the 5th-generation tensor cores (tcgen05 / TMEM / TMA), 3xTF32 error compensation.
//
//   Y[r][o] = epi( sum_k A(r,k) * W[o][k] + bias[o] )        A, W K-major — the same contract as gemm_simt.cu
//
// (Conv1d(k=1) / Linear of models/gmatcher.py:11-24, 105-125, 273 and the SAGE linears.)  Every fp32 operand x
// is split x = hi + lo with hi = tf32(x); the accumulator receives A_hi*W_hi + A_lo*W_hi + A_hi*W_lo in fp32
// (TMEM), i.e. ~2^-21 relative accuracy per product instead of TF32's 2^-11.
//
// Structure (one CTA = one 128 x BN output tile, 224 threads; DESIGN.md section 5 has the measurements behind it):
//   warp 4      TMA producer: raw fp32 A tile, W_hi tile, W_lo tile per k-block (32 floats = one 128-B swizzle row)
//   warps 0..3  thread = tile row = TMEM lane: read the row of the raw A tile from shared memory, split it into tf32
//               hi / lo and tcgen05.st both into a ring of TMEM slots; afterwards the epilogue: tcgen05.ld, sum of the
//               accumulators, transpose through a staging tile, bias / residual / ReLU (or the Q / K / Vt planes of the
//               QKV projection, or the scaled couplings), 128-bit stores that cover whole lines
//   warps 5, 6  one MMA issuer thread each (even / odd k-blocks, own accumulator): three TS-form tcgen05.mma per
//               8-wide k-step (A from TMEM, W from shared memory), tcgen05.commit to free the W stage and the A slot
// W_hi / W_lo are split once at weight-pack time; activations are split on the fly, so no tensor in HBM
// changes layout.  Pipeline: full[s] (TMA bytes) -> conv[slot] (A in TMEM) -> MMA -> empty[s] / tfree[slot].
#include "common.cuh"
#include "tc_common.cuh"

namespace gims {

namespace tc {

EncodeTiledFn get_encode_tiled() {
  static EncodeTiledFn fn = nullptr;
  if (fn) return fn;
  void* p = nullptr;
  cudaDriverEntryPointQueryResult qres;
  if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess ||
      qres != cudaDriverEntryPointSuccess)
    return nullptr;
  fn = reinterpret_cast<EncodeTiledFn>(p);
  return fn;
}

int make_tmap_f32_k32(CUtensorMap* out, const float* base, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return GIMS_ERR_CUDA; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * sizeof(float)};
  cuuint32_t box[2] = {32, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, const_cast<float*>(base), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled failed (%d) base=%p rows=%llu cols=%llu ld=%llu box_rows=%u", (int)r, (const void*)base,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
    return GIMS_ERR_CUDA;
  }
  return GIMS_OK;
}

int make_tmap_16_k64(CUtensorMap* out, const void* base, int fmt, uint64_t rows, uint64_t cols, uint64_t ld, uint32_t box_rows) {
  EncodeTiledFn enc = get_encode_tiled();
  if (!enc) { set_error("cuTensorMapEncodeTiled entry point not available"); return GIMS_ERR_CUDA; }
  cuuint64_t dims[2] = {cols, rows};
  cuuint64_t strides[1] = {ld * 2};
  cuuint32_t box[2] = {64, box_rows};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(out, fmt == 0 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT16 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base),
                   dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                   CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    set_error("cuTensorMapEncodeTiled (16-bit) failed (%d) base=%p rows=%llu cols=%llu ld=%llu box_rows=%u", (int)r, base,
              (unsigned long long)rows, (unsigned long long)cols, (unsigned long long)ld, box_rows);
    return GIMS_ERR_CUDA;
  }
  return GIMS_OK;
}

}  // namespace tc

namespace {

using namespace tc;

constexpr int BM = 128, BK = 32;
constexpr int kThreadsTc = 224;   // 4 data warps + TMA warp + two MMA issuer warps
// PREC 1 adds a second group of 4 data warps (warps 7..10): converting a 64-deep k-block of the A tile to fp16 hi / lo takes
// one thread per row ~1000 clk against 770 clk of MMAs; two threads per row (one per 32-float box) put the loop back on the
// tensor pipe, and the two groups share the epilogue chunks.
__host__ __device__ constexpr int tc_threads(int prec) { return prec ? kThreadsTc + 128 : kThreadsTc; }

// PREC 0: 3xTF32 (operands split into tf32 hi / lo, k-blocks of 32).  PREC 1: fp16 hi + lo (the same error class at twice the
// tensor-pipe rate, see attention_f16.cu): k-blocks of 64 — the raw fp32 A tile is two 32-float boxes, the W planes are 16-bit
// tiles with 128-byte rows (64 elements), K = 16 per MMA.
template <int BN, int PREC>
struct TcCfg {
  static constexpr int kBlockK = PREC ? 64 : 32;         // k-block depth in elements
  static constexpr int kABytes = BM * kBlockK * 4;       // raw fp32 A tile: 16 KB per 32-float box
  static constexpr int kWBytes = PREC ? BN * 64 * 2 : BN * BK * 4;
  static constexpr int kStageBytes = kABytes + 2 * kWBytes;
  static constexpr int kStages = PREC ? (BN <= 64 ? 2 : (BN <= 128 ? 3 : 2))
                                      : ((BN <= 64) ? 3 : (BN <= 128 ? 4 : 3));      // BN = 64: <= 98 KB, two CTAs per SM
  // Two issuer threads (even / odd k-blocks), each with its own accumulators: a single issuer leaves the tensor pipe
  // idle for ~40 % of the time around its commits (profiles/r01_gemm_pipeline_trace.txt).
  static constexpr int kIssuers = 2;
  // The tensor core rounds the fp32 accumulator toward zero on every accumulation, so the error of one accumulation
  // chain grows linearly with its length.  Splitting the k-blocks over two accumulators halves the chains; the epilogue
  // adds the accumulators in fp32 (RN).
  static constexpr bool kFold = true;
  static constexpr int kMainAcc = 2;
  static constexpr int kCorrAcc = kFold ? 0 : kIssuers;
  static constexpr int kAccCols = (kMainAcc + kCorrAcc) * BN;
  // The A operand lives in TMEM: a ring of k-block slots, 32 columns A_hi + 32 columns A_lo each (tf32: 32 values per
  // plane, fp16: 64 values packed two per column).
  static constexpr int kTmemCols = (BN <= 64) ? 256 : 512;                  // BN = 64: two CTAs per SM
  static constexpr int kRing = (kTmemCols - kAccCols) / 64 >= 4 ? 4 : 2;
  static constexpr int kRingCol = kAccCols;
  static_assert(kAccCols + kRing * 64 <= kTmemCols, "accumulators + A ring exceed TMEM");
  static constexpr int kBarriers = 2 * kStages + 2 * kRing + 1;
  static constexpr int kBiasOff = (kBarriers * 8 + 16 + 15) & ~15;       // after the barriers and the TMEM address slot
  static constexpr int kSmemBytes = kStages * kStageBytes + 1024 /*align*/ + kBiasOff + BN * 4;
};

struct TcArgs {
  int K0, K1;                    // k extents served by tensor maps A0 / A1 (multiples of 32)
  const float* bias;
  const float* R; int ldr;
  float* Y; int ldy;
  int N;                         // live output columns (score mode: from n_dev[1])
  int relu;
  float scale;                   // score mode
  int score;                     // 1: Y is the couplings matrix (row-major, ldy = coup_ld(n1_max)), W rows = image-1 rows
  Segs segs;
  int tile_end[kMaxSegs];        // cumulative number of row tiles up to and including segment s
  // qkv mode (N = 768): instead of Y the epilogue writes the tf32 planes the attention kernel consumes:
  //   columns [0,256)   -> Qp [rows_total][256]     fp32 (acc + bias) * log2(e)/8
  //   columns [256,512) -> Kp [2][rows_total][256]
  //   columns [512,768) -> Vt [2][256][ldv]          transposed; rows beyond the live count are zero-filled
  int qkv;
  float* qp; void* kp; void* vt; int ldv; int rows_total; int vbase[kMaxSegs];
  unsigned* status;              // 16-bit planes / fp16 operands: overflow flag
  const float* wscale_inv;       // PREC 1: the W planes hold W * 2^e; the accumulators are multiplied by *wscale_inv = 2^-e
  int single_chain;              // persistent kernel, experiment knob (GIMS_GEMM_SINGLE_CHAIN=1)
};

// Optional pipeline trace (bring-up / profiling): CTA (0,0) stores clock64() stamps, see tools/gemm_trace.py.
__device__ long long* g_gemm_trace = nullptr;

// Where a tile sits: row segment, live rows / columns, first row / column of the tile, first row of the segment.
struct TileCoord { int seg, rows, ncols, r0, c0, rbase; };

// The epilogue of one 128 x BN tile for the warp (lane quarter q, column-chunk group grp of kGroups): tcgen05.ld of the
// accumulator chunks at `lane_base` (n_main / n_corr accumulators BN columns apart), bias / residual / ReLU or the QKV
// planes or the couplings, stores.  `ready()` is called once the residual prefetch has been issued and must return
// when the accumulators may be read.  `stg`: this warp's private 32 x 36-float staging tile in shared memory.
template <int BN, int MODE, int PREC, typename Ready>
__device__ __forceinline__ void epilogue_tile(const TcArgs& g, const TileCoord& tcd, uint32_t lane_base, int n_main, int n_corr,
                                              int grp, int kGroups, int q, int lane, uint32_t stg, const float* bias_s,
                                              Ready&& ready, long long* trace, bool tr0) {
  using Cfg = TcCfg<BN, PREC>;
  const int seg = tcd.seg, rows = tcd.rows, ncols = tcd.ncols, r0 = tcd.r0, c0 = tcd.c0, rbase = tcd.rbase;
  const float wsc = (PREC && g.wscale_inv) ? *g.wscale_inv : 1.f;
  // Each warp transposes its 32x32 chunks through a private staging tile in the (now idle) pipeline stages, so that
  // one store instruction covers four complete 128-byte lines of Y and the residual is read the same way.  An SM
  // takes ~30 bytes of stores per clock (tools/micro/store_rate.cu), which is what this loop should be bound by:
  // all shared-memory reads of a chunk are issued before the first store, and MODE is a template parameter, so
  // there is no branch or address arithmetic between the stores.
  constexpr int kStgLd = 36;                         // floats; 16-byte aligned rows, conflict-free both ways
  const int rl0 = lane >> 3, cj = (lane & 7) * 4;    // read-back: row i*4 + rl0 of the chunk, columns cj..cj+3
  const int rfirst = r0 + 32 * q + rl0;              // tile row of read-back slot i = 0; slot i is 4*i further
  const int nvalid = rows - rfirst;                  // slot i is a live row iff 4*i < nvalid
  const bool use_r = MODE == 0 && g.R != nullptr;
  const float* rrow = use_r ? g.R + (size_t)(rbase + rfirst) * g.ldr + c0 + cj : nullptr;
  float4 rres[8];                                    // residual of the current chunk, fetched one chunk ahead
#pragma unroll
  for (int i = 0; i < 8; ++i)
    rres[i] = (use_r && 4 * i < nvalid && c0 + 32 * grp + cj < ncols) ? *reinterpret_cast<const float4*>(rrow + (size_t)(4 * i) * g.ldr + 32 * grp)
                                                           : make_float4(0.f, 0.f, 0.f, 0.f);
  ready();                                           // the residual's latency hides behind the tail of the k-loop
  if (trace && tr0) trace[3] = clock64();
#pragma unroll 1
  for (int cc = grp; cc < BN / 32; cc += kGroups) {
    uint32_t v[32];
    const uint32_t lane_col = lane_base + (uint32_t)(cc * 32);
    tmem_ld_32x32(lane_col, v);
    tmem_ld_wait();
    for (int a = 1; a < Cfg::kMainAcc + Cfg::kCorrAcc; ++a) {
      if (a < Cfg::kMainAcc ? a >= n_main : a - Cfg::kMainAcc >= n_corr) continue;   // never written (short K)
      uint32_t w[32];
      tmem_ld_32x32(lane_col + (uint32_t)(a * BN), w);       // a >= kMainAcc: the correction accumulators
      tmem_ld_wait();
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(w[j]));
    }
    if (PREC) {
#pragma unroll
      for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * wsc);
    }
    if (trace && tr0 && cc < 2) trace[56 + 3 * cc] = clock64();
    const int c = c0 + cc * 32;
    if (c >= ncols) continue;                                // warp-uniform
    if (MODE >= 2 && c >= 2 * kD) {
      // V: key column of this row in Vt; image 1 starts at a 64-aligned column (TMA box starts must be 16-byte
      // aligned in global memory).  Lanes = consecutive rows = consecutive addresses: coalesced as it is.
      // Rows between the live count and the next multiple of 64 are ZERO-FILLED: the attention kernel multiplies them
      // by P = 0, which must not meet a NaN / Inf left over from whatever used this workspace before.
      const int r = r0 + 32 * q + lane;
      if (r < ((g.segs.nmax[seg] + 63) & ~63)) {             // never touch the other image's columns
        const bool row_ok = r < rows;
        const int cc256 = c & 255;
        const size_t kcol = (size_t)g.vbase[seg] + r;
        if (MODE == 2) {
          float* hi = static_cast<float*>(g.vt) + (size_t)cc256 * g.ldv + kcol;
          float* lo = hi + (size_t)kD * g.ldv;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float h, l;
            split_tf32(row_ok ? __uint_as_float(v[j]) + bias_s[cc * 32 + j] : 0.f, h, l);
            hi[(size_t)j * g.ldv] = h;
            lo[(size_t)j * g.ldv] = l;
          }
        } else {
          constexpr int FMT = MODE == 4 ? 1 : 0;
          uint16_t* hi = static_cast<uint16_t*>(g.vt) + (size_t)cc256 * g.ldv + kcol;
          uint16_t* lo = hi + (size_t)kD * g.ldv;
          bool ovf = false;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const float x = row_ok ? __uint_as_float(v[j]) + bias_s[cc * 32 + j] : 0.f;
            if (FMT == 0) ovf = ovf || fabsf(x) >= 32768.f;
            uint32_t wh, wl;
            if (MODE == 3) split16x2<FMT>(x, 0.f, wh, wl); else wh = pack16<FMT>(x, 0.f);
            hi[(size_t)j * g.ldv] = (uint16_t)wh;
            if (MODE == 3) lo[(size_t)j * g.ldv] = (uint16_t)wl;
          }
          if (ovf && g.status) atomicOr(g.status, GIMS_STATUS_FP16_RANGE);
        }
      }
      continue;
    }
    __syncwarp();                                            // the previous chunk has been read back
#pragma unroll
    for (int j = 0; j < 32; j += 4)
      asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(stg + (uint32_t)((lane * kStgLd + j) * 4)), "r"(v[j]),
                   "r"(v[j + 1]), "r"(v[j + 2]), "r"(v[j + 3])
                   : "memory");
    __syncwarp();
    if (trace && tr0 && cc < 2) trace[57 + 3 * cc] = clock64();
    if (MODE == 1) {
      // couplings rows are n1_max + 1 floats long (not 16-byte aligned) and end raggedly: scalar, one row per store
      if (c + lane < ncols) {
        float* yp = g.Y + (size_t)(r0 + 32 * q) * g.ldy + c + lane;
        const int nrow = min(32, rows - (r0 + 32 * q));
#pragma unroll 8
        for (int i = 0; i < 32; ++i) {
          float x;
          asm volatile("ld.shared.f32 %0, [%1];" : "=f"(x) : "r"(stg + (uint32_t)((i * kStgLd + lane) * 4)) : "memory");
          if (i < nrow) yp[(size_t)i * g.ldy] = x * g.scale;
        }
      }
      continue;
    }
    const float4 bias4 = *reinterpret_cast<const float4*>(bias_s + cc * 32 + cj);
    float4 o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];"
                   : "=f"(o[i].x), "=f"(o[i].y), "=f"(o[i].z), "=f"(o[i].w)
                   : "r"(stg + (uint32_t)(((i * 4 + rl0) * kStgLd + cj) * 4))
                   : "memory");
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      o[i].x += bias4.x + rres[i].x; o[i].y += bias4.y + rres[i].y;
      o[i].z += bias4.z + rres[i].z; o[i].w += bias4.w + rres[i].w;
    }
    if (use_r && cc + kGroups < BN / 32 && c + 32 * kGroups < ncols) {       // my next chunk's residual, one chunk ahead
#pragma unroll
      for (int i = 0; i < 8; ++i)
        if (4 * i < nvalid) rres[i] = *reinterpret_cast<const float4*>(rrow + (size_t)(4 * i) * g.ldr + (cc + kGroups) * 32);
    }
    if (MODE >= 2) {
      if (c < kD) {
        // Q: one fp32 plane, scaled by log2(e)/sqrt(64) (scores in the log2 domain); the attention kernel splits
        // its own query rows when it loads them into TMEM
        float* qd = g.qp + (size_t)(rbase + rfirst) * kD + c + cj;
        const float sc = 0.18033688011112042f;
#pragma unroll
        for (int i = 0; i < 8; ++i)
          if (4 * i < nvalid)
            *reinterpret_cast<float4*>(qd + (size_t)(4 * i) * kD) =
                make_float4(o[i].x * sc, o[i].y * sc, o[i].z * sc, o[i].w * sc);
      } else if (MODE == 2) {
        float* hi = static_cast<float*>(g.kp) + (size_t)(rbase + rfirst) * kD + (c & 255) + cj;
        const size_t plane = (size_t)g.rows_total * kD;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          float4 h, l;
          split_tf32(o[i].x, h.x, l.x); split_tf32(o[i].y, h.y, l.y);
          split_tf32(o[i].z, h.z, l.z); split_tf32(o[i].w, h.w, l.w);
          if (4 * i < nvalid) {
            *reinterpret_cast<float4*>(hi + (size_t)(4 * i) * kD) = h;
            *reinterpret_cast<float4*>(hi + (size_t)(4 * i) * kD + plane) = l;
          }
        }
      } else {
        // K: 16-bit planes, four channels = one 8-byte store per plane
        constexpr int FMT = MODE == 4 ? 1 : 0;
        uint16_t* hi = static_cast<uint16_t*>(g.kp) + (size_t)(rbase + rfirst) * kD + (c & 255) + cj;
        const size_t plane = (size_t)g.rows_total * kD;
        bool ovf = false;
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          uint2 h, l;
          if (FMT == 0 && 4 * i < nvalid)
            ovf = ovf || fmaxf(fmaxf(fabsf(o[i].x), fabsf(o[i].y)), fmaxf(fabsf(o[i].z), fabsf(o[i].w))) >= 32768.f;
          if (MODE == 3) {
            split16x2<FMT>(o[i].x, o[i].y, h.x, l.x);
            split16x2<FMT>(o[i].z, o[i].w, h.y, l.y);
          } else {
            h.x = pack16<FMT>(o[i].x, o[i].y);
            h.y = pack16<FMT>(o[i].z, o[i].w);
          }
          if (4 * i < nvalid) {
            *reinterpret_cast<uint2*>(hi + (size_t)(4 * i) * kD) = h;
            if (MODE == 3) *reinterpret_cast<uint2*>(hi + (size_t)(4 * i) * kD + plane) = l;
          }
        }
        if (ovf && g.status) atomicOr(g.status, GIMS_STATUS_FP16_RANGE);
      }
    } else {
      float* yp = g.Y + (size_t)(rbase + rfirst) * g.ldy + c + cj;
      const bool relu = g.relu != 0;
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float4 y = o[i];
        if (relu) { y.x = fmaxf(y.x, 0.f); y.y = fmaxf(y.y, 0.f); y.z = fmaxf(y.z, 0.f); y.w = fmaxf(y.w, 0.f); }
        if (4 * i < nvalid) *reinterpret_cast<float4*>(yp + (size_t)(4 * i) * g.ldy) = y;
      }
    }
    if (trace && tr0 && cc < 2) trace[58 + 3 * cc] = clock64();
  }
}

// MODE 0: Y = act(acc + bias + R);  1: couplings (score GEMM);  QKV projection: 2: Q / K / Vt tf32 planes,
// 3: fp16 hi + lo planes, 4: one bf16 plane (K / Vt; Q stays fp32)
template <int BN, int MODE, int PREC>
__global__ void __launch_bounds__(tc_threads(PREC), (BN <= 64 && !PREC) ? 2 : 1)
k_gemm_tc(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
          const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo, TcArgs g) {
  using Cfg = TcCfg<BN, PREC>;
  extern __shared__ uint8_t smem_raw[];
  long long* trace = (blockIdx.x | blockIdx.y) == 0 && (threadIdx.x & 31) == 0 ? g_gemm_trace : nullptr;
  if (trace && threadIdx.x == 0) {
    trace[0] = clock64();
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    trace[6] = (long long)gt;
  }
  // ---- tile coordinates (uniform per CTA) -------------------------------------------------------
  int seg, tile;
  if (MODE == 1) { seg = 0; tile = blockIdx.y; }
  else {
    seg = 0;
    while (seg + 1 < g.segs.nseg && (int)blockIdx.y >= g.tile_end[seg]) ++seg;
    tile = blockIdx.y - (seg ? g.tile_end[seg - 1] : 0);
  }
  const int rows = __shfl_sync(0xffffffffu, seg_count(g.segs, seg), 0);
  const int ncols = MODE == 1 ? __shfl_sync(0xffffffffu, seg_count(g.segs, 1), 0) : g.N;
  const int r0 = tile * BM, c0 = blockIdx.x * BN;
  if (r0 >= rows || c0 >= ncols) return;
  const int rbase = g.segs.base[seg];
  const int wbase = MODE == 1 ? g.segs.base[1] : 0;

  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Cfg::kStages * Cfg::kStageBytes);
  uint64_t* full = bars;                                // [kStages] TMA bytes landed (A raw, W_hi, W_lo)
  uint64_t* empty = bars + Cfg::kStages;                // [kStages] the MMAs that read the W tiles have retired
  uint64_t* conv = bars + 2 * Cfg::kStages;             // [kRing]   A_hi / A_lo of a k-block are in TMEM
  uint64_t* tfree = conv + Cfg::kRing;                  // [kRing]   the MMAs that read that TMEM slot have retired
  uint64_t* accum_full = tfree + Cfg::kRing;            // [1]
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_full + 1);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::kBiasOff);   // [BN], 16-byte aligned

  // Warp roles: 0..3 = data warps (TMEM lane quarter = warp), 4 = TMA producer, 5/6 = MMA issuers.  The issuers get
  // the highest warp ids: the arbiter favours high warp ids, and a starved issuer starves the tensor pipe.
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpTma = 4, kWarpMma = 5;          // issuer t (0 or 1) is lane 0 of warp kWarpMma + t
  const int nkb = (g.K0 + g.K1) / Cfg::kBlockK;

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (g.K1) tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapWhi);
    tma_prefetch_desc(&mapWlo);
    for (int s = 0; s < Cfg::kStages; ++s) {
      mbar_init(&full[s], 1);
      mbar_init(&empty[s], 1);
    }
    for (int s = 0; s < Cfg::kRing; ++s) {
      mbar_init(&conv[s], PREC ? 8 : 4);   // one arrival per splitter warp
      mbar_init(&tfree[s], 1);
    }
    mbar_init(accum_full, (Cfg::kIssuers == 2 && nkb >= 2) ? 2 : 1);   // one commit per issuer that has k-blocks
    fence_barrier_init();
  }
  auto stage_ptr = [&](int s) { return smem + (size_t)s * Cfg::kStageBytes; };
  int tma_kb = 0, tma_s = 0; uint32_t tma_ph = 0;
  auto tma_produce = [&](int kb_end) {
    for (; tma_kb < kb_end; ++tma_kb) {
      mbar_wait(&empty[tma_s], tma_ph ^ 1);
      uint8_t* st = stage_ptr(tma_s);
      mbar_arrive_expect_tx(&full[tma_s], Cfg::kABytes + 2 * Cfg::kWBytes);
      const int k = tma_kb * Cfg::kBlockK;
#pragma unroll
      for (int h = 0; h < Cfg::kBlockK / 32; ++h) {                // raw fp32 A: one 128-byte-swizzled box per 32 floats
        const int ka = k + 32 * h;
        if (ka < g.K0) tma_load_2d(st + h * 16384, &mapA0, &full[tma_s], ka, rbase + r0);
        else           tma_load_2d(st + h * 16384, &mapA1, &full[tma_s], ka - g.K0, rbase + r0);
      }
      tma_load_2d(st + Cfg::kABytes, &mapWhi, &full[tma_s], k, wbase + c0);
      tma_load_2d(st + Cfg::kABytes + Cfg::kWBytes, &mapWlo, &full[tma_s], k, wbase + c0);
      if (++tma_s == Cfg::kStages) { tma_s = 0; tma_ph ^= 1; }
    }
  };
  if (warp == kWarpMma) tmem_alloc<Cfg::kTmemCols>(tmem_slot);
  if (threadIdx.x < BN) bias_s[threadIdx.x] = (g.bias && c0 + (int)threadIdx.x < g.N) ? g.bias[c0 + threadIdx.x] : 0.f;
  tcgen05_fence_before();
  // The TMA warp only ARRIVES at the start-up barrier (it has initialised the mbarriers, and needs neither the TMEM
  // address nor the bias) and goes straight on to request the first stages; everybody else waits for it and for the
  // TMEM allocation.  (With a plain __syncthreads the first TMA was issued ~1200 clk into the kernel; issuing it from
  // inside the sync-ing warp before the barrier delayed the barrier for everybody instead.)
  if (warp == kWarpTma) asm volatile("bar.arrive 1, %0;" ::"n"(tc_threads(PREC)) : "memory");
  else                  asm volatile("bar.sync 1, %0;" ::"n"(tc_threads(PREC)) : "memory");
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);   // warp-uniform for the compiler
  if (trace && threadIdx.x == 0) trace[1] = clock64();

  if (warp == kWarpTma) {
    // ===== TMA producer =====
    if (lane == 0) tma_produce(nkb);
  } else if (warp == kWarpMma || warp == kWarpMma + 1) {
    // ===== MMA issuers: lane 0 of warp kWarpMma + t handles the k-blocks kb = t, t + kIssuers, ... =====
    // A comes from TMEM (written by the splitters), W from shared memory: with A in shared memory as well the
    // operand reads of three MMAs per k-step saturated the shared-memory port (profiles/r01_gemm_pipeline_trace.txt).
    const int t = __shfl_sync(0xffffffffu, warp - kWarpMma, 0);   // provably warp-uniform: the MMA operands stay in
    if (lane == 0 && t < Cfg::kIssuers) {                        // uniform registers (no ELECT / R2UR.BROADCAST per MMA)
      constexpr uint32_t idesc = PREC ? umma_idesc_f16(BM, BN, 0) : umma_idesc_tf32(BM, BN);
      // one descriptor per stage, built once: inside the loop an operand advance is a single 64-bit add
      const uint64_t d_stage0 = umma_desc_sw128(smem_u32(stage_ptr(0) + Cfg::kABytes));
      int last = -1;
      for (int kb = t; kb < nkb; kb += Cfg::kIssuers) last = kb;
      for (int kb = t; kb < nkb; kb += Cfg::kIssuers) {
        const int s = kb % Cfg::kStages;
        const int slot = kb % Cfg::kRing;
        mbar_wait(&conv[slot], (uint32_t)(kb / Cfg::kRing) & 1u);   // the splitters arrive after they saw full[s]
        tcgen05_fence_after();
        if (trace && kb < 8) trace[24 + kb] = clock64();
        const uint64_t w_hi = d_stage0 + (uint64_t)((s * Cfg::kStageBytes) >> 4);
        const uint64_t w_lo = w_hi + (Cfg::kWBytes >> 4);
        const uint32_t a_hi = tmem_base + Cfg::kRingCol + slot * 64;
        const uint32_t a_lo = a_hi + 32;
        // issuer t owns main accumulator t (kb % 2 == t) and, unless folded into it, correction accumulator t
        const uint32_t main_acc = tmem_base + t * BN;
        const uint32_t corr = Cfg::kFold ? main_acc : tmem_base + (Cfg::kMainAcc + t) * BN;
        const uint32_t fresh = kb >= Cfg::kIssuers ? 1u : 0u;      // 0: first k-block of this issuer
        // 4 k-steps per k-block either way: 8 tf32 or 16 fp16 values = 32 bytes inside the 128-B swizzle row of W and
        // 8 TMEM columns of A
#pragma unroll
        for (int ks = 0; ks < 4; ++ks) {
          const uint64_t koff = (ks * 32) >> 4;
          if (PREC) {
            umma_f16_ts(corr, a_lo + ks * 8, w_hi + koff, idesc, ks ? 1u : fresh);
            umma_f16_ts(corr, a_hi + ks * 8, w_lo + koff, idesc, 1u);
            umma_f16_ts(main_acc, a_hi + ks * 8, w_hi + koff, idesc, (Cfg::kFold || ks) ? 1u : fresh);
          } else {
            umma_tf32_ts(corr, a_lo + ks * 8, w_hi + koff, idesc, ks ? 1u : fresh);
            umma_tf32_ts(corr, a_hi + ks * 8, w_lo + koff, idesc, 1u);
            umma_tf32_ts(main_acc, a_hi + ks * 8, w_hi + koff, idesc, (Cfg::kFold || ks) ? 1u : fresh);
          }
        }
        if (trace && kb < 8) trace[32 + kb] = clock64();
        umma_commit(&empty[s]);                        // W slot reusable once these MMAs retire
        umma_commit(&tfree[slot]);                     // and so is the A slot in TMEM
        if (kb == last) umma_commit(accum_full);       // this issuer's accumulators are complete
        if (trace && kb < 8) trace[40 + kb] = clock64();
      }
    }
  } else if (warp < 4 || warp >= 7) {
    // ===== A splitters (warps 0..3 and, at PREC 1, 7..10; thread = tile row = TMEM lane), then epilogue =====
    const int grp = warp >= 7 ? 1 : 0;                 // PREC 1: group g converts box g of every k-block, and takes every
    constexpr int kGroups = PREC ? 2 : 1;              // second epilogue chunk
    const int q = warp & 3;                            // TMEM lane quarter this warp may access (warp id mod 4)
    const int t = 32 * q + lane;                       // tile row 0..127
    const bool tr0 = threadIdx.x == 0;                 // the thread that writes the trace
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    float amax = 0.f;                                  // PREC 1: largest |A| this thread has converted (fp16 range guard)
    {
      int s = 0; uint32_t ph = 0;
      for (int kb = 0; kb < nkb; ++kb) {
        const int slot = kb % Cfg::kRing;
        mbar_wait(&full[s], ph);
        if (trace && tr0 && kb < 8) trace[8 + kb] = clock64();
        // row t of the 128-byte-swizzled tile: 16-byte chunk j sits at position j ^ (t & 7)
        uint32_t x[32], hi[32];
        if (PREC) {
          // my 32 floats of this row (box grp) -> 16 packed fp16 hi / lo words: value 2 w in the low half of word w
          const uint4* rowp = reinterpret_cast<const uint4*>(stage_ptr(s) + grp * 16384 + t * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            const uint4 v = rowp[j ^ (t & 7)];
            const float f0 = __uint_as_float(v.x), f1 = __uint_as_float(v.y), f2 = __uint_as_float(v.z), f3 = __uint_as_float(v.w);
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(f0), fabsf(f1)), fmaxf(fabsf(f2), fabsf(f3))));
            split16x2<0>(f0, f1, hi[2 * j], x[2 * j]);
            split16x2<0>(f2, f3, hi[2 * j + 1], x[2 * j + 1]);
          }
        } else {
          const uint4* rowp = reinterpret_cast<const uint4*>(stage_ptr(s) + t * 128);
#pragma unroll
          for (int j = 0; j < 8; ++j) {
            uint4 v = rowp[j ^ (t & 7)];
            x[4 * j] = v.x; x[4 * j + 1] = v.y; x[4 * j + 2] = v.z; x[4 * j + 3] = v.w;
          }
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            float h, l;
            split_tf32(__uint_as_float(x[j]), h, l);
            hi[j] = __float_as_uint(h); x[j] = __float_as_uint(l);
          }
        }
        mbar_wait(&tfree[slot], ((uint32_t)(kb / Cfg::kRing) & 1u) ^ 1u);   // MMAs of k-block kb - kRing are done
        tcgen05_fence_after();
        if (trace && tr0 && kb < 8) trace[48 + kb] = clock64();
        if (PREC) {
          uint32_t h16[16], l16[16];
#pragma unroll
          for (int j = 0; j < 16; ++j) { h16[j] = hi[j]; l16[j] = x[j]; }
          tmem_st_32x16(lane_base + Cfg::kRingCol + slot * 64 + 16 * grp, h16);
          tmem_st_32x16(lane_base + Cfg::kRingCol + slot * 64 + 32 + 16 * grp, l16);
        } else {
          tmem_st_32x32(lane_base + Cfg::kRingCol + slot * 64, hi);
          tmem_st_32x32(lane_base + Cfg::kRingCol + slot * 64 + 32, x);
        }
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[slot]);
        if (trace && tr0 && kb < 8) trace[16 + kb] = clock64();
        if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
      }
    }
    // rows past the live count may hold anything (other segments, stale scratch): only live rows raise the range flag
    if (PREC && g.status && amax >= 32768.f && r0 + t < rows) atomicOr(g.status, GIMS_STATUS_FP16_RANGE);
    // ---- epilogue ----
    {
      const int n_main = nkb < Cfg::kMainAcc ? nkb : Cfg::kMainAcc;
      const int n_corr = nkb < Cfg::kCorrAcc ? nkb : Cfg::kCorrAcc;
      const TileCoord tcd = {seg, rows, ncols, r0, c0, rbase};
      // each warp transposes through a private staging tile in the (now idle) pipeline stages
      const uint32_t stg = smem_u32(smem) + (uint32_t)((q + 4 * grp) * 32 * 36 * 4);
      epilogue_tile<BN, MODE, PREC>(g, tcd, lane_base, n_main, n_corr, grp, kGroups, q, lane, stg, bias_s,
                                    [&] { mbar_wait(accum_full, 0); tcgen05_fence_after(); }, trace, tr0);
    }
    tcgen05_fence_before();
    if (trace && tr0) trace[4] = clock64();
  }
  __syncthreads();
  if (warp == kWarpMma) {
    tcgen05_fence_after();
    tmem_dealloc<Cfg::kTmemCols>(tmem_base);
  }
  if (trace && threadIdx.x == 0) {
    trace[5] = clock64();
    unsigned long long gt; asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
    trace[7] = (long long)gt;
  }
}

// =====================================================================================================================
// Persistent variant of the fp16 hi+lo kernel (BN = 128): a CTA walks over several output tiles, and the store-bound
// epilogue of tile i runs on its own warps while the k-loop of tile i+1 is already feeding the tensor pipe.  One tile
// per CTA spends more than half of its life outside the k-loop (prologue 1.4 k clk, first TMA round trip 1.4 k, epilogue
// 3.3 k of 11.8 k clk: profiles/r01_gemm_pipeline_trace.txt, tools/gemm_trace.py); here that is paid once per CTA.
//
// TMEM (512 columns): accumulators 0 / 1 of the tile in flight (2 x 128, even / odd k-blocks as in k_gemm_tc — the chains
// stay short, the arithmetic is identical) | PARK 128 | A ring 2 x 64.  When a tile's MMAs have retired the splitter
// warps fold accumulator 0 + 1 into PARK (tcgen05.ld, add, tcgen05.st: ~400 clk) and hand the accumulators back to the
// issuers; the four epilogue warps drain PARK with the code of the one-tile kernel (epilogue_tile).
//   warps 0-3, 7-10  A splitters (as k_gemm_tc at PREC 1) + the fold      warp 4  TMA producer
//   warps 5, 6       MMA issuers                                          warps 11-14  epilogue (lane quarter = warp & 3)
// Barriers per tile: accum_full (issuers -> splitters), acc_free + park_full (fold -> issuers / epilogue),
// park_free (epilogue -> next fold).  The TMA / conversion rings count k-blocks across tiles.
// =====================================================================================================================
constexpr int kThreadsP = 480;
struct TcpCfg {
  static constexpr int BN = 128;
  static constexpr int kBlockK = 64;
  static constexpr int kABytes = BM * kBlockK * 4;          // 32 KB raw fp32 A (two 32-float boxes)
  static constexpr int kWBytes = BN * 64 * 2;               // 16 KB per 16-bit W plane
  static constexpr int kStageBytes = kABytes + 2 * kWBytes; // 64 KB
  static constexpr int kStages = 3;
  static constexpr int kRing = 2;
  static constexpr int kParkCol = 2 * BN;
  static constexpr int kRingCol = 3 * BN;
  static constexpr int kStgBytes = 4 * 32 * 36 * 4;         // one staging tile per epilogue warp
  static constexpr int kBarriers = 2 * kStages + 2 * kRing + 4;
  static constexpr int kBiasOff = (kBarriers * 8 + 16 + 15) & ~15;
  static constexpr int kSmemBytes = kStages * kStageBytes + kStgBytes + 1024 /*align*/ + kBiasOff + 2 * BN * 4;
};

// PLANES 2: fp16 hi + lo operands, three MMAs per k-step (fp32 parity).  PLANES 1: ONE bf16 plane per operand, one MMA per
// k-step — the bf16 variant of the score GEMM (BASELINE configs[4]), reported separately, never the fp32 line.
template <int MODE, int PLANES>
__global__ void __launch_bounds__(kThreadsP, 1)
k_gemm_tcp(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
           const __grid_constant__ CUtensorMap mapWhi, const __grid_constant__ CUtensorMap mapWlo, TcArgs g, int col_tiles,
           int total_work) {
  using Cfg = TcpCfg;
  constexpr int BN = Cfg::BN;
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* stg_base = smem + Cfg::kStages * Cfg::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(stg_base + Cfg::kStgBytes);
  uint64_t* full = bars;                                // [kStages] TMA bytes landed (A raw, W_hi, W_lo)
  uint64_t* empty = bars + Cfg::kStages;                // [kStages] the MMAs that read the W tiles have retired
  uint64_t* conv = bars + 2 * Cfg::kStages;             // [kRing]   A_hi / A_lo of a k-block are in TMEM
  uint64_t* tfree = conv + Cfg::kRing;                  // [kRing]   the MMAs that read that TMEM slot have retired
  uint64_t* accum_full = tfree + Cfg::kRing;            // the tile's MMAs have retired
  uint64_t* acc_free = accum_full + 1;                  // accumulators folded into PARK: the next tile may overwrite them
  uint64_t* park_full = acc_free + 1;                   // PARK holds a finished tile
  uint64_t* park_free = park_full + 1;                  // the epilogue has read PARK
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(park_free + 1);
  float* bias_s = reinterpret_cast<float*>(reinterpret_cast<uint8_t*>(bars) + Cfg::kBiasOff);   // [2][BN]

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr int kWarpTma = 4, kWarpMma = 5, kWarpEpi = 11;
  // optional timeline of CTA 0 (tools/gemm_trace.py): 8 stamps per tile for the first 4 tiles, after the entry stamp
  long long* trace = (blockIdx.x == 0 && lane == 0) ? g_gemm_trace : nullptr;
  if (trace && threadIdx.x == 0) trace[0] = clock64();
#define TCP_TRACE(k) do { if (trace && lt < 4) trace[8 + lt * 8 + (k)] = clock64(); } while (0)
  const int nkb = (g.K0 + g.K1) / Cfg::kBlockK;
  const bool single = g.single_chain != 0;              // experiment: one issuer, one accumulation chain per tile
  const int n_issuers = (nkb >= 2 && !single) ? 2 : 1;

  // the live row counts are device scalars: read once (a global load per tile per role cost ~800 clk at every tile
  // boundary), before the start-up barrier below
  __shared__ int cnt_s[kMaxSegs];
  if (threadIdx.x < kMaxSegs) cnt_s[threadIdx.x] = (int)threadIdx.x < g.segs.nseg ? seg_count(g.segs, threadIdx.x) : 0;
  // work item -> tile; identical in every role.  Tiles outside the live rows / columns are skipped by everybody.
  auto decode = [&](int w, TileCoord& t, bool from_global = false) -> bool {
    const int rt = w / col_tiles, ct = w - rt * col_tiles;
    int seg = 0, tile = rt;
    if (MODE != 1) {
      while (seg + 1 < g.segs.nseg && rt >= g.tile_end[seg]) ++seg;
      tile = rt - (seg ? g.tile_end[seg - 1] : 0);
    }
    t.seg = seg;
    // (the TMA warp does not wait at the start-up barrier and reads the counts itself; it runs ahead anyway)
    t.rows = from_global ? seg_count(g.segs, seg) : cnt_s[seg];
    t.ncols = MODE == 1 ? (from_global ? seg_count(g.segs, 1) : cnt_s[1]) : g.N;
    t.r0 = tile * BM; t.c0 = ct * BN;
    t.rbase = g.segs.base[seg];
    return t.r0 < t.rows && t.c0 < t.ncols;
  };
  const int wbase = MODE == 1 ? g.segs.base[1] : 0;

  if (warp == kWarpTma && lane == 0) {
    tma_prefetch_desc(&mapA0);
    if (g.K1) tma_prefetch_desc(&mapA1);
    tma_prefetch_desc(&mapWhi);
    tma_prefetch_desc(&mapWlo);
    for (int s = 0; s < Cfg::kStages; ++s) { mbar_init(&full[s], 1); mbar_init(&empty[s], 1); }
    for (int s = 0; s < Cfg::kRing; ++s) { mbar_init(&conv[s], 8); mbar_init(&tfree[s], 1); }
    mbar_init(accum_full, n_issuers);
    mbar_init(acc_free, 8);
    mbar_init(park_full, 8);
    mbar_init(park_free, 4);
    fence_barrier_init();
  }
  if (warp == kWarpMma) tmem_alloc<512>(tmem_slot);
  tcgen05_fence_before();
  if (warp == kWarpTma) asm volatile("bar.arrive 1, %0;" ::"n"(kThreadsP) : "memory");
  else                  asm volatile("bar.sync 1, %0;" ::"n"(kThreadsP) : "memory");
  tcgen05_fence_after();
  const uint32_t tmem_base = __shfl_sync(0xffffffffu, *tmem_slot, 0);
  auto stage_ptr = [&](int s) { return smem + (size_t)s * Cfg::kStageBytes; };

  if (warp == kWarpTma) {
    // ===== TMA producer: runs ahead across tile boundaries as far as the stage ring allows =====
    if (lane == 0) {
      int s = 0; uint32_t ph = 0;
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        TileCoord t;
        if (!decode(w, t, true)) continue;
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(&empty[s], ph ^ 1);
          uint8_t* st = stage_ptr(s);
          mbar_arrive_expect_tx(&full[s], Cfg::kABytes + PLANES * Cfg::kWBytes);
          const int k = kb * Cfg::kBlockK;
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            const int ka = k + 32 * h;
            if (ka < g.K0) tma_load_2d(st + h * 16384, &mapA0, &full[s], ka, t.rbase + t.r0);
            else           tma_load_2d(st + h * 16384, &mapA1, &full[s], ka - g.K0, t.rbase + t.r0);
          }
          tma_load_2d(st + Cfg::kABytes, &mapWhi, &full[s], k, wbase + t.c0);
          if (PLANES == 2) tma_load_2d(st + Cfg::kABytes + Cfg::kWBytes, &mapWlo, &full[s], k, wbase + t.c0);
          if (++s == Cfg::kStages) { s = 0; ph ^= 1; }
        }
      }
    }
  } else if (warp == kWarpMma || warp == kWarpMma + 1) {
    // ===== MMA issuers: lane 0 of warp kWarpMma + t handles the k-blocks kb = t, t + 2, ... of every tile =====
    const int t = __shfl_sync(0xffffffffu, warp - kWarpMma, 0);
    if (lane == 0) {
      constexpr uint32_t idesc = umma_idesc_f16(BM, BN, PLANES == 2 ? 0 : 1);
      const uint64_t d_stage0 = umma_desc_sw128(smem_u32(stage_ptr(0) + Cfg::kABytes));
      const uint32_t acc = tmem_base + t * BN;
      const int kstep = single ? 1 : 2;
      int last = -1;
      if (!(single && t > 0))
        for (int kb = t; kb < nkb; kb += kstep) last = kb;
      int lt = 0;                                          // live tiles so far
      for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
        TileCoord tcd;
        if (!decode(w, tcd)) continue;
        if (lt > 0 && last >= 0) { mbar_wait(acc_free, (uint32_t)(lt - 1) & 1u); tcgen05_fence_after(); }
        for (int kb = t; kb < nkb && last >= 0; kb += kstep) {
          const int gk = lt * nkb + kb;
          const int s = gk % Cfg::kStages, slot = gk % Cfg::kRing;
          mbar_wait(&conv[slot], (uint32_t)(gk / Cfg::kRing) & 1u);
          tcgen05_fence_after();
          const uint64_t w_hi = d_stage0 + (uint64_t)((s * Cfg::kStageBytes) >> 4);
          const uint64_t w_lo = w_hi + (Cfg::kWBytes >> 4);
          const uint32_t a_hi = tmem_base + Cfg::kRingCol + slot * 64;
          const uint32_t a_lo = a_hi + 32;
          const uint32_t fresh = kb >= kstep ? 1u : 0u;    // 0: first k-block of this issuer in this tile
          if (t == 0 && kb == 0) TCP_TRACE(7);
#pragma unroll
          for (int ks = 0; ks < 4; ++ks) {
            const uint64_t koff = (ks * 32) >> 4;
            if (PLANES == 2) {
              umma_f16_ts(acc, a_lo + ks * 8, w_hi + koff, idesc, ks ? 1u : fresh);
              umma_f16_ts(acc, a_hi + ks * 8, w_lo + koff, idesc, 1u);
              umma_f16_ts(acc, a_hi + ks * 8, w_hi + koff, idesc, 1u);
            } else {
              umma_f16_ts(acc, a_hi + ks * 8, w_hi + koff, idesc, ks ? 1u : fresh);
            }
          }
          umma_commit(&empty[s]);
          umma_commit(&tfree[slot]);
          if (kb == last) umma_commit(accum_full);
        }
        ++lt;
      }
    }
  } else if (warp < 4 || (warp >= 7 && warp < kWarpEpi)) {
    // ===== A splitters (thread = tile row = TMEM lane; group grp converts box grp of every k-block), then the fold =====
    const int grp = warp >= 7 ? 1 : 0;
    const int q = warp & 3;
    const int t = 32 * q + lane;
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16);
    const int n_main = (nkb < 2 || single) ? 1 : 2;
    // PARK = accumulator 0 + accumulator 1 of tile f (fp32, RN — what the one-tile epilogue computes), then the
    // accumulators go back to the issuers
    auto fold = [&](int f) {
      int lt = f;                                          // (for the trace macro)
      mbar_wait(accum_full, (uint32_t)f & 1u);
      tcgen05_fence_after();
      if (warp == 0) TCP_TRACE(2);
      if (f > 0) { mbar_wait(park_free, (uint32_t)(f - 1) & 1u); tcgen05_fence_after(); }
      if (warp == 0) TCP_TRACE(3);
#pragma unroll
      for (int cc = grp; cc < BN / 32; cc += 2) {
        uint32_t v[32];
        tmem_ld_32x32(lane_base + (uint32_t)(cc * 32), v);
        if (n_main > 1) {
          uint32_t x[32];
          tmem_ld_32x32(lane_base + (uint32_t)(BN + cc * 32), x);
          tmem_ld_wait();
#pragma unroll
          for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(x[j]));
        } else {
          tmem_ld_wait();
        }
        tmem_st_32x32(lane_base + (uint32_t)(Cfg::kParkCol + cc * 32), v);
      }
      tmem_st_wait();
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) { mbar_arrive(acc_free); mbar_arrive(park_full); }
      if (warp == 0) TCP_TRACE(4);
    };
    // (Converting the first k-blocks of the next tile BEFORE the fold was tried: the fold then starts later by as much
    // as the issuers gain afterwards — the stage ring, not the conversion, decides when the next tile's operands land.)
    int lt = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      TileCoord tcd;
      if (!decode(w, tcd)) continue;
      float amax = 0.f;
      for (int kb = 0; kb < nkb; ++kb) {
        const int gk = lt * nkb + kb;
        const int s = gk % Cfg::kStages, slot = gk % Cfg::kRing;
        mbar_wait(&full[s], (uint32_t)(gk / Cfg::kStages) & 1u);
        if (warp == 0 && kb == 0) TCP_TRACE(0);
        uint32_t hi[16], lo[16];
        const uint4* rowp = reinterpret_cast<const uint4*>(stage_ptr(s) + grp * 16384 + t * 128);
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          const uint4 v = rowp[j ^ (t & 7)];
          const float f0 = __uint_as_float(v.x), f1 = __uint_as_float(v.y), f2 = __uint_as_float(v.z), f3 = __uint_as_float(v.w);
          if (PLANES == 2) {
            amax = fmaxf(amax, fmaxf(fmaxf(fabsf(f0), fabsf(f1)), fmaxf(fabsf(f2), fabsf(f3))));
            split16x2<0>(f0, f1, hi[2 * j], lo[2 * j]);
            split16x2<0>(f2, f3, hi[2 * j + 1], lo[2 * j + 1]);
          } else {
            hi[2 * j] = pack16<1>(f0, f1);
            hi[2 * j + 1] = pack16<1>(f2, f3);
          }
        }
        mbar_wait(&tfree[slot], ((uint32_t)(gk / Cfg::kRing) & 1u) ^ 1u);   // MMAs of k-block gk - kRing are done
        tcgen05_fence_after();
        tmem_st_32x16(lane_base + Cfg::kRingCol + slot * 64 + 16 * grp, hi);
        if (PLANES == 2) tmem_st_32x16(lane_base + Cfg::kRingCol + slot * 64 + 32 + 16 * grp, lo);
        tmem_st_wait();
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(&conv[slot]);
      }
      if (warp == 0) TCP_TRACE(1);
      // rows past the live count may hold anything (other segments, stale scratch): only live rows raise the range flag
      if (PLANES == 2 && g.status && amax >= 32768.f && tcd.r0 + t < tcd.rows) atomicOr(g.status, GIMS_STATUS_FP16_RANGE);
      fold(lt);
      ++lt;
    }
  } else {
    // ===== epilogue warps: drain PARK while the next tile's k-loop runs =====
    const int q = warp & 3;
    const int et = threadIdx.x - kWarpEpi * 32;            // 0..127
    const uint32_t lane_base = tmem_base + ((uint32_t)(32 * q) << 16) + Cfg::kParkCol;
    const uint32_t stg = smem_u32(stg_base) + (uint32_t)((warp - kWarpEpi) * 32 * 36 * 4);
    int lt = 0;
    for (int w = blockIdx.x; w < total_work; w += gridDim.x) {
      TileCoord tcd;
      if (!decode(w, tcd)) continue;
      float* bias_t = bias_s + (lt & 1) * BN;             // the other half may still be read by a slower epilogue warp
      bias_t[et] = (g.bias && tcd.c0 + et < g.N) ? g.bias[tcd.c0 + et] : 0.f;
      asm volatile("bar.sync 2, 128;" ::: "memory");
      epilogue_tile<BN, MODE, 1>(g, tcd, lane_base, 1, 0, 0, 1, q, lane, stg, bias_t,
                                 [&] { mbar_wait(park_full, (uint32_t)lt & 1u); tcgen05_fence_after();
                                       if (warp == kWarpEpi) TCP_TRACE(5); }, nullptr, false);
      tcgen05_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(park_free);
      if (warp == kWarpEpi) TCP_TRACE(6);
      ++lt;
    }
  }
#undef TCP_TRACE
  __syncthreads();
  if (warp == kWarpMma) {
    tcgen05_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  if (trace && threadIdx.x == 0) trace[1] = clock64();
}

// x -> (hi, lo) planes, hi = tf32(x); used for activations that act as the "weight" operand (score GEMM)
__global__ void k_split_planes(const float* __restrict__ x, float* __restrict__ hi, float* __restrict__ lo, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  float4 v = reinterpret_cast<const float4*>(x)[i], h, l;
  tc::split_tf32(v.x, h.x, l.x); tc::split_tf32(v.y, h.y, l.y); tc::split_tf32(v.z, h.z, l.z); tc::split_tf32(v.w, h.w, l.w);
  reinterpret_cast<float4*>(hi)[i] = h;
  reinterpret_cast<float4*>(lo)[i] = l;
}

// x -> fp16 (hi, lo) planes, hi = fp16(x), lo = fp16(x - hi); the score GEMM's "weight" operand at PREC 1
__global__ void k_split_planes16(const float* __restrict__ x, uint16_t* __restrict__ hi, uint16_t* __restrict__ lo, size_t n4,
                                 unsigned* status) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  uint2 h, l;
  tc::split16x2<0>(v.x, v.y, h.x, l.x);
  tc::split16x2<0>(v.z, v.w, h.y, l.y);
  reinterpret_cast<uint2*>(hi)[i] = h;
  reinterpret_cast<uint2*>(lo)[i] = l;
  // (rows past the live counts hold whatever the projection left there: only finite overflow is flagged)
  const float m = fmaxf(fmaxf(fabsf(v.x), fabsf(v.y)), fmaxf(fabsf(v.z), fabsf(v.w)));
  if (status && m >= 32768.f && m < 3e38f) atomicOr(status, GIMS_STATUS_FP16_RANGE);
}

// x -> one bf16 plane (the bf16 variant of the score GEMM)
__global__ void k_plane_bf16(const float* __restrict__ x, uint16_t* __restrict__ hi, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n4) return;
  const float4 v = reinterpret_cast<const float4*>(x)[i];
  uint2 h;
  h.x = tc::pack16<1>(v.x, v.y);
  h.y = tc::pack16<1>(v.z, v.w);
  reinterpret_cast<uint2*>(hi)[i] = h;
}

template <int BN, int MODE, int PREC = 0>
int launch_tc(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& whi, const CUtensorMap& wlo,
              const TcArgs& g, int col_tiles, int row_tiles, int prof_class, cudaStream_t st) {
  using Cfg = TcCfg<BN, PREC>;
  GIMS_CUDA_OK(cudaFuncSetAttribute(k_gemm_tc<BN, MODE, PREC>, cudaFuncAttributeMaxDynamicSharedMemorySize, Cfg::kSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(col_tiles, row_tiles);
  cfg.blockDim = dim3(tc_threads(PREC));
  cfg.dynamicSmemBytes = Cfg::kSmemBytes;
  cfg.stream = st;
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  ProfScope prof(prof_class, st);
  GIMS_CUDA_OK(cudaLaunchKernelEx(&cfg, k_gemm_tc<BN, MODE, PREC>, a0, a1, whi, wlo, g));
  count_launch();
  return GIMS_OK;
}

// Persistent launch: `total` = col_tiles * row_tiles work items, row-tile major (the column tiles of one row tile run on
// neighbouring CTAs at the same time and share the A tile in L2).  Grid: enough CTAs that each gets about
// GIMS_GEMM_TPC (default 2) tiles — what matters when several streams keep the GPU full is SM-time per tile, and a CTA's
// fixed cost (prologue, first TMA round trip, last epilogue) is then paid once per two tiles; small problems (fewer
// work items than half the SMs) keep one tile per CTA for latency.
int single_chain_knob() {
  static const int on = [] { const char* e = getenv("GIMS_GEMM_SINGLE_CHAIN"); return (e && e[0] == '1') ? 1 : 0; }();
  return on;
}
bool persist_enabled() {
  static const bool on = [] { const char* e = getenv("GIMS_GEMM_PERSIST"); return !(e && e[0] == '0'); }();
  return on;
}
template <int MODE, int PLANES = 2>
int launch_tcp(const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& whi, const CUtensorMap& wlo,
               const TcArgs& g, int col_tiles, int row_tiles, int prof_class, cudaStream_t st) {
  static const int tpc_env = [] { const char* e = getenv("GIMS_GEMM_TPC"); int v = e ? atoi(e) : 2; return v < 1 ? 1 : v; }();
  int dev = 0, sms = 0;
  GIMS_CUDA_OK(cudaGetDevice(&dev));
  GIMS_CUDA_OK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
  const int total = col_tiles * row_tiles;
  const int tpc = (2 * total <= sms) ? 1 : tpc_env;
  int grid = (total + tpc - 1) / tpc;
  if (grid > sms) grid = sms;
  GIMS_CUDA_OK(cudaFuncSetAttribute(k_gemm_tcp<MODE, PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize, TcpCfg::kSmemBytes));
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(kThreadsP);
  cfg.dynamicSmemBytes = TcpCfg::kSmemBytes;
  cfg.stream = st;
  cfg.attrs = nullptr;
  cfg.numAttrs = 0;
  ProfScope prof(prof_class, st);
  GIMS_CUDA_OK(cudaLaunchKernelEx(&cfg, k_gemm_tcp<MODE, PLANES>, a0, a1, whi, wlo, g, col_tiles, total));
  count_launch();
  return GIMS_OK;
}

int pick_bn(int n) {
  static const int forced = [] { const char* e = getenv("GIMS_GEMM_BN"); return e ? atoi(e) : 0; }();   // tuning knob
  if (forced == 64 || forced == 128 || forced == 192) return forced;
  // Fewer, fatter CTAs: per output column a 128-wide tile costs half the tensor-pipe time of a 64-wide one, and with
  // several pairs in flight the SMs a launch leaves idle are used by the other streams (N = 256 at BN = 128 is 64 CTAs).
  if (n >= 768) return 192;
  if (n >= 256) return 128;
  return 64;
}

}  // namespace

// W planes: [N][K] produced at pack time (gims_b200/packing.py) — same layout as the fp32 weight.  prec 0: tf32 hi / lo
// (fp32 words); prec 1: fp16 hi / lo of W * 2^e with *sinv = 2^-e.
int launch_gemm_tc(const GemmArgs& a, const WPlanes& w, int prec, cudaStream_t st, const QkvPlanes* qkv) {
  int K = a.K0 + a.K1;
  if (prec && (a.K0 % 64 || a.K1 % 64 || !w.h16 || !w.l16 || !w.sinv)) prec = 0;     // shapes the fp16 kernel does not take
  if (a.K0 % BK || a.K1 % BK || K == 0 || a.N % 32 || (a.lda0 % 4) || (a.K1 && (a.lda1 % 4)) || !w.hi32 || !w.lo32) {
    set_error("launch_gemm_tc: unsupported shape K0=%d K1=%d N=%d", a.K0, a.K1, a.N);
    return GIMS_ERR_ARG;
  }
  int total_rows = segs_rows(a.segs);
  int bn = pick_bn(a.N);
  if (qkv && (bn != 64 || qkv->fmt >= 0)) bn = 192;
  if (prec) bn = (qkv || (a.N >= 256 && bn != 64)) ? 128 : 64;      // (GIMS_GEMM_BN=64 forces the two-CTAs-per-SM shape)
  CUtensorMap mA0, mA1, mWh, mWl;
  GIMS_TRY(tc::make_tmap_f32_k32(&mA0, a.A0, total_rows, a.K0, a.lda0, BM));
  if (a.K1) GIMS_TRY(tc::make_tmap_f32_k32(&mA1, a.A1, total_rows, a.K1, a.lda1, BM));
  else mA1 = mA0;
  if (prec) {
    GIMS_TRY(tc::make_tmap_16_k64(&mWh, w.h16, 0, a.N, K, K, bn));
    GIMS_TRY(tc::make_tmap_16_k64(&mWl, w.l16, 0, a.N, K, K, bn));
  } else {
    GIMS_TRY(tc::make_tmap_f32_k32(&mWh, w.hi32, a.N, K, K, bn));
    GIMS_TRY(tc::make_tmap_f32_k32(&mWl, w.lo32, a.N, K, K, bn));
  }
  TcArgs g;
  g.K0 = a.K0; g.K1 = a.K1; g.bias = a.bias; g.R = a.R; g.ldr = a.ldr; g.Y = a.Y; g.ldy = a.ldy; g.N = a.N;
  g.relu = a.relu; g.scale = 1.f; g.score = 0; g.segs = a.segs;
  g.qkv = 0; g.qp = nullptr; g.kp = g.vt = nullptr; g.ldv = 0; g.rows_total = total_rows; g.status = a.status;
  g.wscale_inv = prec ? w.sinv : nullptr;
  g.single_chain = single_chain_knob();
  for (int i = 0; i < kMaxSegs; ++i) g.vbase[i] = 0;
  if (qkv) {
    if (a.N != 3 * kD || !a.bias) { set_error("launch_gemm_tc: qkv mode needs N = 768 and a bias"); return GIMS_ERR_ARG; }
    g.qkv = 1; g.qp = qkv->qp; g.kp = qkv->kp; g.vt = qkv->vt; g.ldv = qkv->ldv;
    for (int i = 0; i < kMaxSegs; ++i) g.vbase[i] = qkv->vbase[i];
    if (qkv->status) g.status = qkv->status;
  }
  int tiles = 0;
  for (int i = 0; i < a.segs.nseg; ++i) { tiles += cdiv(a.segs.nmax[i], BM); g.tile_end[i] = tiles; }
  if (tiles == 0) return GIMS_OK;
  int ct = cdiv(a.N, bn);
  if (qkv) {
    if (qkv->fmt >= 0 && !((qkv->fmt == 0 && qkv->planes == 2) || (qkv->fmt == 1 && qkv->planes == 1))) {
      set_error("launch_gemm_tc: unsupported 16-bit plane format");
      return GIMS_ERR_ARG;
    }
    if (prec && persist_enabled()) {
      if (qkv->fmt == 0) return launch_tcp<3>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
      if (qkv->fmt == 1) return launch_tcp<4>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
      return launch_tcp<2>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
    }
    if (prec) {
      if (qkv->fmt == 0) return launch_tc<128, 3, 1>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
      if (qkv->fmt == 1) return launch_tc<128, 4, 1>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
      return launch_tc<128, 2, 1>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
    }
    if (qkv->fmt == 0) return launch_tc<192, 3>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
    if (qkv->fmt == 1) return launch_tc<192, 4>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
    if (bn == 64) return launch_tc<64, 2>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
    return launch_tc<192, 2>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
  }
  if (prec) {
    if (bn == 128 && persist_enabled()) return launch_tcp<0>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
    if (bn == 128) return launch_tc<128, 0, 1>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
    return launch_tc<64, 0, 1>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
  }
  switch (bn) {
    case 192: return launch_tc<192, 0>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
    case 128: return launch_tc<128, 0>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
    default:  return launch_tc<64, 0>(mA0, mA1, mWh, mWl, g, ct, tiles, GIMS_PROF_GEMM, st);
  }
}

int set_gemm_trace(long long* dev_buf) {
  GIMS_CUDA_OK(cudaMemcpyToSymbol(g_gemm_trace, &dev_buf, sizeof(dev_buf)));
  return GIMS_OK;
}

int launch_split_planes(const float* x, float* hi, float* lo, size_t n, cudaStream_t st) {
  size_t n4 = n / 4;
  k_split_planes<<<(unsigned)((n4 + 255) / 256), 256, 0, st>>>(x, hi, lo, n4);
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}

// couplings[i][j] = <mdesc0_i, mdesc1_j> / 16 on tensor cores; `planes` = 2 * (n0_max+n1_max) * 256 floats scratch.
// prec 1: fp16 hi / lo planes of mdesc (|mdesc| >= 32768 raises GIMS_STATUS_FP16_RANGE in *status)
int launch_score_gemm_tc(const float* mdesc, int n0_max, int n1_max, const int* n_dev, float* planes, float* couplings,
                         int prec, unsigned* status, cudaStream_t st) {
  size_t rows = (size_t)n0_max + n1_max;
  constexpr int bn = 128;
  CUtensorMap mA, mWh, mWl;
  GIMS_TRY(tc::make_tmap_f32_k32(&mA, mdesc, rows, kD, kD, BM));
  if (prec == 2) {
    // bf16 variant: one plane per operand, one MMA per k-step (no range concern: bf16 has the fp32 exponent)
    uint16_t* hi = reinterpret_cast<uint16_t*>(planes);
    k_plane_bf16<<<(unsigned)((rows * kD / 4 + 255) / 256), 256, 0, st>>>(mdesc, hi, rows * kD / 4);
    GIMS_LAUNCH_OK();
    GIMS_TRY(tc::make_tmap_16_k64(&mWh, hi, 1, rows, kD, kD, bn));
    mWl = mWh;
  } else if (prec) {
    uint16_t* hi = reinterpret_cast<uint16_t*>(planes);
    uint16_t* lo = hi + rows * kD;
    k_split_planes16<<<(unsigned)((rows * kD / 4 + 255) / 256), 256, 0, st>>>(mdesc, hi, lo, rows * kD / 4, status);
    GIMS_LAUNCH_OK();
    GIMS_TRY(tc::make_tmap_16_k64(&mWh, hi, 0, rows, kD, kD, bn));
    GIMS_TRY(tc::make_tmap_16_k64(&mWl, lo, 0, rows, kD, kD, bn));
  } else {
    float* hi = planes;
    float* lo = planes + rows * kD;
    GIMS_TRY(launch_split_planes(mdesc, hi, lo, rows * kD, st));
    GIMS_TRY(tc::make_tmap_f32_k32(&mWh, hi, rows, kD, kD, bn));
    GIMS_TRY(tc::make_tmap_f32_k32(&mWl, lo, rows, kD, kD, bn));
  }
  TcArgs g;
  g.K0 = kD; g.K1 = 0; g.bias = nullptr; g.R = nullptr; g.ldr = 0; g.Y = couplings; g.ldy = coup_ld(n1_max); g.N = n1_max;
  g.relu = 0; g.scale = 0.0625f; g.score = 1;
  g.qkv = 0; g.qp = nullptr; g.kp = g.vt = nullptr; g.ldv = 0; g.rows_total = (int)rows; g.status = status;
  g.wscale_inv = nullptr;
  g.single_chain = single_chain_knob();
  for (int i = 0; i < kMaxSegs; ++i) { g.vbase[i] = 0; g.tile_end[i] = 0; }
  g.segs = two_segs(n0_max, n1_max, n_dev);
  g.tile_end[0] = cdiv(n0_max, BM);
  if (prec == 2) return launch_tcp<1, 1>(mA, mA, mWh, mWl, g, cdiv(n1_max, bn), g.tile_end[0], GIMS_PROF_SCORE, st);
  if (prec && persist_enabled()) return launch_tcp<1>(mA, mA, mWh, mWl, g, cdiv(n1_max, bn), g.tile_end[0], GIMS_PROF_SCORE, st);
  if (prec) return launch_tc<bn, 1, 1>(mA, mA, mWh, mWl, g, cdiv(n1_max, bn), g.tile_end[0], GIMS_PROF_SCORE, st);
  return launch_tc<bn, 1>(mA, mA, mWh, mWl, g, cdiv(n1_max, bn), g.tile_end[0], GIMS_PROF_SCORE, st);
}

}  // namespace gims
