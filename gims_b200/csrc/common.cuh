// Shared helpers for the gims_b200 CUDA library (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>

#include "../../include/gims_b200.h"

namespace gims {

void set_error(const char* fmt, ...);
void count_launch(int n = 1);

#define GIMS_CUDA_OK(expr)                                                                   \
  do {                                                                                       \
    cudaError_t e__ = (expr);                                                                \
    if (e__ != cudaSuccess) {                                                                \
      ::gims::set_error("%s:%d CUDA error %s: %s", __FILE__, __LINE__, #expr,                \
                        cudaGetErrorString(e__));                                            \
      return GIMS_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

#define GIMS_LAUNCH_OK()                                                                     \
  do {                                                                                       \
    ::gims::count_launch();                                                                  \
    cudaError_t e__ = cudaGetLastError();                                                    \
    if (e__ != cudaSuccess) {                                                                \
      ::gims::set_error("%s:%d kernel launch failed: %s", __FILE__, __LINE__,                \
                        cudaGetErrorString(e__));                                            \
      return GIMS_ERR_CUDA;                                                                  \
    }                                                                                        \
  } while (0)

#define GIMS_TRY(expr)                                                                       \
  do {                                                                                       \
    int rc__ = (expr);                                                                       \
    if (rc__ != GIMS_OK) return rc__;                                                        \
  } while (0)

// RAII event pair around a launch of the profiled kernel class (see gims_profile_begin).
struct ProfScope {
  int slot;
  cudaStream_t st;
  ProfScope(int kernel_class, cudaStream_t s);
  ~ProfScope();
};

// Cooperative (grid-synchronising) kernels of different streams must never be partially co-resident with more CTAs in
// total than the GPU has SMs: they are chained through per-device events.  There are two chains ("lanes"): a kernel that
// occupies at most half the SMs joins one lane (alternating), so two such kernels of different streams may run side by
// side; a kernel that needs every SM waits for both lanes and both lanes wait for it.  The scope holds a process-wide
// lock from the wait to the record, so concurrent host threads (one stream each) cannot both chain onto the same
// predecessor.
int coop_chain_wait(cudaStream_t st, int lane);      // lane 0 / 1, or -1 = both
int coop_chain_record(cudaStream_t st, int lane);
int coop_chain_pick_lane();
void coop_chain_lock();
void coop_chain_unlock();
struct CoopChainScope {
  cudaStream_t st;
  int lane;
  int rc;
  CoopChainScope(cudaStream_t s, bool half_gpu) : st(s) {
    coop_chain_lock();
    lane = half_gpu ? coop_chain_pick_lane() : -1;
    rc = coop_chain_wait(st, lane);
  }
  ~CoopChainScope() { if (rc == 0) coop_chain_record(st, lane); coop_chain_unlock(); }
};

constexpr int kD = GIMS_DESC_DIM;     // 256
constexpr int kHeads = GIMS_NUM_HEADS;
constexpr int kHeadDim = kD / kHeads;  // 64

static inline int cdiv(int a, int b) { return (a + b - 1) / b; }
static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

// Bump allocator over a caller-provided workspace.
struct Arena {
  char* base;
  size_t cap;
  size_t off;
  Arena(void* p, size_t c) : base(static_cast<char*>(p)), cap(c), off(0) {}
  template <typename T>
  T* take(size_t count) {
    off = align_up(off, 256);
    T* r = reinterpret_cast<T*>(base + off);
    off += count * sizeof(T);
    return r;
  }
  bool ok() const { return off <= cap; }
};

// Row segments inside one stacked activation buffer: segment i lives in rows [base[i], base[i] + count),
// count = n_ptr[i] ? min(*n_ptr[i], nmax[i]) : nmax[i] (live counts are device scalars).  A pair contributes two segments
// (image 0, image 1); a batch of P pairs stacks 2 P of them, pair-major, so that segment s ^ 1 is the other image of the
// same pair.
constexpr int kMaxSegs = 8;          // up to 4 pairs per launch
struct Segs {
  int base[kMaxSegs];
  int nmax[kMaxSegs];
  const int* n_ptr[kMaxSegs];        // may be null
  int nseg;
};

__device__ __forceinline__ int seg_count(const Segs& s, int i) {
  return s.n_ptr[i] ? min(*s.n_ptr[i], s.nmax[i]) : s.nmax[i];
}
static inline Segs one_seg(int n_max, const int* n_dev) {
  Segs s = {};
  s.nmax[0] = n_max; s.n_ptr[0] = n_dev; s.nseg = 1;
  return s;
}
static inline Segs two_segs(int n0_max, int n1_max, const int* n_dev) {
  Segs s = {};
  s.base[1] = n0_max; s.nmax[0] = n0_max; s.nmax[1] = n1_max;
  s.n_ptr[0] = n_dev; s.n_ptr[1] = n_dev ? n_dev + 1 : nullptr; s.nseg = 2;
  return s;
}
static inline int segs_rows(const Segs& s) { return s.nseg ? s.base[s.nseg - 1] + s.nmax[s.nseg - 1] : 0; }
static inline int segs_nmax(const Segs& s) { int m = 0; for (int i = 0; i < s.nseg; ++i) m = s.nmax[i] > m ? s.nmax[i] : m; return m; }

// ---------------------------------------------------------------------------------------------
// internal (non-ABI) launchers shared between translation units
// ---------------------------------------------------------------------------------------------
struct GemmArgs {
  const float* A0; int lda0; int K0;      // A(r,k) = k < K0 ? A0[r*lda0+k] : A1[r*lda1 + k-K0]
  const float* A1; int lda1; int K1;
  const float* W;                          // [N][K0+K1] row-major (K-major), nn.Linear / Conv1d(k=1) layout
  const float* bias;                       // [N] or null
  const float* R; int ldr;                 // residual added in the epilogue (may alias Y) or null
  float* Y; int ldy;
  int N;
  int relu;
  Segs segs;
  unsigned* status;                        // fp16-operand kernels: GIMS_STATUS_FP16_RANGE is OR-ed in (may be null)
};
int launch_gemm(const GemmArgs& a, cudaStream_t st);

// tensor-core planes of one weight matrix [N][K] (gims_b200/packing.py)
struct WPlanes {
  const float* hi32;            // tf32(W)
  const float* lo32;            // W - tf32(W)
  const void* h16;              // fp16(W * 2^e)
  const void* l16;              // fp16(W * 2^e - h16)
  const float* sinv;            // one float: 2^-e
};

// tensor-core (tcgen05, 3xTF32) variants — gemm_tc.cu
struct QkvPlanes {            // outputs of the QKV projection in the layout the attention kernels consume
  float* qp;                  // [rows_total][256] fp32, scaled by log2(e)/8
  void* kp;                   // tf32: float [2][rows_total][256];  16-bit: [planes][rows_total][256]
  void* vt;                   // tf32: float [2][256][ldv];         16-bit: [planes][256][ldv]
                              // key columns of segment s at [vbase[s], vbase[s] + n_s)
  int ldv;                    // multiple of 64
  int vbase[kMaxSegs];        // first key column of segment s: sum of round_up(nmax, 64) of the segments before it
  int fmt;                    // -1: tf32 hi / lo planes (attention_tc.cu); 0: fp16, 1: bf16 (attention_f16.cu)
  int planes;                 // 16-bit: 2 = hi + lo (fp32-class), 1 = single plane (bf16 variant)
  unsigned* status;           // fp16: GIMS_STATUS_FP16_RANGE is OR-ed in when a value reaches 32768 (may be null)
};
int launch_gemm_tc(const GemmArgs& a, const WPlanes& w, int prec, cudaStream_t st, const QkvPlanes* qkv = nullptr);
int launch_attention_tc(const QkvPlanes& pl, float* out, const Segs& segs, int cross, cudaStream_t st);
int launch_attention_f16(const QkvPlanes& pl, float* out, const Segs& segs, int cross, cudaStream_t st);
int set_attention_trace(long long* dev_buf);
int set_gemm_trace(long long* dev_buf);
// row pitch (floats) of the couplings matrix (n0_max+1) x (n1_max+1): rows start 16-byte aligned
static inline int coup_ld(int n1_max) { return (n1_max + 1 + 3) & ~3; }
// key-column layout of the transposed V planes: fills vbase[], returns ldv
static inline int attn_vt_layout(const Segs& s, int* vbase) {
  int c = 0;
  for (int i = 0; i < s.nseg; ++i) { vbase[i] = c; c += (s.nmax[i] + 63) & ~63; }
  return c;
}
int launch_score_gemm_tc(const float* mdesc, int n0_max, int n1_max, const int* n_dev, float* planes, float* couplings,
                         int prec, unsigned* status, cudaStream_t st);
int launch_split_planes(const float* x, float* hi, float* lo, size_t n, cudaStream_t st);
int launch_score_border(int n0_max, int n1_max, const int* n_dev, const float* bin_score, float* couplings,
                        cudaStream_t st);

// scores GEMM: couplings[i][j] = scale * <A_i, B_j>, with the dustbin border (gmatcher.py:59-60)
int launch_score_gemm(const float* mdesc, int n0_max, int n1_max, const int* n_dev, const float* bin_score,
                      float* couplings, cudaStream_t st);

// flash attention over the stacked QKV buffer [rows][768] (Q|K|V, each head-major h*64+d)
int launch_attention(const float* qkv, float* out, const Segs& segs, int cross, cudaStream_t st);

int launch_sage_aggregate(const float* src, int lds, int width, const int* indptr, const int* indices,
                          int n_max, const int* n_dev, const float* self_add, int ldself, const float* bias,
                          int relu, float* out, int ldo, cudaStream_t st);

int launch_kenc_first(const float* kpts, int n_max, const int* n_dev, float img_w, float img_h, const float* W,
                      const float* b, int cout, float* out, cudaStream_t st);

}  // namespace gims
