// fp32 CUDA-core GEMM family (exact fp32 FMA arithmetic; the tensor-core path lives in gemm_tc.cu).
//
//   Y[r][o] = epi( sum_k A(r,k) * W[o][k] + bias[o] )          A, W both K-major (row-major, K contiguous)
//
// which is nn.Conv1d(kernel_size=1) / nn.Linear of the reference on node-major activations:
//   Q/K/V/merge projections  models/gmatcher.py:105-114      MLP convs (BN folded)  gmatcher.py:11-24, 116-125
//   SAGE fc_self/fc_neigh    dgl SAGEConv, gmatcher.py:145-162   final_proj          gmatcher.py:273
//   score matrix             gmatcher.py:274-275 (einsum 'bdn,bdm->bnm' / sqrt(D)) + dustbin border :59-60
// `A` may be the K-concatenation of two buffers (torch.cat([x, message], dim=1), gmatcher.py:125).
// Rows are processed per image segment with device-side row counts (ragged N' after AGC pruning).
#include "common.cuh"

namespace gims {

namespace {

constexpr int BM = 128, BN = 128, BK = 16;
constexpr int kPad = 4;

enum { kModeLinear = 0, kModeScore = 1 };

struct GemmKernArgs {
  const float* A0; int lda0; int K0;
  const float* A1; int lda1; int K1;
  const float* W; int ldw;
  const float* bias;
  const float* R; int ldr;
  float* Y; int ldy;
  int N;                    // columns (outputs); for score mode the max, live count from n_dev[1]
  int relu;
  float scale;
  Segs segs;
  int tile_end[kMaxSegs];   // cumulative number of row tiles up to and including segment s
};

template <int MODE>
__global__ void __launch_bounds__(256) k_gemm_simt(GemmKernArgs g) {
  __shared__ __align__(16) float As[2][BK][BM + kPad];
  __shared__ __align__(16) float Bs[2][BK][BN + kPad];

  int seg, tile;
  if (MODE == kModeScore) { seg = 0; tile = blockIdx.y; }
  else {
    seg = 0;
    while (seg + 1 < g.segs.nseg && (int)blockIdx.y >= g.tile_end[seg]) ++seg;
    tile = blockIdx.y - (seg ? g.tile_end[seg - 1] : 0);
  }
  int rows = seg_count(g.segs, seg);
  int ncols = g.N;
  if (MODE == kModeScore) ncols = seg_count(g.segs, 1);
  int r0 = tile * BM, c0 = blockIdx.x * BN;
  if (r0 >= rows || c0 >= ncols) return;
  int rbase = g.segs.base[seg];
  int wbase = (MODE == kModeScore) ? g.segs.base[1] : 0;   // score mode: "weights" are image-1 rows

  int tid = threadIdx.x;
  int tx = tid & 15, ty = tid >> 4;
  int lrow = tid >> 2, lk = (tid & 3) * 4;
  int K = g.K0 + g.K1;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.f;

  float4 ra[2], rb[2];
  auto gload = [&](int k0) {
    const float* Ap; int lda; int kk;
    if (k0 < g.K0) { Ap = g.A0; lda = g.lda0; kk = k0; } else { Ap = g.A1; lda = g.lda1; kk = k0 - g.K0; }
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int r = r0 + lrow + 64 * h;
      ra[h] = (r < rows) ? *reinterpret_cast<const float4*>(Ap + (size_t)(rbase + r) * lda + kk + lk)
                         : make_float4(0.f, 0.f, 0.f, 0.f);
      int c = c0 + lrow + 64 * h;
      rb[h] = (c < ncols) ? *reinterpret_cast<const float4*>(g.W + (size_t)(wbase + c) * g.ldw + k0 + lk)
                          : make_float4(0.f, 0.f, 0.f, 0.f);
    }
  };
  auto sstore = [&](int buf) {
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      int m = lrow + 64 * h;
      As[buf][lk + 0][m] = ra[h].x; As[buf][lk + 1][m] = ra[h].y; As[buf][lk + 2][m] = ra[h].z; As[buf][lk + 3][m] = ra[h].w;
      Bs[buf][lk + 0][m] = rb[h].x; Bs[buf][lk + 1][m] = rb[h].y; Bs[buf][lk + 2][m] = rb[h].z; Bs[buf][lk + 3][m] = rb[h].w;
    }
  };

  gload(0);
  sstore(0);
  __syncthreads();
  int nk = K / BK;
  for (int kt = 0; kt < nk; ++kt) {
    int buf = kt & 1;
    if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      float a[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      float b[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(a[i], b[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      sstore(buf ^ 1);
      __syncthreads();
    }
  }

#pragma unroll
  for (int i = 0; i < 8; ++i) {
    int r = r0 + ((i < 4) ? (ty * 4 + i) : (64 + ty * 4 + i - 4));
    if (r >= rows) continue;
#pragma unroll
    for (int jh = 0; jh < 2; ++jh) {
      int c = c0 + jh * 64 + tx * 4;
      if (c >= ncols) continue;
      float v[4] = {acc[i][jh * 4 + 0], acc[i][jh * 4 + 1], acc[i][jh * 4 + 2], acc[i][jh * 4 + 3]};
      if (MODE == kModeScore) {
        float* y = g.Y + (size_t)r * g.ldy + c;
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (c + q < ncols) y[q] = v[q] * g.scale;
      } else {
        size_t row = (size_t)(rbase + r);
        if (g.bias) {
          float4 bb = *reinterpret_cast<const float4*>(g.bias + c);
          v[0] += bb.x; v[1] += bb.y; v[2] += bb.z; v[3] += bb.w;
        }
        if (g.R) {
          float4 rr = *reinterpret_cast<const float4*>(g.R + row * g.ldr + c);
          v[0] += rr.x; v[1] += rr.y; v[2] += rr.z; v[3] += rr.w;
        }
        if (g.relu) {
#pragma unroll
          for (int q = 0; q < 4; ++q) v[q] = fmaxf(v[q], 0.f);
        }
        *reinterpret_cast<float4*>(g.Y + row * g.ldy + c) = make_float4(v[0], v[1], v[2], v[3]);
      }
    }
  }
}

// dustbin row/column of the couplings matrix (torch.cat at gmatcher.py:59-60)
__global__ void k_score_border(float* __restrict__ Z, int ld, Segs segs, const float* __restrict__ bin_score) {
  int n0 = seg_count(segs, 0), n1 = seg_count(segs, 1);
  float a = *bin_score;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  if (t <= n1) Z[(size_t)n0 * ld + t] = a;
  if (t < n0) Z[(size_t)t * ld + n1] = a;
}

}  // namespace

int launch_gemm(const GemmArgs& a, cudaStream_t st) {
  int K = a.K0 + a.K1;
  if (a.K0 % BK || a.K1 % BK || K == 0 || a.N % 4) {
    set_error("launch_gemm: unsupported shape K0=%d K1=%d N=%d", a.K0, a.K1, a.N);
    return GIMS_ERR_ARG;
  }
  GemmKernArgs g;
  g.A0 = a.A0; g.lda0 = a.lda0; g.K0 = a.K0;
  g.A1 = a.A1; g.lda1 = a.lda1; g.K1 = a.K1;
  g.W = a.W; g.ldw = K; g.bias = a.bias; g.R = a.R; g.ldr = a.ldr; g.Y = a.Y; g.ldy = a.ldy;
  g.N = a.N; g.relu = a.relu; g.scale = 1.f; g.segs = a.segs;
  int tiles = 0;
  for (int i = 0; i < a.segs.nseg; ++i) { tiles += cdiv(a.segs.nmax[i], BM); g.tile_end[i] = tiles; }
  if (tiles == 0) return GIMS_OK;
  ProfScope prof(GIMS_PROF_GEMM, st);
  k_gemm_simt<kModeLinear><<<dim3(cdiv(a.N, BN), tiles), 256, 0, st>>>(g);
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}

int launch_score_gemm(const float* mdesc, int n0_max, int n1_max, const int* n_dev, const float* bin_score,
                      float* couplings, cudaStream_t st) {
  GemmKernArgs g;
  g.A0 = mdesc; g.lda0 = kD; g.K0 = kD;
  g.A1 = nullptr; g.lda1 = 0; g.K1 = 0;
  g.W = mdesc; g.ldw = kD; g.bias = nullptr; g.R = nullptr; g.ldr = 0;
  g.Y = couplings; g.ldy = coup_ld(n1_max);
  g.N = n1_max; g.relu = 0; g.scale = 0.0625f;    // 1/sqrt(256), exact
  g.segs = two_segs(n0_max, n1_max, n_dev);
  g.tile_end[0] = cdiv(n0_max, BM);
  ProfScope prof(GIMS_PROF_SCORE, st);
  k_gemm_simt<kModeScore><<<dim3(cdiv(n1_max, BN), g.tile_end[0]), 256, 0, st>>>(g);
  GIMS_LAUNCH_OK();
  return launch_score_border(n0_max, n1_max, n_dev, bin_score, couplings, st);
}

int launch_score_border(int n0_max, int n1_max, const int* n_dev, const float* bin_score, float* couplings,
                        cudaStream_t st) {
  Segs s = two_segs(n0_max, n1_max, n_dev);
  int m = (n0_max > n1_max ? n0_max : n1_max) + 1;
  k_score_border<<<cdiv(m, 256), 256, 0, st>>>(couplings, coup_ld(n1_max), s, bin_score);
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}

}  // namespace gims
