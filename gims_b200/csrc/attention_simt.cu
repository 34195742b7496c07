// Dense multi-head attention, flash style, fp32 CUDA-core arithmetic (stage-1 / reference-grade path).
//
// Replaces `attention` + the einsums of MultiHeadedAttention (models/gmatcher.py:35-39, 108-113):
//   scores = q^T k / sqrt(64) per head, softmax over the source keypoints, out = P v
// without materialising the (4, N, M) probability tensor.  Q|K|V come from one stacked buffer
// [rows][768] whose channels were de-interleaved at pack time (reference channel c = d*4 + h,
// gmatcher.py:111; packed channel = h*64 + d).  Image s attends to image s ('self') or 1-s ('cross').
#include <math_constants.h>

#include "common.cuh"

namespace gims {

namespace {

constexpr int TQ = 64, TK = 64, HD = 64;
constexpr int LDT = TQ + 4;   // padded leading dim of transposed tiles

struct AttnSmem {
  float Qt[HD][LDT];    // Qt[d][i]
  float Kt[HD][LDT];    // Kt[d][j]; reused as Pt[j][i] after the score tile is computed
  float Vs[TK][HD];     // Vs[j][d]
};

__global__ void __launch_bounds__(256) k_attention_simt(const float* __restrict__ qkv, float* __restrict__ out, Segs segs,
                                                        int cross) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AttnSmem& sm = *reinterpret_cast<AttnSmem*>(smem_raw);
  int img = blockIdx.z, head = blockIdx.y;
  int src = cross ? (img ^ 1) : img;
  int nq = seg_count(segs, img), nk = seg_count(segs, src);
  int q0 = blockIdx.x * TQ;
  if (q0 >= nq) return;
  const float* Qg = qkv + (size_t)segs.base[img] * (3 * kD) + head * HD;
  const float* Kg = qkv + (size_t)segs.base[src] * (3 * kD) + kD + head * HD;
  const float* Vg = qkv + (size_t)segs.base[src] * (3 * kD) + 2 * kD + head * HD;

  int tid = threadIdx.x, tx = tid & 15, ty = tid >> 4;
  // tile loader mapping: 64 rows x 16 float4; thread -> (row = tid/16 + 16*p, c4 = tid%16)
  int lr = tid >> 4, lc = (tid & 15) * 4;
#pragma unroll
  for (int p = 0; p < 4; ++p) {
    int r = lr + 16 * p;
    float4 v = make_float4(0.f, 0.f, 0.f, 0.f);
    if (q0 + r < nq) v = *reinterpret_cast<const float4*>(Qg + (size_t)(q0 + r) * (3 * kD) + lc);
    sm.Qt[lc + 0][r] = v.x; sm.Qt[lc + 1][r] = v.y; sm.Qt[lc + 2][r] = v.z; sm.Qt[lc + 3][r] = v.w;
  }

  float m_run[4], l_run[4], o[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    m_run[i] = -CUDART_INF_F; l_run[i] = 0.f;
#pragma unroll
    for (int d = 0; d < 4; ++d) o[i][d] = 0.f;
  }
  const float sc = 0.125f * 1.4426950408889634f;   // 1/sqrt(64) * log2(e)

  for (int k0 = 0; k0 < nk; k0 += TK) {
    __syncthreads();   // previous tile fully consumed (Pt / Vs)
#pragma unroll
    for (int p = 0; p < 4; ++p) {
      int r = lr + 16 * p;
      float4 kv = make_float4(0.f, 0.f, 0.f, 0.f), vv = kv;
      if (k0 + r < nk) {
        kv = *reinterpret_cast<const float4*>(Kg + (size_t)(k0 + r) * (3 * kD) + lc);
        vv = *reinterpret_cast<const float4*>(Vg + (size_t)(k0 + r) * (3 * kD) + lc);
      }
      sm.Kt[lc + 0][r] = kv.x; sm.Kt[lc + 1][r] = kv.y; sm.Kt[lc + 2][r] = kv.z; sm.Kt[lc + 3][r] = kv.w;
      *reinterpret_cast<float4*>(&sm.Vs[r][lc]) = vv;
    }
    __syncthreads();
    // S tile: rows i = ty*4+a, cols j = tx*4+b
    float s[4][4];
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int b = 0; b < 4; ++b) s[a][b] = 0.f;
#pragma unroll 8
    for (int d = 0; d < HD; ++d) {
      float4 qa = *reinterpret_cast<const float4*>(&sm.Qt[d][ty * 4]);
      float4 kb = *reinterpret_cast<const float4*>(&sm.Kt[d][tx * 4]);
      float qv[4] = {qa.x, qa.y, qa.z, qa.w}, kv[4] = {kb.x, kb.y, kb.z, kb.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int b = 0; b < 4; ++b) s[a][b] = fmaf(qv[a], kv[b], s[a][b]);
    }
    // online softmax (base-2 domain)
    float p[4][4], alpha[4];
#pragma unroll
    for (int a = 0; a < 4; ++a) {
      float mx = -CUDART_INF_F;
#pragma unroll
      for (int b = 0; b < 4; ++b) {
        s[a][b] = (k0 + tx * 4 + b < nk) ? s[a][b] * sc : -CUDART_INF_F;
        mx = fmaxf(mx, s[a][b]);
      }
#pragma unroll
      for (int off = 8; off; off >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
      float m_new = fmaxf(m_run[a], mx);     // finite: every tile holds at least one live key
      alpha[a] = exp2f(m_run[a] - m_new);
      float rs = 0.f;
#pragma unroll
      for (int b = 0; b < 4; ++b) { p[a][b] = exp2f(s[a][b] - m_new); rs += p[a][b]; }
#pragma unroll
      for (int off = 8; off; off >>= 1) rs += __shfl_xor_sync(0xffffffffu, rs, off);
      l_run[a] = l_run[a] * alpha[a] + rs;
      m_run[a] = m_new;
    }
    __syncthreads();   // everyone done reading Kt
    float (*Pt)[LDT] = sm.Kt;
#pragma unroll
    for (int b = 0; b < 4; ++b)
      *reinterpret_cast<float4*>(&Pt[tx * 4 + b][ty * 4]) = make_float4(p[0][b], p[1][b], p[2][b], p[3][b]);
    __syncthreads();
    // O tile: rows i = ty*4+a, dims d = tx*4+c
#pragma unroll
    for (int a = 0; a < 4; ++a)
#pragma unroll
      for (int c = 0; c < 4; ++c) o[a][c] *= alpha[a];
#pragma unroll 8
    for (int j = 0; j < TK; ++j) {
      float4 pa = *reinterpret_cast<const float4*>(&Pt[j][ty * 4]);
      float4 vb = *reinterpret_cast<const float4*>(&sm.Vs[j][tx * 4]);
      float pv[4] = {pa.x, pa.y, pa.z, pa.w}, vv[4] = {vb.x, vb.y, vb.z, vb.w};
#pragma unroll
      for (int a = 0; a < 4; ++a)
#pragma unroll
        for (int c = 0; c < 4; ++c) o[a][c] = fmaf(pv[a], vv[c], o[a][c]);
    }
  }
#pragma unroll
  for (int a = 0; a < 4; ++a) {
    int r = q0 + ty * 4 + a;
    if (r >= nq) continue;
    float inv = 1.f / l_run[a];
    float4 v = make_float4(o[a][0] * inv, o[a][1] * inv, o[a][2] * inv, o[a][3] * inv);
    *reinterpret_cast<float4*>(out + (size_t)(segs.base[img] + r) * kD + head * HD + tx * 4) = v;
  }
}

}  // namespace

int launch_attention(const float* qkv, float* out, const Segs& s, int cross, cudaStream_t st) {
  GIMS_CUDA_OK(cudaFuncSetAttribute(k_attention_simt, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                    (int)sizeof(AttnSmem)));
  ProfScope prof(GIMS_PROF_ATTENTION, st);
  k_attention_simt<<<dim3(cdiv(segs_nmax(s), TQ), kHeads, s.nseg), 256, sizeof(AttnSmem), st>>>(qkv, out, s, cross);
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}

}  // namespace gims
