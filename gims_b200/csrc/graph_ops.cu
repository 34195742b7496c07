// Per-node kernels around the dense contractions:
//   * GraphSAGE mean aggregation over the AGC CSR (dgl SAGEConv('mean'), models/gmatcher.py:145-162):
//     warp per destination node, 128-bit loads, neighbours summed in ascending order, then / degree.
//   * first layer of the keypoint encoder fused with normalize_keypoints (gmatcher.py:26-33, 87-97).
#include "common.cuh"

namespace gims {

namespace {

// out[i][c] = act( self[i][c] + bias[c] + mean_{j in N(i)} src[j][c] ), c < width (width % 128 == 0)
__global__ void __launch_bounds__(256) k_sage_aggregate(const float* __restrict__ src, int lds, int width,
                                                        const int* __restrict__ indptr, const int* __restrict__ indices,
                                                        int n_max, const int* __restrict__ n_dev,
                                                        const float* __restrict__ self_add, int ldself,
                                                        const float* __restrict__ bias, int relu,
                                                        float* __restrict__ out, int ldo) {
  int n = n_dev ? min(*n_dev, n_max) : n_max;
  int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  int lane = threadIdx.x & 31;
  if (i >= n) return;
  int e0 = indptr[i], e1 = indptr[i + 1];
  float deg = (float)max(e1 - e0, 1);
  for (int c = lane * 4; c < width; c += 128) {
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    int e = e0;
    for (; e + 4 <= e1; e += 4) {
      int j0 = indices[e], j1 = indices[e + 1], j2 = indices[e + 2], j3 = indices[e + 3];
      float4 v0 = *reinterpret_cast<const float4*>(src + (size_t)j0 * lds + c);
      float4 v1 = *reinterpret_cast<const float4*>(src + (size_t)j1 * lds + c);
      float4 v2 = *reinterpret_cast<const float4*>(src + (size_t)j2 * lds + c);
      float4 v3 = *reinterpret_cast<const float4*>(src + (size_t)j3 * lds + c);
      s.x += v0.x; s.y += v0.y; s.z += v0.z; s.w += v0.w;
      s.x += v1.x; s.y += v1.y; s.z += v1.z; s.w += v1.w;
      s.x += v2.x; s.y += v2.y; s.z += v2.z; s.w += v2.w;
      s.x += v3.x; s.y += v3.y; s.z += v3.z; s.w += v3.w;
    }
    for (; e < e1; ++e) {
      float4 v = *reinterpret_cast<const float4*>(src + (size_t)indices[e] * lds + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    s.x /= deg; s.y /= deg; s.z /= deg; s.w /= deg;
    if (self_add) {
      float4 v = *reinterpret_cast<const float4*>(self_add + (size_t)i * ldself + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    if (bias) {
      float4 v = *reinterpret_cast<const float4*>(bias + c);
      s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w;
    }
    if (relu) { s.x = fmaxf(s.x, 0.f); s.y = fmaxf(s.y, 0.f); s.z = fmaxf(s.z, 0.f); s.w = fmaxf(s.w, 0.f); }
    *reinterpret_cast<float4*>(out + (size_t)i * ldo + c) = s;
  }
}

// out[i][o] = relu(W[o][0]*xn + W[o][1]*yn + b[o]); (xn,yn) = normalize_keypoints(kpts[i]) — gmatcher.py:26-33
__global__ void k_kenc_first(const float2* __restrict__ kpts, int n_max, const int* __restrict__ n_dev, float img_w,
                             float img_h, const float* __restrict__ W, const float* __restrict__ b, int cout,
                             float* __restrict__ out) {
  int n = n_dev ? min(*n_dev, n_max) : n_max;
  int t = blockIdx.x * blockDim.x + threadIdx.x;
  int i = t / cout, o = t - i * cout;
  if (i >= n) return;
  float2 p = kpts[i];
  float cx = img_w / 2.f, cy = img_h / 2.f;
  float sc = fmaxf(img_w, img_h) * 0.7f;
  float xn = (p.x - cx) / sc, yn = (p.y - cy) / sc;
  float v = fmaf(W[2 * o + 1], yn, W[2 * o] * xn) + b[o];
  out[(size_t)i * cout + o] = fmaxf(v, 0.f);
}

}  // namespace

int launch_sage_aggregate(const float* src, int lds, int width, const int* indptr, const int* indices, int n_max,
                          const int* n_dev, const float* self_add, int ldself, const float* bias, int relu, float* out,
                          int ldo, cudaStream_t st) {
  if (width % 128) { set_error("sage aggregate: width %d not a multiple of 128", width); return GIMS_ERR_ARG; }
  ProfScope prof(GIMS_PROF_SAGE_GATHER, st);
  k_sage_aggregate<<<cdiv(n_max, 8), 256, 0, st>>>(src, lds, width, indptr, indices, n_max, n_dev, self_add, ldself,
                                                  bias, relu, out, ldo);
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}

int launch_kenc_first(const float* kpts, int n_max, const int* n_dev, float img_w, float img_h, const float* W,
                      const float* b, int cout, float* out, cudaStream_t st) {
  long long total = (long long)n_max * cout;
  k_kenc_first<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float2*>(kpts), n_max, n_dev,
                                                              img_w, img_h, W, b, cout, out);
  GIMS_LAUNCH_OK();
  return GIMS_OK;
}

}  // namespace gims
