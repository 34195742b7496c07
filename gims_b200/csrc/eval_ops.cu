// Caller-side result consumption on the device: the ground-truth correspondence search and the
// precision / recall counts that eval_homography.py:205-229 computes after every matcher call, so that the pair's results
// never travel to the host before they are reduced to three integers.
//
//   gims_gt_matches   = utils/preprocess_utils.py:98-132 `torch_find_matches`: keypoints of image 0 projected through the
//                       homography (warp_keypoints :82-96), then n_iters rounds of { mutual nearest neighbours among the still
//                       unmatched points of both images, accepted where the distance is < dist_thresh }.
//   gims_match_counts = eval_homography.py:224-228: true positives, predicted matches, missed ground-truth matches.
//
// Arithmetic follows the reference's fp32 expressions operation by operation (no FMA contraction in the distance:
// (dx*dx + dy*dy) with both squares rounded, then sqrt), because nearest-neighbour ties and the threshold test decide
// indices.  torch.argmin returns the FIRST minimum; the remaining points are kept in ascending index order by the
// reference (torch.unique), so "first" = lowest index here.
#include <math_constants.h>

#include "common.cuh"

namespace gims {
namespace {

// dest = (H @ [x, y, 1]) / dest_z, fp32, the accumulation order of a 3-term dot product
__global__ void k_gt_project(const float2* __restrict__ kp, int n, const float* __restrict__ H, float2* __restrict__ out) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float x = kp[i].x, y = kp[i].y;
  const float dx = __fadd_rn(__fadd_rn(__fmul_rn(H[0], x), __fmul_rn(H[1], y)), H[2]);
  const float dy = __fadd_rn(__fadd_rn(__fmul_rn(H[3], x), __fmul_rn(H[4], y)), H[5]);
  const float dz = __fadd_rn(__fadd_rn(__fmul_rn(H[6], x), __fmul_rn(H[7], y)), H[8]);
  out[i] = make_float2(__fdiv_rn(dx, dz), __fdiv_rn(dy, dz));
}

__device__ __forceinline__ float gt_dist(float2 a, float2 b) {
  const float dx = __fsub_rn(a.x, b.x), dy = __fsub_rn(a.y, b.y);
  return __fsqrt_rn(__fadd_rn(__fmul_rn(dx, dx), __fmul_rn(dy, dy)));
}

// warp per point of `a` that is still unmatched: nearest unmatched point of `b` (lowest index among equal distances)
__global__ void k_gt_nearest(const float2* __restrict__ a, const int* __restrict__ a_match, int na, const float2* __restrict__ b,
                             const int* __restrict__ b_match, int nb, int* __restrict__ nn, float* __restrict__ nnd) {
  const int i = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (i >= na) return;
  if (a_match[i] >= 0) { if (lane == 0) nn[i] = -1; return; }
  const float2 p = a[i];
  float best = CUDART_INF_F;
  int bj = 0x7fffffff;
  for (int j = lane; j < nb; j += 32) {
    if (b_match[j] >= 0) continue;
    const float d = gt_dist(p, b[j]);
    if (d < best) { best = d; bj = j; }          // ascending j per lane: the first minimum stays
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    const float ob = __shfl_xor_sync(0xffffffffu, best, o);
    const int oj = __shfl_xor_sync(0xffffffffu, bj, o);
    if (ob < best || (ob == best && oj < bj)) { best = ob; bj = oj; }
  }
  if (lane == 0) { nn[i] = (bj == 0x7fffffff) ? -1 : bj; nnd[i] = best; }
}

// point j of image 1 and its nearest point i of image 0 are a ground-truth match if i's nearest point is j and they are
// closer than the threshold.  New matches go to separate arrays: this round's searches saw the old state.
__global__ void k_gt_resolve(const int* __restrict__ nn0, const int* __restrict__ nn1, const float* __restrict__ nnd1, int n1,
                             float thresh, int round, int* __restrict__ new0, int* __restrict__ new1, int* __restrict__ round0) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= n1) return;
  const int i = nn1[j];
  if (i < 0 || nn0[i] != j || !(nnd1[j] < thresh)) return;
  new1[j] = i;
  new0[i] = j;
  round0[i] = round;
}

__global__ void k_gt_commit(int* __restrict__ m, const int* __restrict__ fresh, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n && fresh[i] >= 0) m[i] = fresh[i];
}

__global__ void k_fill_int(int* __restrict__ p, int v, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) p[i] = v;
}

// counts[0] = true positives  (matches0[i] == gt0[i], gt0[i] >= 0)       eval_homography.py:224
// counts[1] = predicted matches (matches0[i] > -1)                       :187, 225
// counts[2] = missed: gt0[i] >= 0 and matches0[i] == -1                  :226
__global__ void k_match_counts(const int64_t* __restrict__ matches0, const int* __restrict__ gt0, int n, const int* __restrict__ n_dev,
                               int* __restrict__ counts) {
  const int live = n_dev ? min(*n_dev, n) : n;
  int tp = 0, pred = 0, fn = 0;
  for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < live; i += gridDim.x * blockDim.x) {
    const long long m = matches0[i];
    const int g = gt0[i];
    pred += m > -1;
    tp += (g >= 0 && m == (long long)g);
    fn += (g >= 0 && m == -1);
  }
#pragma unroll
  for (int o = 16; o; o >>= 1) {
    tp += __shfl_xor_sync(0xffffffffu, tp, o);
    pred += __shfl_xor_sync(0xffffffffu, pred, o);
    fn += __shfl_xor_sync(0xffffffffu, fn, o);
  }
  if ((threadIdx.x & 31) == 0) {
    if (tp) atomicAdd(&counts[0], tp);
    if (pred) atomicAdd(&counts[1], pred);
    if (fn) atomicAdd(&counts[2], fn);
  }
}

}  // namespace
}  // namespace gims

using namespace gims;

extern "C" size_t gims_gt_workspace_bytes(int n0, int n1) {
  size_t a = (size_t)(n0 > 0 ? n0 : 0), b = (size_t)(n1 > 0 ? n1 : 0);
  // projected points | nn0, nnd0, new0 | nn1, nnd1, new1, each 256-byte aligned
  return align_up(a * 8, 256) + 3 * align_up(a * 4, 256) + 3 * align_up(b * 4, 256) + 256;
}

extern "C" int gims_gt_matches(const float* kpts0, int n0, const float* kpts1, int n1, const float* homography_dev,
                               float dist_thresh, int n_iters, void* workspace, size_t workspace_bytes, int* gt0, int* gt1,
                               int* round0, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!kpts0 || !kpts1 || !homography_dev || !workspace || !gt0 || !gt1 || !round0) { set_error("gims_gt_matches: null argument"); return GIMS_ERR_ARG; }
  if (n0 < 1 || n1 < 1 || n_iters < 0) { set_error("gims_gt_matches: bad sizes (%d, %d, %d)", n0, n1, n_iters); return GIMS_ERR_ARG; }
  if (workspace_bytes < gims_gt_workspace_bytes(n0, n1)) { set_error("gims_gt_matches: workspace too small"); return GIMS_ERR_WORKSPACE; }
  Arena a(workspace, workspace_bytes);
  float2* proj = a.take<float2>(n0);
  int* nn0 = a.take<int>(n0); float* nnd0 = a.take<float>(n0); int* new0 = a.take<int>(n0);
  int* nn1 = a.take<int>(n1); float* nnd1 = a.take<float>(n1); int* new1 = a.take<int>(n1);
  const float2* k0 = reinterpret_cast<const float2*>(kpts0);
  const float2* k1 = reinterpret_cast<const float2*>(kpts1);
  k_gt_project<<<cdiv(n0, 256), 256, 0, st>>>(k0, n0, homography_dev, proj);
  GIMS_LAUNCH_OK();
  k_fill_int<<<cdiv(n0, 256), 256, 0, st>>>(gt0, -1, n0); GIMS_LAUNCH_OK();
  k_fill_int<<<cdiv(n0, 256), 256, 0, st>>>(round0, -1, n0); GIMS_LAUNCH_OK();
  k_fill_int<<<cdiv(n1, 256), 256, 0, st>>>(gt1, -1, n1); GIMS_LAUNCH_OK();
  for (int it = 0; it < n_iters; ++it) {
    k_fill_int<<<cdiv(n0, 256), 256, 0, st>>>(new0, -1, n0); GIMS_LAUNCH_OK();
    k_fill_int<<<cdiv(n1, 256), 256, 0, st>>>(new1, -1, n1); GIMS_LAUNCH_OK();
    k_gt_nearest<<<cdiv(n0, 8), 256, 0, st>>>(proj, gt0, n0, k1, gt1, n1, nn0, nnd0); GIMS_LAUNCH_OK();
    k_gt_nearest<<<cdiv(n1, 8), 256, 0, st>>>(k1, gt1, n1, proj, gt0, n0, nn1, nnd1); GIMS_LAUNCH_OK();
    k_gt_resolve<<<cdiv(n1, 256), 256, 0, st>>>(nn0, nn1, nnd1, n1, dist_thresh, it, new0, new1, round0); GIMS_LAUNCH_OK();
    k_gt_commit<<<cdiv(n0, 256), 256, 0, st>>>(gt0, new0, n0); GIMS_LAUNCH_OK();
    k_gt_commit<<<cdiv(n1, 256), 256, 0, st>>>(gt1, new1, n1); GIMS_LAUNCH_OK();
    for (int k = 0; k < 7; ++k) count_launch();
  }
  for (int k = 0; k < 4; ++k) count_launch();
  return GIMS_OK;
}

extern "C" int gims_match_counts(const int64_t* matches0, const int* gt0, int n0_max, const int* n_dev, int* counts_dev, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!matches0 || !gt0 || !counts_dev || n0_max < 1) { set_error("gims_match_counts: bad argument"); return GIMS_ERR_ARG; }
  GIMS_CUDA_OK(cudaMemsetAsync(counts_dev, 0, 3 * sizeof(int), st));
  k_match_counts<<<min(cdiv(n0_max, 256), 64), 256, 0, st>>>(matches0, gt0, n0_max, n_dev, counts_dev);
  GIMS_LAUNCH_OK();
  count_launch();
  return GIMS_OK;
}
