// C-ABI entry points of libgims_b200.so (declared in include/gims_b200.h): model handle, per-stage
// forwards and the whole-pair forward that chains them (GMatcher.forward, models/gmatcher.py:219-307).
#include <stdarg.h>
#include <string.h>

#include <atomic>
#include <mutex>
#include <new>
#include <vector>

#include "common.cuh"

namespace gims {

static thread_local char g_err[512] = "";
static std::atomic<long long> g_launches{0};

void set_error(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
}
void count_launch(int n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

// ---- profiler ---------------------------------------------------------------------------------
static std::mutex g_prof_mu;
static int g_prof_class = GIMS_PROF_NONE;
static int g_prof_max = 0;
static std::vector<cudaEvent_t> g_prof_ev;   // pairs (start, stop)

ProfScope::ProfScope(int kernel_class, cudaStream_t s) : slot(-1), st(s) {
  if (g_prof_class != kernel_class) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (g_prof_class != kernel_class || (int)g_prof_ev.size() / 2 >= g_prof_max) return;
  cudaEvent_t a, b;
  if (cudaEventCreate(&a) != cudaSuccess) return;
  if (cudaEventCreate(&b) != cudaSuccess) { cudaEventDestroy(a); return; }
  slot = (int)g_prof_ev.size();
  g_prof_ev.push_back(a);
  g_prof_ev.push_back(b);
  cudaEventRecord(a, st);
}
ProfScope::~ProfScope() {
  if (slot < 0) return;
  std::lock_guard<std::mutex> lk(g_prof_mu);
  if (slot + 1 < (int)g_prof_ev.size()) cudaEventRecord(g_prof_ev[slot + 1], st);
}

// ---- cooperative-kernel chain -------------------------------------------------------------------
static std::mutex g_coop_mu;
static cudaEvent_t g_coop_ev[64];
static bool g_coop_has[64];

static std::mutex g_coop_launch_mu;
void coop_chain_lock() { g_coop_launch_mu.lock(); }
void coop_chain_unlock() { g_coop_launch_mu.unlock(); }

int coop_chain_wait(cudaStream_t st) {
  int dev = 0;
  GIMS_CUDA_OK(cudaGetDevice(&dev));
  std::lock_guard<std::mutex> lk(g_coop_mu);
  if (dev < 64 && g_coop_has[dev]) GIMS_CUDA_OK(cudaStreamWaitEvent(st, g_coop_ev[dev], 0));
  return GIMS_OK;
}
int coop_chain_record(cudaStream_t st) {
  int dev = 0;
  GIMS_CUDA_OK(cudaGetDevice(&dev));
  if (dev >= 64) return GIMS_OK;
  std::lock_guard<std::mutex> lk(g_coop_mu);
  if (!g_coop_has[dev]) {
    GIMS_CUDA_OK(cudaEventCreateWithFlags(&g_coop_ev[dev], cudaEventDisableTiming));
    g_coop_has[dev] = true;
  }
  GIMS_CUDA_OK(cudaEventRecord(g_coop_ev[dev], st));
  return GIMS_OK;
}

}  // namespace gims

extern "C" int gims_profile_begin(int kernel_class, int max_launches) {
  std::lock_guard<std::mutex> lk(gims::g_prof_mu);
  for (cudaEvent_t e : gims::g_prof_ev) cudaEventDestroy(e);
  gims::g_prof_ev.clear();
  gims::g_prof_class = kernel_class;
  gims::g_prof_max = max_launches;
  return GIMS_OK;
}

extern "C" int gims_profile_end(double* total_ms, int* launches) {
  std::lock_guard<std::mutex> lk(gims::g_prof_mu);
  gims::g_prof_class = GIMS_PROF_NONE;
  double tot = 0.0;
  int n = 0;
  for (size_t i = 0; i + 1 < gims::g_prof_ev.size(); i += 2) {
    float ms = 0.f;
    if (cudaEventSynchronize(gims::g_prof_ev[i + 1]) == cudaSuccess &&
        cudaEventElapsedTime(&ms, gims::g_prof_ev[i], gims::g_prof_ev[i + 1]) == cudaSuccess) {
      tot += ms;
      ++n;
    }
  }
  for (cudaEvent_t e : gims::g_prof_ev) cudaEventDestroy(e);
  gims::g_prof_ev.clear();
  if (total_ms) *total_ms = tot;
  if (launches) *launches = n;
  return GIMS_OK;
}

using namespace gims;

// Packed weight blobs, in the order gims_b200/packing.py emits them.
struct Wt {                    // one weight matrix [N][K]: fp32 + its tensor-core planes
  const float* w;
  const float* hi;             // tf32(W)
  const float* lo;             // W - tf32(W)
};

struct gims_model {
  gims_config cfg;
  const float* bin_score;
  Wt kenc_w[GIMS_MAX_KENC];
  const float* kenc_b[GIMS_MAX_KENC];
  Wt sage_w[3];                // L0: [256][256] rows 0..127 fc_neigh, 128..255 fc_self
                               // L1: [128][256] cols 0..127 fc_self, 128..255 fc_neigh;  L2: [256][256] likewise
  const float* sage_b[3];
  Wt wqkv[GIMS_MAX_LAYERS];    // [768][256]  Q|K|V rows, head-major
  const float* bqkv[GIMS_MAX_LAYERS];
  Wt w1[GIMS_MAX_LAYERS];      // [512][512]  BN folded, `merge` composed in: input = [x | attention output]
  const float* b1[GIMS_MAX_LAYERS];
  Wt w2[GIMS_MAX_LAYERS];      // [256][512]
  const float* b2[GIMS_MAX_LAYERS];
  Wt wfinal;
  const float* bfinal;
};

static std::atomic<int> g_gemm_mode{GIMS_GEMM_TC_F16};

extern "C" int gims_version(void) { return 100; }
extern "C" const char* gims_last_error(void) { return g_err; }
extern "C" long long gims_launch_count(void) { return g_launches.load(); }

extern "C" int gims_set_gemm_mode(int mode) {
  if (mode < GIMS_GEMM_SIMT || mode > GIMS_GEMM_BF16) { set_error("gims_set_gemm_mode: bad mode %d", mode); return GIMS_ERR_ARG; }
  g_gemm_mode.store(mode);
  return GIMS_OK;
}
extern "C" int gims_get_gemm_mode(void) { return g_gemm_mode.load(); }

extern "C" int gims_packed_blob_count(const gims_config* c) {
  if (!c) return -1;
  // every weight matrix contributes 3 blobs (W, W_hi, W_lo), every bias 1
  return 1 + 4 * c->kenc_num + 4 * 3 + 12 * c->num_layers + 4;
}

extern "C" int gims_model_create(const gims_config* c, const float* packed, const int64_t* off, int n_off,
                                 gims_model** out) {
  if (!c || !packed || !off || !out) { set_error("gims_model_create: null argument"); return GIMS_ERR_ARG; }
  if (c->descriptor_dim != kD) { set_error("gims_model_create: descriptor_dim %d unsupported (built for 256)", c->descriptor_dim); return GIMS_ERR_ARG; }
  if (c->num_layers < 0 || c->num_layers > GIMS_MAX_LAYERS || c->kenc_num < 2 || c->kenc_num > GIMS_MAX_KENC) {
    set_error("gims_model_create: bad layer counts"); return GIMS_ERR_ARG;
  }
  if (c->kenc_dims[0] != 2 || c->kenc_dims[c->kenc_num] != kD) { set_error("gims_model_create: kenc must map 2 -> 256"); return GIMS_ERR_ARG; }
  for (int i = 1; i < c->kenc_num; ++i)
    if (c->kenc_dims[i] % 16) { set_error("gims_model_create: kenc width %d not a multiple of 16", c->kenc_dims[i]); return GIMS_ERR_ARG; }
  if (n_off != gims_packed_blob_count(c)) { set_error("gims_model_create: expected %d blobs, got %d", gims_packed_blob_count(c), n_off); return GIMS_ERR_ARG; }
  gims_model* m = new (std::nothrow) gims_model;
  if (!m) { set_error("gims_model_create: out of host memory"); return GIMS_ERR_ARG; }
  m->cfg = *c;
  int k = 0;
  auto next = [&]() { return packed + off[k++]; };
  auto next_w = [&]() { Wt w; w.w = next(); w.hi = next(); w.lo = next(); return w; };
  m->bin_score = next();
  for (int i = 0; i < c->kenc_num; ++i) { m->kenc_w[i] = next_w(); m->kenc_b[i] = next(); }
  for (int i = 0; i < 3; ++i) { m->sage_w[i] = next_w(); m->sage_b[i] = next(); }
  for (int l = 0; l < c->num_layers; ++l) {
    m->wqkv[l] = next_w(); m->bqkv[l] = next();
    m->w1[l] = next_w(); m->b1[l] = next(); m->w2[l] = next_w(); m->b2[l] = next();
  }
  m->wfinal = next_w(); m->bfinal = next();
  *out = m;
  return GIMS_OK;
}

extern "C" void gims_model_destroy(gims_model* m) { delete m; }

static Segs one_seg(int n_max, const int* n_dev) {
  Segs s; s.base[0] = 0; s.base[1] = 0; s.nmax[0] = n_max; s.nmax[1] = 0; s.n_dev = n_dev; s.nseg = 1; return s;
}
static Segs two_segs(int n0_max, int n1_max, const int* n_dev) {
  Segs s; s.base[0] = 0; s.base[1] = n0_max; s.nmax[0] = n0_max; s.nmax[1] = n1_max; s.n_dev = n_dev; s.nseg = 2; return s;
}

// Y = epi(A W^T + bias): tcgen05 3xTF32 kernel when the shape allows and the mode asks for it, else fp32 SIMT.
static int gemm(const float* A0, int lda0, int K0, const float* A1, int lda1, int K1, const Wt& W, const float* bias,
                const float* R, int ldr, float* Y, int ldy, int N, int relu, Segs s, int mode, cudaStream_t st) {
  GemmArgs g;
  g.A0 = A0; g.lda0 = lda0; g.K0 = K0; g.A1 = A1; g.lda1 = lda1; g.K1 = K1; g.W = W.w; g.bias = bias;
  g.R = R; g.ldr = ldr; g.Y = Y; g.ldy = ldy; g.N = N; g.relu = relu; g.segs = s;
  bool tc_ok = W.hi && W.lo && K0 % 32 == 0 && K1 % 32 == 0 && N % 32 == 0 && lda0 % 4 == 0 && (K1 == 0 || lda1 % 4 == 0);
  if (mode != GIMS_GEMM_SIMT && tc_ok) return launch_gemm_tc(g, W.hi, W.lo, st);
  return launch_gemm(g, st);
}

// a-9 ------------------------------------------------------------------------------------------
static int sage_fwd(const gims_model* m, const float* feat, const int* indptr, const int* indices, int n_max,
                    const int* n_dev, float* out, float* scratch, int mode, cudaStream_t st) {
  if (!m || n_max < 1) { set_error("gims_sage_forward: bad arguments"); return GIMS_ERR_ARG; }
  const int H = kD / 2;
  float* y0 = scratch;                          // [n][256]  = feat @ [Wn0; Ws0]^T
  float* h1 = y0 + (size_t)n_max * kD;          // [n][128]
  float* agg = h1 + (size_t)n_max * H;          // [n][128]
  float* h2 = agg + (size_t)n_max * H;          // [n][128]   (total 2.5 * n * 256 floats)
  Segs s = one_seg(n_max, n_dev);
  // layer 0 (256 -> 128, fc_neigh before aggregation)
  GIMS_TRY(gemm(feat, kD, kD, nullptr, 0, 0, m->sage_w[0], nullptr, nullptr, 0, y0, kD, kD, 0, s, mode, st));
  GIMS_TRY(launch_sage_aggregate(y0, kD, H, indptr, indices, n_max, n_dev, y0 + H, kD, m->sage_b[0], 1, h1, H, st));
  // layer 1 (128 -> 128, aggregate then fc_neigh)
  GIMS_TRY(launch_sage_aggregate(h1, H, H, indptr, indices, n_max, n_dev, nullptr, 0, nullptr, 0, agg, H, st));
  GIMS_TRY(gemm(h1, H, H, agg, H, H, m->sage_w[1], m->sage_b[1], nullptr, 0, h2, H, H, 1, s, mode, st));
  // layer 2 (128 -> 256)
  GIMS_TRY(launch_sage_aggregate(h2, H, H, indptr, indices, n_max, n_dev, nullptr, 0, nullptr, 0, agg, H, st));
  GIMS_TRY(gemm(h2, H, H, agg, H, H, m->sage_w[2], m->sage_b[2], nullptr, 0, out, kD, kD, 0, s, mode, st));
  return GIMS_OK;
}
extern "C" int gims_sage_forward(const gims_model* m, const float* feat, const int* indptr, const int* indices,
                                 int n_max, const int* n_dev, float* out, float* scratch, void* stream) {
  return sage_fwd(m, feat, indptr, indices, n_max, n_dev, out, scratch, g_gemm_mode.load(), static_cast<cudaStream_t>(stream));
}

// a-8 + a-10 -----------------------------------------------------------------------------------
static int kenc_fwd(const gims_model* m, const float* kpts, int n_max, const int* n_dev, float img_w, float img_h,
                    const float* add, float* desc, float* scratch, int mode, cudaStream_t st) {
  if (!m || n_max < 1) { set_error("gims_kenc_forward: bad arguments"); return GIMS_ERR_ARG; }
  const gims_config& c = m->cfg;
  float* buf[2] = {scratch, scratch + (size_t)n_max * kD};
  Segs s = one_seg(n_max, n_dev);
  GIMS_TRY(launch_kenc_first(kpts, n_max, n_dev, img_w, img_h, m->kenc_w[0].w, m->kenc_b[0], c.kenc_dims[1], buf[0], st));
  int cur = 0;
  for (int i = 1; i < c.kenc_num; ++i) {
    int cin = c.kenc_dims[i], cout = c.kenc_dims[i + 1];
    bool last = (i == c.kenc_num - 1);
    float* y = last ? desc : buf[cur ^ 1];
    GIMS_TRY(gemm(buf[cur], cin, cin, nullptr, 0, 0, m->kenc_w[i], m->kenc_b[i], last ? add : nullptr, kD, y,
                              last ? kD : cout, cout, last ? 0 : 1, s, mode, st));
    cur ^= 1;
  }
  return GIMS_OK;
}
extern "C" int gims_kenc_forward(const gims_model* m, const float* kpts, int n_max, const int* n_dev, float img_w,
                                 float img_h, const float* add, float* desc, float* scratch, void* stream) {
  return kenc_fwd(m, kpts, n_max, n_dev, img_w, img_h, add, desc, scratch, g_gemm_mode.load(), static_cast<cudaStream_t>(stream));
}

// a-11 + a-12 ----------------------------------------------------------------------------------
// qkv (or its tf32 planes: 6 * rows * 256 + padding of the transposed V rows) | att | msg | hid
extern "C" size_t gims_attn_scratch_floats(int rows) { return (size_t)rows * (6 * kD + kD + kD + 2 * kD) + 2 * 128 * kD; }

static int attn_layer(const gims_model* m, int layer, float* desc, int n0_max, int n1_max, const int* n_dev,
                      float* scratch, unsigned* status_dev, int mode, cudaStream_t st) {
  if (!m || layer < 0 || layer >= m->cfg.num_layers) { set_error("gims_attn_layer_forward: bad layer %d", layer); return GIMS_ERR_ARG; }
  size_t rows = (size_t)n0_max + n1_max;
  int ldv = attn_ldv(n0_max, n1_max);      // <= rows + 126
  float* qkv = scratch;                    // [rows][768]  (SIMT)  |  Qp, Kp, Vt planes (tensor cores)
  float* att = qkv + rows * 6 * kD + 2 * 128 * kD;   // [rows][256]
  float* msg = att + rows * kD;            // [rows][256]
  float* hid = msg + rows * kD;            // [rows][512]
  Segs s = two_segs(n0_max, n1_max, n_dev);
  if (mode != GIMS_GEMM_SIMT) {
    QkvPlanes pl;
    pl.qp = qkv; pl.kp = pl.qp + 2 * rows * kD; pl.vt = static_cast<float*>(pl.kp) + 2 * rows * kD; pl.ldv = ldv;
    pl.vbase1 = attn_vbase1(n0_max);
    pl.fmt = mode == GIMS_GEMM_TC_F16 ? 0 : (mode == GIMS_GEMM_BF16 ? 1 : -1);
    pl.planes = mode == GIMS_GEMM_BF16 ? 1 : 2;
    pl.status = status_dev;
    GemmArgs g;
    g.A0 = desc; g.lda0 = kD; g.K0 = kD; g.A1 = nullptr; g.lda1 = 0; g.K1 = 0; g.W = m->wqkv[layer].w;
    g.bias = m->bqkv[layer]; g.R = nullptr; g.ldr = 0; g.Y = nullptr; g.ldy = 0; g.N = 3 * kD; g.relu = 0; g.segs = s;
    GIMS_TRY(launch_gemm_tc(g, m->wqkv[layer].hi, m->wqkv[layer].lo, st, &pl));
    if (pl.fmt >= 0) GIMS_TRY(launch_attention_f16(pl, att, n0_max, n1_max, n_dev, m->cfg.layer_is_cross[layer], st));
    else             GIMS_TRY(launch_attention_tc(pl, att, n0_max, n1_max, n_dev, m->cfg.layer_is_cross[layer], st));
  } else {
    GIMS_TRY(gemm(desc, kD, kD, nullptr, 0, 0, m->wqkv[layer], m->bqkv[layer], nullptr, 0, qkv, 3 * kD, 3 * kD, 0, s, mode, st));
    GIMS_TRY(launch_attention(qkv, att, n0_max, n1_max, n_dev, m->cfg.layer_is_cross[layer], st));
  }
  (void)msg;   // the merge conv is composed into W1 at pack time
  GIMS_TRY(gemm(desc, kD, kD, att, kD, kD, m->w1[layer], m->b1[layer], nullptr, 0, hid, 2 * kD, 2 * kD, 1, s, mode, st));
  GIMS_TRY(gemm(hid, 2 * kD, 2 * kD, nullptr, 0, 0, m->w2[layer], m->b2[layer], desc, kD, desc, kD, kD, 0, s, mode, st));
  return GIMS_OK;
}
extern "C" int gims_attn_layer_forward(const gims_model* m, int layer, float* desc, int n0_max, int n1_max,
                                       const int* n_dev, float* scratch, unsigned* status_dev, void* stream) {
  return attn_layer(m, layer, desc, n0_max, n1_max, n_dev, scratch, status_dev, g_gemm_mode.load(), static_cast<cudaStream_t>(stream));
}

// a-13 -----------------------------------------------------------------------------------------
static int final_scores(const gims_model* m, const float* desc, int n0_max, int n1_max, const int* n_dev, float* mdesc,
                        float* couplings, float* scratch, int mode, cudaStream_t st) {
  if (!m) { set_error("gims_final_scores: null model"); return GIMS_ERR_ARG; }
  Segs s = two_segs(n0_max, n1_max, n_dev);
  GIMS_TRY(gemm(desc, kD, kD, nullptr, 0, 0, m->wfinal, m->bfinal, nullptr, 0, mdesc, kD, kD, 0, s, mode, st));
  if (mode != GIMS_GEMM_SIMT && scratch) {
    GIMS_TRY(launch_score_gemm_tc(mdesc, n0_max, n1_max, n_dev, scratch, couplings, st));
    GIMS_TRY(launch_score_border(n0_max, n1_max, n_dev, m->bin_score, couplings, st));
  } else {
    GIMS_TRY(launch_score_gemm(mdesc, n0_max, n1_max, n_dev, m->bin_score, couplings, st));
  }
  return GIMS_OK;
}
extern "C" int gims_final_scores(const gims_model* m, const float* desc, int n0_max, int n1_max, const int* n_dev,
                                 float* mdesc, float* couplings, float* scratch, void* stream) {
  return final_scores(m, desc, n0_max, n1_max, n_dev, mdesc, couplings, scratch, g_gemm_mode.load(), static_cast<cudaStream_t>(stream));
}

// test / bring-up entry points ----------------------------------------------------------------
extern "C" int gims_linear(const float* A0, int lda0, int K0, const float* A1, int lda1, int K1, const float* W,
                           const float* W_hi, const float* W_lo, const float* bias, const float* R, int ldr, float* Y,
                           int ldy, int N, int relu, int rows_max, const int* rows_dev, int mode, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  GemmArgs g;
  g.A0 = A0; g.lda0 = lda0; g.K0 = K0; g.A1 = A1; g.lda1 = lda1; g.K1 = K1; g.W = W; g.bias = bias;
  g.R = R; g.ldr = ldr; g.Y = Y; g.ldy = ldy; g.N = N; g.relu = relu; g.segs = one_seg(rows_max, rows_dev);
  if (mode == GIMS_GEMM_TC) {
    if (!W_hi || !W_lo) { set_error("gims_linear: tensor-core mode needs W_hi / W_lo"); return GIMS_ERR_ARG; }
    return launch_gemm_tc(g, W_hi, W_lo, st);
  }
  return launch_gemm(g, st);
}

extern "C" int gims_debug_attention_trace(long long* dev_buf) { return set_attention_trace(dev_buf); }
extern "C" int gims_debug_gemm_trace(long long* dev_buf) { return set_gemm_trace(dev_buf); }

extern "C" int gims_split_tf32(const float* x, float* hi, float* lo, size_t n, void* stream) {
  if (n % 4) { set_error("gims_split_tf32: n must be a multiple of 4"); return GIMS_ERR_ARG; }
  return launch_split_planes(x, hi, lo, n, static_cast<cudaStream_t>(stream));
}

// whole pair -----------------------------------------------------------------------------------
namespace {
struct PairWs {
  void* agc;        size_t agc_bytes;
  float* sage_out;  // [rows][256]
  float* desc;      // [rows][256]
  float* scratch;   // attention scratch (also SAGE / kenc scratch)
  float* couplings;
  void* sink;       size_t sink_bytes;
};
size_t carve_pair(PairWs& w, void* base, size_t cap, int n0, int n1, int edge_cap) {
  Arena a(base, cap);
  size_t rows = (size_t)n0 + n1;
  int nmax = n0 > n1 ? n0 : n1;
  w.agc_bytes = gims_agc_workspace_bytes(nmax, edge_cap);
  w.agc = a.take<char>(w.agc_bytes);
  w.sage_out = a.take<float>(rows * kD);
  w.desc = a.take<float>(rows * kD);
  w.scratch = a.take<float>(gims_attn_scratch_floats((int)rows));
  w.couplings = a.take<float>((size_t)(n0 + 1) * coup_ld(n1));
  w.sink_bytes = gims_sinkhorn_workspace_bytes(n0, n1);
  w.sink = a.take<char>(w.sink_bytes);
  return align_up(a.off, 256);
}
__global__ void k_copy_rows(const float* __restrict__ src, float* __restrict__ dst, size_t n4) {
  size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n4) reinterpret_cast<float4*>(dst)[i] = reinterpret_cast<const float4*>(src)[i];
}
}  // namespace

extern "C" size_t gims_pair_workspace_bytes(const gims_model* m, int n0, int n1, int edge_cap) {
  (void)m;
  PairWs w;
  return carve_pair(w, nullptr, 0, n0, n1, edge_cap);
}

extern "C" int gims_forward_pair(const gims_model* m, const gims_pair_inputs* in, const gims_pair_outputs* o,
                                 void* workspace, size_t workspace_bytes, void* stream) {
  cudaStream_t st = static_cast<cudaStream_t>(stream);
  if (!m || !in || !o || !workspace) { set_error("gims_forward_pair: null argument"); return GIMS_ERR_ARG; }
  int n0 = in->n[0], n1 = in->n[1];
  // validate everything BEFORE the first launch: a late error would leave half a forward enqueued
  if (n0 < 2 || n1 < 2 || n0 > GIMS_MAX_KPTS || n1 > GIMS_MAX_KPTS) {
    set_error("gims_forward_pair: keypoint counts (%d, %d) outside [2, %d]", n0, n1, GIMS_MAX_KPTS);
    return GIMS_ERR_ARG;
  }
  if (n1 > gims_sinkhorn_max_columns()) {
    set_error("gims_forward_pair: n1=%d exceeds the Sinkhorn kernel's shared-memory limit of %d columns", n1, gims_sinkhorn_max_columns());
    return GIMS_ERR_ARG;
  }
  if (in->edge_cap < 1) { set_error("gims_forward_pair: edge_cap %d", in->edge_cap); return GIMS_ERR_ARG; }
  if (!o->n_kept_dev || !o->n_edges_dev || !o->n_comp_dev || !o->thr_dev || !o->mdesc || !o->u || !o->v || !o->status_dev) {
    set_error("gims_forward_pair: null output pointer");
    return GIMS_ERR_ARG;
  }
  for (int s = 0; s < 2; ++s) {
    if (!in->kpts[s] || !in->desc[s] || !in->scores[s] || !o->kept_idx[s] || !o->csr_indptr[s] || !o->csr_indices[s] ||
        !o->kpts[s] || !o->feat[s] || !o->scores[s] || !o->matches[s] || !o->mscores[s] || !o->indices[s]) {
      set_error("gims_forward_pair: null pointer for image %d", s);
      return GIMS_ERR_ARG;
    }
    long long len = (long long)in->n[s] * (in->n[s] - 1) / 2;
    if (in->k_rank[s] < 0 || in->k_rank[s] >= len) {
      set_error("gims_forward_pair: k_rank[%d]=%lld outside [0, %lld)", s, in->k_rank[s], len);
      return GIMS_ERR_ARG;
    }
  }
  PairWs w;
  size_t need = carve_pair(w, workspace, workspace_bytes, n0, n1, in->edge_cap);
  if (need > workspace_bytes) { set_error("gims_forward_pair: workspace %zu < %zu", workspace_bytes, need); return GIMS_ERR_WORKSPACE; }
  size_t rows = (size_t)n0 + n1;
  const int mode = in->gemm_mode > 0 ? in->gemm_mode - 1 : g_gemm_mode.load();
  if (mode < GIMS_GEMM_SIMT || mode > GIMS_GEMM_BF16) { set_error("gims_forward_pair: gemm_mode %d", in->gemm_mode); return GIMS_ERR_ARG; }
  GIMS_CUDA_OK(cudaMemsetAsync(o->status_dev, 0, sizeof(unsigned), st));
  // a-1 .. a-7: graphs + pruning, both images (gmatcher.py:233-252)
  for (int s = 0; s < 2; ++s) {
    GIMS_TRY(gims_agc_build(in->kpts[s], in->desc[s], in->desc_channel_major, in->scores[s], in->n[s], in->radius,
                            in->k_rank[s], in->min_size, w.agc, w.agc_bytes, o->kept_idx[s], o->n_kept_dev + s,
                            o->csr_indptr[s], o->csr_indices[s], in->edge_cap, o->n_edges_dev + s, o->kpts[s], o->feat[s],
                            o->scores[s], o->thr_dev + s, o->n_comp_dev + s, o->status_dev, stream));
  }
  // a-9, a-8, a-10: desc = SAGE(feat) + kenc(normalize(kpts))   (gmatcher.py:265-271)
  for (int s = 0; s < 2; ++s) {
    size_t base = s ? (size_t)n0 : 0;
    GIMS_TRY(sage_fwd(m, o->feat[s], o->csr_indptr[s], o->csr_indices[s], in->n[s], o->n_kept_dev + s,
                      w.sage_out + base * kD, w.scratch, mode, st));
    GIMS_TRY(kenc_fwd(m, o->kpts[s], in->n[s], o->n_kept_dev + s, in->img_w[s], in->img_h[s],
                      w.sage_out + base * kD, w.desc + base * kD, w.scratch, mode, st));
  }
  if (o->desc_in) {
    k_copy_rows<<<(unsigned)((rows * kD / 4 + 255) / 256), 256, 0, st>>>(w.desc, o->desc_in, rows * kD / 4);
    GIMS_LAUNCH_OK();
  }
  // a-11, a-12: attention stack (gmatcher.py:272)
  for (int l = 0; l < m->cfg.num_layers; ++l)
    GIMS_TRY(attn_layer(m, l, w.desc, n0, n1, o->n_kept_dev, w.scratch, o->status_dev, mode, st));
  if (o->desc_gnn) {
    k_copy_rows<<<(unsigned)((rows * kD / 4 + 255) / 256), 256, 0, st>>>(w.desc, o->desc_gnn, rows * kD / 4);
    GIMS_LAUNCH_OK();
  }
  // a-13 .. a-15
  float* coup = o->couplings ? o->couplings : w.couplings;
  GIMS_TRY(final_scores(m, w.desc, n0, n1, o->n_kept_dev, o->mdesc, coup, w.scratch, mode, st));
  GIMS_TRY(gims_sinkhorn_match(coup, coup_ld(n1), n0, n1, o->n_kept_dev, m->cfg.sinkhorn_iterations, m->cfg.match_threshold, w.sink,
                               w.sink_bytes, o->u, o->v, o->indices[0], o->indices[1], o->matches[0], o->matches[1],
                               o->mscores[0], o->mscores[1], o->status_dev, stream));
  return GIMS_OK;
}
