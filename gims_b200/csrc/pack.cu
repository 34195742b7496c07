// gims_pack_weights: a reference GMatcher.state_dict() -> the flat fp32 buffer gims_model_create consumes (SURVEY.md §8 a-0,
// §8b).  Host code only; the same transformations as gims_b200/packing.py (which stays the Python host's packer and is the
// specification: tests/test_abi_cpu.py compares the two), for integrators that do not run Python:
//   * eval-mode BatchNorm1d folded into the preceding Conv1d(k=1)          (models/gmatcher.py:11-24)
//   * attention heads de-interleaved: reference channel c = d*4 + h  ->  packed channel c' = h*64 + d   (gmatcher.py:108-111)
//   * proj[0] | proj[1] | proj[2] stacked into one [768][256] matrix; `merge` composed into the first MLP conv (fp64)
//   * SAGEConv: layer 0 rows [fc_neigh; fc_self], layers 1, 2 columns [fc_self | fc_neigh]
//   * every weight matrix followed by its tensor-core planes: tf32 hi / lo, fp16 hi / lo of W * 2^e, and 2^-e
// All arithmetic in double, rounded to fp32 once, exactly like the Python packer.
#include <cmath>
#include <cstring>
#include <string>
#include <vector>

#include <cuda_fp16.h>

#include "common.cuh"

namespace gims {
namespace {

constexpr double kBnEps = 1e-5;          // torch.nn.BatchNorm1d default
constexpr int kSageLayers = 3;

struct Blob { std::vector<float> data; };

struct Packer {
  const gims_named_tensor* t;
  int n;
  std::string missing;

  const float* find(const std::string& name, int64_t numel) {
    for (int i = 0; i < n; ++i)
      if (t[i].name && name == t[i].name) {
        if (t[i].numel != numel || !t[i].data) {
          if (missing.empty()) missing = name + " (wrong element count)";
          return nullptr;
        }
        return t[i].data;
      }
    if (missing.empty()) missing = name;
    return nullptr;
  }
  bool has(const std::string& name) const {
    for (int i = 0; i < n; ++i)
      if (t[i].name && name == t[i].name) return true;
    return false;
  }
  std::vector<double> get(const std::string& name, int64_t numel) {
    std::vector<double> v((size_t)numel, 0.0);
    const float* p = find(name, numel);
    if (p)
      for (int64_t i = 0; i < numel; ++i) v[(size_t)i] = (double)p[i];
    return v;
  }
};

// W [rows][cols], b [rows]  <-  BatchNorm(prefix) o (W, b)
void fold_bn(Packer& pk, const std::string& bn, std::vector<double>& w, std::vector<double>& b, int rows, int cols) {
  std::vector<double> gamma = pk.get(bn + ".weight", rows), beta = pk.get(bn + ".bias", rows),
                      mean = pk.get(bn + ".running_mean", rows), var = pk.get(bn + ".running_var", rows);
  for (int r = 0; r < rows; ++r) {
    const double s = gamma[r] / std::sqrt(var[r] + kBnEps);
    for (int c = 0; c < cols; ++c) w[(size_t)r * cols + c] *= s;
    b[r] = (b[r] - mean[r]) * s + beta[r];
  }
}

struct Out {
  std::vector<Blob> blobs;
  void add_vec(const std::vector<double>& v) {
    Blob b;
    b.data.resize(v.size());
    for (size_t i = 0; i < v.size(); ++i) b.data[i] = (float)v[i];
    blobs.push_back(std::move(b));
  }
  // weight matrix: W, tf32 hi / lo, fp16 hi / lo (raw bits, two per float word), 2^-e
  void add_mat(const std::vector<double>& v) {
    add_vec(v);
    const std::vector<float> w = blobs.back().data;      // fp32-rounded
    const size_t n = w.size();
    Blob hi, lo, h16, l16, sinv;
    hi.data.resize(n); lo.data.resize(n);
    for (size_t i = 0; i < n; ++i) {
      uint32_t bits;
      std::memcpy(&bits, &w[i], 4);
      bits = (bits + 0x1000u) & ~0x1FFFu;                 // 10-bit mantissa, ties away from zero (cvt.rna.tf32.f32)
      float h;
      std::memcpy(&h, &bits, 4);
      hi.data[i] = h;
      lo.data[i] = w[i] - h;
    }
    float amax = 0.f;
    for (size_t i = 0; i < n; ++i) amax = std::fmax(amax, std::fabs(w[i]));
    int e = 0;
    if (amax > 0.f) {
      const float q = (float)(1024.0 / (double)amax);
      e = (int)std::floor(std::log2(q));
    }
    e = e < -12 ? -12 : (e > 24 ? 24 : e);
    const float scale = (float)std::ldexp(1.0, e);
    h16.data.assign((n + 1) / 2, 0.f); l16.data.assign((n + 1) / 2, 0.f);
    uint16_t* hp = reinterpret_cast<uint16_t*>(h16.data.data());
    uint16_t* lp = reinterpret_cast<uint16_t*>(l16.data.data());
    for (size_t i = 0; i < n; ++i) {
      const float s = w[i] * scale;
      const __half h = __float2half_rn(s);
      const __half l = __float2half_rn(s - __half2float(h));
      std::memcpy(&hp[i], &h, 2);
      std::memcpy(&lp[i], &l, 2);
    }
    sinv.data.assign(1, (float)std::ldexp(1.0, -e));
    blobs.push_back(std::move(hi)); blobs.push_back(std::move(lo));
    blobs.push_back(std::move(h16)); blobs.push_back(std::move(l16)); blobs.push_back(std::move(sinv));
  }
};

int check_cfg(const gims_config* c) {
  if (!c) { set_error("gims_pack_weights: null config"); return GIMS_ERR_ARG; }
  if (c->descriptor_dim != kD || c->num_layers < 0 || c->num_layers > GIMS_MAX_LAYERS || c->kenc_num < 2 ||
      c->kenc_num > GIMS_MAX_KENC || c->kenc_dims[0] != 2 || c->kenc_dims[c->kenc_num] != kD) {
    set_error("gims_pack_weights: unsupported config");
    return GIMS_ERR_ARG;
  }
  return GIMS_OK;
}

// sizes of the blobs in packing order (floats, before the 64-float alignment of every blob)
void blob_sizes(const gims_config* c, std::vector<size_t>& sz) {
  auto mat = [&](size_t rows, size_t cols) {
    const size_t n = rows * cols;
    sz.push_back(n); sz.push_back(n); sz.push_back(n); sz.push_back((n + 1) / 2); sz.push_back((n + 1) / 2); sz.push_back(1);
  };
  const size_t d = kD, h = kD / 2;
  sz.push_back(1);
  for (int i = 0; i < c->kenc_num; ++i) { mat(c->kenc_dims[i + 1], c->kenc_dims[i]); sz.push_back(c->kenc_dims[i + 1]); }
  mat(2 * h, d); sz.push_back(h);          // SAGE 0: [fc_neigh; fc_self], 256 -> 128
  mat(h, 2 * h); sz.push_back(h);          // SAGE 1: [fc_self | fc_neigh], 128 -> 128
  mat(d, 2 * h); sz.push_back(d);          // SAGE 2: 128 -> 256
  for (int l = 0; l < c->num_layers; ++l) {
    mat(3 * d, d); sz.push_back(3 * d);
    mat(2 * d, 2 * d); sz.push_back(2 * d);
    mat(d, 2 * d); sz.push_back(d);
  }
  mat(d, d); sz.push_back(d);
}
size_t padded(size_t n) { return (n + 63) / 64 * 64; }     // every blob 256-byte aligned

}  // namespace
}  // namespace gims

using namespace gims;

extern "C" size_t gims_pack_weights_floats(const gims_config* cfg) {
  if (check_cfg(cfg) != GIMS_OK) return 0;
  std::vector<size_t> sz;
  blob_sizes(cfg, sz);
  size_t total = 0;
  for (size_t s : sz) total += padded(s);
  return total;
}

extern "C" int gims_pack_weights(const gims_config* cfg, const gims_named_tensor* tensors, int n_tensors, float* packed_host,
                                 size_t capacity_floats, int64_t* offsets_host, int n_offsets) {
  GIMS_TRY(check_cfg(cfg));
  if (!tensors || n_tensors < 1 || !packed_host || !offsets_host) { set_error("gims_pack_weights: null argument"); return GIMS_ERR_ARG; }
  if (n_offsets != gims_packed_blob_count(cfg)) {
    set_error("gims_pack_weights: %d offsets, the config has %d blobs", n_offsets, gims_packed_blob_count(cfg));
    return GIMS_ERR_ARG;
  }
  if (capacity_floats < gims_pack_weights_floats(cfg)) {
    set_error("gims_pack_weights: buffer of %zu floats, need %zu", capacity_floats, gims_pack_weights_floats(cfg));
    return GIMS_ERR_WORKSPACE;
  }
  Packer pk{tensors, n_tensors, {}};
  Out out;
  const int d = kD, hd = kD / GIMS_NUM_HEADS, H = kD / 2;
  char key[160];
  auto K = [&](const char* fmt, int a, int b = 0) { std::snprintf(key, sizeof key, fmt, a, b); return std::string(key); };

  out.add_vec(pk.get("bin_score", 1));
  // keypoint encoder (gmatcher.py:87-97): Conv1d - BN - ReLU ... Conv1d; modules 3 i (conv), 3 i + 1 (bn)
  for (int i = 0; i < cfg->kenc_num; ++i) {
    const int cin = cfg->kenc_dims[i], cout = cfg->kenc_dims[i + 1];
    std::vector<double> w = pk.get(K("kenc.encoder.%d.weight", 3 * i), (int64_t)cout * cin);
    std::vector<double> b = pk.get(K("kenc.encoder.%d.bias", 3 * i), cout);
    if (i < cfg->kenc_num - 1) fold_bn(pk, K("kenc.encoder.%d", 3 * i + 1), w, b, cout, cin);
    out.add_mat(w);
    out.add_vec(b);
  }
  // GraphSAGE (gmatcher.py:192-197)
  const int sin[kSageLayers] = {d, H, H}, sout[kSageLayers] = {H, H, d};
  for (int l = 0; l < kSageLayers; ++l) {
    const int ci = sin[l], co = sout[l];
    std::vector<double> wn = pk.get(K("gnn_encoder.layers.%d.fc_neigh.weight", l), (int64_t)co * ci);
    std::vector<double> ws = pk.get(K("gnn_encoder.layers.%d.fc_self.weight", l), (int64_t)co * ci);
    const std::string bkey = pk.has(K("gnn_encoder.layers.%d.bias", l)) ? K("gnn_encoder.layers.%d.bias", l)
                                                                         : K("gnn_encoder.layers.%d.fc_self.bias", l);
    std::vector<double> w((size_t)2 * co * ci);
    if (ci > co) {                                   // rows: [fc_neigh; fc_self]
      std::copy(wn.begin(), wn.end(), w.begin());
      std::copy(ws.begin(), ws.end(), w.begin() + (size_t)co * ci);
    } else {                                         // columns: [fc_self | fc_neigh]
      for (int r = 0; r < co; ++r)
        for (int c = 0; c < ci; ++c) {
          w[(size_t)r * 2 * ci + c] = ws[(size_t)r * ci + c];
          w[(size_t)r * 2 * ci + ci + c] = wn[(size_t)r * ci + c];
        }
    }
    out.add_mat(w);
    out.add_vec(pk.get(bkey, co));
  }
  // attention layers (gmatcher.py:99-143)
  std::vector<int> perm(d);                          // perm[c'] = reference channel of packed channel c' = h*64 + dd
  for (int cp = 0; cp < d; ++cp) perm[cp] = (cp % hd) * GIMS_NUM_HEADS + cp / hd;
  for (int l = 0; l < cfg->num_layers; ++l) {
    std::vector<double> wqkv((size_t)3 * d * d), bqkv((size_t)3 * d);
    for (int j = 0; j < 3; ++j) {
      std::vector<double> w = pk.get(K("gnn.layers.%d.attn.proj.%d.weight", l, j), (int64_t)d * d);
      std::vector<double> b = pk.get(K("gnn.layers.%d.attn.proj.%d.bias", l, j), d);
      for (int r = 0; r < d; ++r) {
        std::copy(w.begin() + (size_t)perm[r] * d, w.begin() + (size_t)(perm[r] + 1) * d, wqkv.begin() + ((size_t)j * d + r) * d);
        bqkv[(size_t)j * d + r] = b[perm[r]];
      }
    }
    out.add_mat(wqkv);
    out.add_vec(bqkv);
    // `merge` (gmatcher.py:114) is linear and feeds only the first MLP conv (gmatcher.py:125):
    //   W1 [x | merge(att)] + b1 = [W1x | W1m Wm] [x | att] + (b1 + W1m bm),  Wm's input columns in packed head order
    std::vector<double> wm_raw = pk.get(K("gnn.layers.%d.attn.merge.weight", l), (int64_t)d * d);
    std::vector<double> bm = pk.get(K("gnn.layers.%d.attn.merge.bias", l), d);
    std::vector<double> wm((size_t)d * d);
    for (int r = 0; r < d; ++r)
      for (int c = 0; c < d; ++c) wm[(size_t)r * d + c] = wm_raw[(size_t)r * d + perm[c]];
    std::vector<double> w1_raw = pk.get(K("gnn.layers.%d.mlp.0.weight", l), (int64_t)2 * d * 2 * d);
    std::vector<double> b1 = pk.get(K("gnn.layers.%d.mlp.0.bias", l), 2 * d);
    std::vector<double> w1((size_t)2 * d * 2 * d, 0.0);
    for (int r = 0; r < 2 * d; ++r) {
      const double* src = &w1_raw[(size_t)r * 2 * d];
      double* dst = &w1[(size_t)r * 2 * d];
      std::copy(src, src + d, dst);
      double acc_b = 0.0;
      for (int k = 0; k < d; ++k) {                  // dst[d + c] = sum_k W1m[r][k] Wm[k][c]
        const double a = src[d + k];
        const double* wk = &wm[(size_t)k * d];
        for (int c = 0; c < d; ++c) dst[d + c] += a * wk[c];
        acc_b += a * bm[k];
      }
      b1[r] += acc_b;
    }
    fold_bn(pk, K("gnn.layers.%d.mlp.1", l), w1, b1, 2 * d, 2 * d);
    out.add_mat(w1);
    out.add_vec(b1);
    out.add_mat(pk.get(K("gnn.layers.%d.mlp.3.weight", l), (int64_t)d * 2 * d));
    out.add_vec(pk.get(K("gnn.layers.%d.mlp.3.bias", l), d));
  }
  out.add_mat(pk.get("final_proj.weight", (int64_t)d * d));
  out.add_vec(pk.get("final_proj.bias", d));
  if (!pk.missing.empty()) { set_error("gims_pack_weights: state_dict entry missing: %s", pk.missing.c_str()); return GIMS_ERR_ARG; }
  if ((int)out.blobs.size() != n_offsets) { set_error("gims_pack_weights: internal blob count %zu != %d", out.blobs.size(), n_offsets); return GIMS_ERR_ARG; }
  size_t off = 0;
  for (size_t i = 0; i < out.blobs.size(); ++i) {
    const std::vector<float>& v = out.blobs[i].data;
    offsets_host[i] = (int64_t)off;
    std::memcpy(packed_host + off, v.data(), v.size() * sizeof(float));
    const size_t pad = padded(v.size()) - v.size();
    if (pad) std::memset(packed_host + off + v.size(), 0, pad * sizeof(float));
    off += v.size() + pad;
  }
  return GIMS_OK;
}
