// unet_exec.cu -- the network boundary of SURVEY 8b as C entry points: _LIB_unet_forward_2d / _LIB_unet_forward_3d.
//
// Reference boundary: `keras_model.predict(x[np.newaxis])` (stardist/models/base.py:408-410) for the graph built in
// stardist/models/model2d.py:310-349 / model3d.py:360-399 (csbdeep unet_block: conv+bias+ReLU x n per level, max-pool 2,
// nearest up-sampling 2, Concatenate([up, skip]), `features` conv, 1x1 heads prob (sigmoid) / dist (linear)).
// A caller without Python (the reference's Fiji / Java consumer binds stardist3d_lib.h the same way) creates a network
// object from the Keras kernels (layout (k..., Cin, Cout), names in topology order from sdb_unet_layer_name) and runs it on
// device buffers.  The executor walks the same layer list as stardist_b200/models/unet_device.py and launches the same
// kernels (sdb_stem_split / sdb_conv3x3_tc / sdb_maxpool_split / sdb_conv3x3_heads_tc in 2-D; sdb_conv3_nd + sdb_split_f32 /
// sdb_conv3x3x3_tc / sdb_maxpool3d_split / sdb_heads_tc in 3-D), so its maps are bit-equal to the Python path's
// (tests/test_gpu_unet_tc.py::test_c_unet_forward_equals_python_executor).
// Scope: the default architecture family -- U-Net backbone, 3^d kernels, pool 2, ReLU, no batch norm, grid 1, 2 <= R;
// 2-D needs n_rays <= 32 and net_conv_after_unet == 128 (fused features + heads kernel), 3-D n_rays <= 143.
#include <cuda_fp16.h>
#include <math.h>
#include <map>
#include <string>
#include <vector>
#include "common.cuh"
#include "../../include/stardist_b200.h"

namespace {

struct Layer { std::string name; int kind; int cin = 0, cout = 0, cin_lo = 0, skip = -1; bool relu = true; };   // kind 0 conv, 1 pool, 2 up
struct Wt { float* k = nullptr; float* b = nullptr; __half* hi = nullptr; __half* lo = nullptr; float scale = 1.f; int cin = 0, cout = 0; };

static std::vector<Layer> topology(const sdb_unet_config& c) {
  std::vector<Layer> L;
  const int depth = c.unet_n_depth, base = c.unet_n_filter_base, nconv = c.unet_n_conv_per_depth;
  int ch = c.n_channel_in;
  std::vector<int> skips;
  auto conv = [&](const std::string& name, int cin, int cout, int cin_lo) { Layer l; l.name = name; l.kind = 0; l.cin = cin; l.cout = cout; l.cin_lo = cin_lo; L.push_back(l); };
  for (int n = 0; n < depth; ++n) {
    for (int i = 0; i < nconv; ++i) { conv("down_level_" + std::to_string(n) + "_no_" + std::to_string(i), ch, base << n, 0); ch = base << n; }
    Layer p; p.name = "max_" + std::to_string(n); p.kind = 1; p.skip = n; L.push_back(p);
    skips.push_back(ch);
  }
  for (int i = 0; i < nconv - 1; ++i) { conv("middle_" + std::to_string(i), ch, base << depth, 0); ch = base << depth; }
  conv("middle_" + std::to_string(nconv), ch, base << std::max(0, depth - 1), 0); ch = base << std::max(0, depth - 1);
  for (int n = depth - 1; n >= 0; --n) {
    Layer u; u.name = "up_sampling_" + std::to_string(n); u.kind = 2; u.skip = n; L.push_back(u);
    int cin = ch + skips[n];
    for (int i = 0; i < nconv - 1; ++i) { conv("up_level_" + std::to_string(n) + "_no_" + std::to_string(i), cin, base << n, i == 0 ? ch : 0); cin = base << n; }
    conv("up_level_" + std::to_string(n) + "_no_" + std::to_string(nconv), cin, base << std::max(0, n - 1), nconv == 1 ? ch : 0);
    ch = base << std::max(0, n - 1);
  }
  conv("features", ch, c.net_conv_after_unet, 0);
  return L;
}

static float weight_scale(const float* w, size_t n) {      // tc_weight_scale of models/unet_device.py
  float m = 0.f;
  for (size_t i = 0; i < n; ++i) m = fmaxf(m, fabsf(w[i]));
  if (!isfinite(m) || !(m > 0.f)) return 1.f;
  return ldexpf(1.f, 10 - (int)floor(log2((double)m)));      // float64 log2 like numpy's
}

}  // namespace

struct sdb_unet {
  sdb_unet_config cfg;
  std::vector<Layer> layers;
  std::vector<std::string> names;      // weight-carrying layers in topology order + "prob", "dist"
  std::map<std::string, Wt> w;
  float* fuse_w = nullptr; float* fuse_b = nullptr;                          // 2-D: [cf][36], [36]
  __half* heads_hi = nullptr; __half* heads_lo = nullptr; float* heads_b = nullptr; float heads_scale = 1.f; int heads_np = 0;   // 3-D
  std::vector<void*> owned;
};

static std::vector<std::string> names_of(const sdb_unet_config& c) {
  std::vector<std::string> n;
  for (auto& l : topology(c)) if (l.kind == 0) n.push_back(l.name);
  n.push_back("prob"); n.push_back("dist");
  return n;
}

extern "C" int sdb_unet_layer_count(const sdb_unet_config* cfg) { return cfg ? (int)names_of(*cfg).size() : 0; }
extern "C" const char* sdb_unet_layer_name(const sdb_unet_config* cfg, int i) {
  static thread_local std::string s;
  if (!cfg) return nullptr;
  auto n = names_of(*cfg);
  if (i < 0 || i >= (int)n.size()) return nullptr;
  s = n[i];
  return s.c_str();
}

extern "C" void sdb_unet_destroy(sdb_unet* net) {
  if (!net) return;
  for (void* p : net->owned) cudaFree(p);
  delete net;
}

extern "C" sdb_unet* sdb_unet_create(const sdb_unet_config* cfg, const float* const* kernels, const float* const* biases) {
  if (!cfg || !kernels || !biases) { sdb::set_error("unet_create: null argument"); return nullptr; }
  const sdb_unet_config& c = *cfg;
  const int nd = c.ndim;
  if ((nd != 2 && nd != 3) || c.n_channel_in < 1 || c.n_channel_in > 4 || c.unet_n_depth < 1 || c.unet_n_conv_per_depth < 1 ||
      c.unet_n_filter_base % 32 != 0 || c.net_conv_after_unet % 64 != 0 || c.net_conv_after_unet <= 0 || c.n_rays < 2) {
    sdb::set_error("unet_create: unsupported configuration (U-Net, filters % 32 == 0, net_conv_after_unet % 64 == 0, <= 4 input channels)"); return nullptr;
  }
  for (int a = 0; a < nd; ++a) if (c.grid[a] != 1) { sdb::set_error("unet_create: grid must be 1 on every axis"); return nullptr; }
  if (nd == 2 && (c.n_rays > 32 || c.net_conv_after_unet != 128)) { sdb::set_error("unet_create: 2-D needs n_rays <= 32 and net_conv_after_unet == 128"); return nullptr; }
  if (nd == 3 && c.n_rays + 1 > 144) { sdb::set_error("unet_create: 3-D needs n_rays <= 143"); return nullptr; }
  sdb_unet* net = new sdb_unet();
  net->cfg = c; net->layers = topology(c); net->names = names_of(c);
  const int taps = nd == 2 ? 9 : 27;
  cudaStream_t st = 0;
  auto fail = [&](const char* msg) -> sdb_unet* { sdb::set_error(msg); sdb_unet_destroy(net); return nullptr; };
  auto dev_alloc = [&](size_t bytes) -> void* { void* p = nullptr; if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) return nullptr; net->owned.push_back(p); return p; };
  int wi = 0;
  for (auto& l : net->layers) {
    if (l.kind != 0) continue;
    const size_t nk = (size_t)taps * l.cin * l.cout;
    Wt W; W.cin = l.cin; W.cout = l.cout;
    W.k = (float*)dev_alloc(nk * 4); W.b = (float*)dev_alloc((size_t)l.cout * 4);
    if (!W.k || !W.b) return fail("unet_create: device allocation failed");
    if (cudaMemcpy(W.k, kernels[wi], nk * 4, cudaMemcpyHostToDevice) != cudaSuccess || cudaMemcpy(W.b, biases[wi], (size_t)l.cout * 4, cudaMemcpyHostToDevice) != cudaSuccess)
      return fail("unet_create: weight upload failed");
    if (l.cin % 32 == 0) {
      W.scale = weight_scale(kernels[wi], nk);
      W.hi = (__half*)dev_alloc(nk * 2); W.lo = (__half*)dev_alloc(nk * 2);
      if (!W.hi || !W.lo) return fail("unet_create: device allocation failed");
      const int rc = nd == 2 ? sdb_split_weights(W.k, l.cin, l.cout, W.scale, W.hi, W.lo, (sdb_stream_t)st)
                             : sdb_split_weights_3d(W.k, l.cin, l.cout, W.scale, W.hi, W.lo, (sdb_stream_t)st);
      if (rc) { sdb_unet_destroy(net); return nullptr; }
    }
    net->w[l.name] = W;
    ++wi;
  }
  const float* kp = kernels[wi]; const float* bp = biases[wi];          // prob: (1.., cf, 1)
  const float* kd = kernels[wi + 1]; const float* bd = biases[wi + 1];  // dist: (1.., cf, R)
  const int cf = c.net_conv_after_unet, R = c.n_rays;
  if (nd == 2) {
    std::vector<float> Wh((size_t)cf * 36, 0.f), bh(36, 0.f);
    for (int ch = 0; ch < cf; ++ch) { for (int r = 0; r < R; ++r) Wh[(size_t)ch * 36 + r] = kd[(size_t)ch * R + r]; Wh[(size_t)ch * 36 + 32] = kp[ch]; }
    for (int r = 0; r < R; ++r) bh[r] = bd[r];
    bh[32] = bp[0];
    net->fuse_w = (float*)dev_alloc(Wh.size() * 4); net->fuse_b = (float*)dev_alloc(36 * 4);
    if (!net->fuse_w || !net->fuse_b) return fail("unet_create: device allocation failed");
    cudaMemcpy(net->fuse_w, Wh.data(), Wh.size() * 4, cudaMemcpyHostToDevice); cudaMemcpy(net->fuse_b, bh.data(), 36 * 4, cudaMemcpyHostToDevice);
  } else {
    const int cands[4] = {48, 80, 112, 144};
    int np = 144; for (int v : cands) if (v >= R + 1) { np = v; break; }
    net->heads_np = np;
    std::vector<float> Wf((size_t)np * cf, 0.f), hb(np, 0.f);
    for (int ch = 0; ch < cf; ++ch) { Wf[ch] = kp[ch]; for (int r = 0; r < R; ++r) Wf[(size_t)(1 + r) * cf + ch] = kd[(size_t)ch * R + r]; }
    const float sc = weight_scale(Wf.data(), Wf.size());
    std::vector<__half> hi(Wf.size()), lo(Wf.size());
    for (size_t i = 0; i < Wf.size(); ++i) { const float v = Wf[i] * sc; const __half h = __float2half_rn(v); hi[i] = h; lo[i] = __float2half_rn(v - __half2float(h)); }
    hb[0] = bp[0]; for (int r = 0; r < R; ++r) hb[1 + r] = bd[r];
    net->heads_scale = sc;
    net->heads_hi = (__half*)dev_alloc(hi.size() * 2); net->heads_lo = (__half*)dev_alloc(lo.size() * 2); net->heads_b = (float*)dev_alloc((size_t)np * 4);
    if (!net->heads_hi || !net->heads_lo || !net->heads_b) return fail("unet_create: device allocation failed");
    cudaMemcpy(net->heads_hi, hi.data(), hi.size() * 2, cudaMemcpyHostToDevice); cudaMemcpy(net->heads_lo, lo.data(), lo.size() * 2, cudaMemcpyHostToDevice);
    cudaMemcpy(net->heads_b, hb.data(), (size_t)np * 4, cudaMemcpyHostToDevice);
  }
  if (cudaDeviceSynchronize() != cudaSuccess) return fail("unet_create: CUDA error while preparing the weights");
  return net;
}

namespace {
struct Act { __half* hi = nullptr; __half* lo = nullptr; int d = 1, h = 0, w = 0, c = 0; };

static int forward(sdb_unet* net, const float* d_x, int D, int H, int W, float* d_prob, float* d_dist, cudaStream_t st) {
  if (!net) { sdb::set_error("unet_forward: null network"); return 1; }
  const sdb_unet_config& c = net->cfg;
  const int nd = c.ndim, div = 1 << c.unet_n_depth;
  if (H % div || W % div || (nd == 3 && D % div)) { sdb::set_error("unet_forward: extents must be multiples of 2^depth (pad as StarDistPadAndCropResizer does)"); return 1; }
  std::vector<sdb::DevBuf*> bufs;
  struct Guard { std::vector<sdb::DevBuf*>& b; ~Guard() { for (auto* p : b) delete p; } } guard{bufs};
  auto new_act = [&](int d, int h, int w, int ch, Act* a) -> int {
    const size_t n = (size_t)d * h * w * ch;
    sdb::DevBuf* b = new sdb::DevBuf(); bufs.push_back(b);
    SDB_CUDA(b->alloc(n * 2 * sizeof(__half), st));
    a->hi = b->as<__half>(); a->lo = a->hi + n; a->d = d; a->h = h; a->w = w; a->c = ch;
    return 0;
  };
  Act cur, lo; bool have_lo = false, first = true;
  std::map<int, Act> skips;
  const auto& L = net->layers;
  const sdb_stream_t s = (sdb_stream_t)st;
  for (size_t i = 0; i < L.size(); ++i) {
    const Layer& l = L[i];
    if (l.kind == 0) {
      const Wt& Wl = net->w[l.name];
      const bool up = i + 1 < L.size() && L[i + 1].kind == 2;
      if (first) {
        Act out;
        if (new_act(D, H, W, l.cout, &out)) return 1;
        if (nd == 2) {
          if (sdb_stem_split(d_x, 1, H, W, c.n_channel_in, Wl.k, Wl.b, l.cout, 1, out.hi, out.lo, s)) return 1;
        } else {
          sdb::DevBuf* y = new sdb::DevBuf(); bufs.push_back(y);
          const size_t n = (size_t)D * H * W * l.cout;
          SDB_CUDA(y->alloc(n * 4, st));
          if (sdb_conv3_nd(d_x, nullptr, 1, D, H, W, c.n_channel_in, 0, 1, 1, 1, Wl.k, Wl.b, l.cout, 3, 1, y->as<float>(), s)) return 1;
          if (sdb_split_f32(y->as<float>(), (long long)n, out.hi, out.lo, s)) return 1;
        }
        cur = out; first = false;
        continue;
      }
      const int c0 = have_lo ? lo.c : 0;
      if (l.name == "features" && nd == 2) {
        if (sdb_conv3x3_heads_tc(have_lo ? lo.hi : nullptr, have_lo ? lo.lo : nullptr, c0, cur.hi, cur.lo, cur.c, 1, cur.h, cur.w, Wl.hi, Wl.lo, Wl.scale, Wl.b, 1,
                                 net->fuse_w, net->fuse_b, c.n_rays, d_prob, d_dist, s)) return 1;
        return sdb_tc_error_check(s);
      }
      Act out;
      const int f = up ? 2 : 1;
      if (new_act(nd == 3 ? cur.d * f : 1, cur.h * f, cur.w * f, l.cout, &out)) return 1;
      if (nd == 2) {
        if (sdb_conv3x3_tc(have_lo ? lo.hi : nullptr, have_lo ? lo.lo : nullptr, c0, cur.hi, cur.lo, cur.c, 1, cur.h, cur.w, Wl.hi, Wl.lo, Wl.scale, Wl.b,
                           l.cout, 1, up ? 1 : 0, out.hi, out.lo, s)) return 1;
      } else {
        if (sdb_conv3x3x3_tc(have_lo ? lo.hi : nullptr, have_lo ? lo.lo : nullptr, c0, cur.hi, cur.lo, cur.c, cur.d, cur.h, cur.w, Wl.hi, Wl.lo, Wl.scale, Wl.b,
                             l.cout, 1, up ? 2 : 0, out.hi, out.lo, s)) return 1;
      }
      have_lo = false; cur = out;
    } else if (l.kind == 1) {
      skips[l.skip] = cur;
      Act out;
      if (new_act(nd == 3 ? cur.d / 2 : 1, cur.h / 2, cur.w / 2, cur.c, &out)) return 1;
      if (nd == 2) { if (sdb_maxpool_split(cur.hi, cur.lo, 1, cur.h, cur.w, cur.c, out.hi, out.lo, s)) return 1; }
      else { if (sdb_maxpool3d_split(cur.hi, cur.lo, cur.d, cur.h, cur.w, cur.c, 2, 2, 2, out.hi, out.lo, s)) return 1; }
      cur = out;
    } else {
      lo = cur; have_lo = true;            // written at 2x resolution by its producer
      cur = skips[l.skip];
    }
  }
  // 3-D: 1x1x1 heads over the volume viewed as a [D*H, W] image
  if (sdb_heads_tc(cur.hi, cur.lo, cur.c, 1, cur.d * cur.h, cur.w, net->heads_hi, net->heads_lo, net->heads_scale, net->heads_b, net->heads_np, c.n_rays, d_prob, d_dist, s)) return 1;
  return sdb_tc_error_check(s);
}
}  // namespace

extern "C" int _LIB_unet_forward_2d(sdb_unet* net, const float* d_x, int h, int w, float* d_prob, float* d_dist, sdb_stream_t stream) {
  if (net && net->cfg.ndim != 2) { sdb::set_error("unet_forward_2d: network is not 2-D"); return 1; }
  return forward(net, d_x, 1, h, w, d_prob, d_dist, (cudaStream_t)stream);
}
extern "C" int _LIB_unet_forward_3d(sdb_unet* net, const float* d_x, int d, int h, int w, float* d_prob, float* d_dist, sdb_stream_t stream) {
  if (net && net->cfg.ndim != 3) { sdb::set_error("unet_forward_3d: network is not 3-D"); return 1; }
  return forward(net, d_x, d, h, w, d_prob, d_dist, (cudaStream_t)stream);
}
