// Handle-level VAE: aether_vae_encode / aether_vae_decode run the WHOLE schedule of AutoencoderKLCogVideoX.encode /
// .decode for one batch item -- 3 x 3 spatial tiling with the linear tile blends, frame batching with the causal
// conv caches, every resnet / norm / resampling stage -- as ONE C call that only enqueues kernels on the caller's
// stream: no host synchronisation, no allocation (scratch comes from the caller's workspace, sized by
// aether_vae_workspace_bytes), no Python between the ~7 700 launches of a 41 x 480 x 720 decode.
//
// Replaces the device work behind  vae.encode(x).latent_dist  and  vae.decode(z).sample  (third-party diffusers
// module; reference call sites aether/pipelines/aetherv1_pipeline_cogvideox.py:557-620, :931, :936; SURVEY.md 8(b)
// minimum export set `vae_encode_tile` / `vae_decode_tile`).  Control flow restated from the published diffusers
// v0.32 algorithm exactly like aether_b200/vae.py, which keeps the per-op path for the parity tests.
//
// Memory: every activation is a channels-last bf16 tensor taken from a first-fit free-list arena over the workspace.
// The same schedule code runs in DRY mode (no launches, arena without backing memory) to obtain the peak size, so the
// size query and the real run can never disagree.  Everything is stream-ordered on one stream, so a block freed by
// the host-side bookkeeping may be reused by a later launch immediately.
#include <algorithm>
#include <exception>
#include <map>
#include <new>
#include <string>
#include <unordered_map>
#include <vector>

#include "host_util.h"
#include "vae_internal.h"

namespace {
using namespace aether;

struct ConvP { const void* w; const float* b; int kt, kh, kw, cin, cout; };
struct NormP { const float* g; const float* b; };
struct YbP { const void* w; const float* b; int col, n; };     // fused conv_y | conv_b 1x1x1 weights [2C, L]; column offset in the concatenation

class Arena {
 public:
  Arena(char* base, int64_t cap, bool dry) : base_(base), cap_(cap), dry_(dry) { free_[0] = dry ? (int64_t(1) << 60) : cap; }
  char* alloc(int64_t bytes) {
    bytes = (bytes + 1023) & ~int64_t(1023);
    for (auto it = free_.begin(); it != free_.end(); ++it) {
      if (it->second >= bytes) {
        const int64_t off = it->first, len = it->second;
        free_.erase(it);
        if (len > bytes) free_[off + bytes] = len - bytes;
        used_[off] = bytes;
        if (off + bytes > peak_) peak_ = off + bytes;
        return fake() + off;
      }
    }
    failed_ = true;
    return nullptr;
  }
  void release(const void* p) {
    if (!p) return;
    const int64_t off = reinterpret_cast<const char*>(p) - fake();
    auto u = used_.find(off);
    if (u == used_.end()) return;
    int64_t o = off, len = u->second;
    used_.erase(u);
    auto nx = free_.lower_bound(o);
    if (nx != free_.end() && o + len == nx->first) { len += nx->second; nx = free_.erase(nx); }
    if (nx != free_.begin()) {
      auto pv = std::prev(nx);
      if (pv->first + pv->second == o) { o = pv->first; len += pv->second; free_.erase(pv); }
    }
    free_[o] = len;
  }
  int64_t peak() const { return peak_; }
  bool failed() const { return failed_; }
  bool dry() const { return dry_; }

 private:
  // dry mode hands out distinct fake addresses (never dereferenced) from a non-null base so that nullptr stays "none"
  char* fake() const { return dry_ ? reinterpret_cast<char*>(uintptr_t(1) << 40) : base_; }
  char* base_;
  int64_t cap_;
  bool dry_;
  std::map<int64_t, int64_t> free_, used_;
  int64_t peak_ = 0;
  bool failed_ = false;
};

struct Tn {          // channels-last activation [T, H, W, C] bf16
  char* p = nullptr;
  int T = 0, H = 0, W = 0, C = 0;
  int64_t frame_bytes() const { return int64_t(H) * W * C * 2; }
  int64_t bytes() const { return T * frame_bytes(); }
  int64_t rows() const { return int64_t(T) * H * W; }
};

}  // namespace

struct AetherVae {
  AetherVaeConfig cfg;
  std::unordered_map<std::string, ConvP> conv;
  std::unordered_map<std::string, NormP> norm;
  std::unordered_map<std::string, YbP> yb;
  YbP yb_all{nullptr, nullptr, 0, 0};                // every decoder conv_yb row-concatenated (optional)
};

namespace {

using Cache = std::unordered_map<std::string, Tn>;

struct Exec {
  const AetherVae* h;
  Arena* ar;
  cudaStream_t st;
  int rc = AETHER_OK;
  int64_t launches = 0;
  unsigned* counter = nullptr;    // zeroed word for the one-launch GroupNorm statistics (head of the workspace)
  char* zyb_all = nullptr;        // [rows(zq), yb_all.n] conv_y | conv_b of EVERY decoder SpatialNorm for the current batch

  bool dry() const { return ar->dry(); }
  bool ok() const { return rc == AETHER_OK; }
  void run(int r) {
    if (rc == AETHER_OK && r != AETHER_OK) rc = r;
  }
  Tn make(int T, int H, int W, int C) {
    Tn t;
    t.T = T; t.H = H; t.W = W; t.C = C;
    t.p = ar->alloc(t.bytes());
    if (!t.p) run(AETHER_ERR_WORKSPACE);
    return t;
  }
  void drop(Tn& t) {
    ar->release(t.p);
    t.p = nullptr;
  }
  void copy(void* dst, const void* src, int64_t bytes) {
    ++launches;
    if (dry() || !ok() || bytes <= 0) return;
    if (cudaMemcpyAsync(dst, src, size_t(bytes), cudaMemcpyDeviceToDevice, st) != cudaSuccess) run(AETHER_ERR_CUDA);
  }
  const ConvP* convp(const std::string& n) {
    auto it = h->conv.find(n);
    if (it == h->conv.end()) { fprintf(stderr, "[aether_b200] vae: missing conv parameter %s\n", n.c_str()); run(AETHER_ERR_INVALID); return nullptr; }
    return &it->second;
  }

  // ---- conv of an already time-padded input; `dst` (optional) = where the output goes (e.g. a slice of the tile output)
  Tn conv(const Tn& xpad, const std::string& name, int T_out, int stride, int pad_h, int pad_w, const Tn* resid,
          int Ho, int Wo, char* dst = nullptr) {
    Tn y;
    const ConvP* p = convp(name);
    if (!p) return y;
    if (xpad.C != p->cin || xpad.T != T_out + p->kt - 1) { run(AETHER_ERR_INVALID); return y; }
    y.T = T_out; y.H = Ho; y.W = Wo; y.C = p->cout;
    y.p = dst ? dst : ar->alloc(y.bytes());
    if (!y.p) { run(AETHER_ERR_WORKSPACE); return y; }
    ++launches;
    if (!dry() && ok())
      run(conv3d_bf16(xpad.p, xpad.T, xpad.H, xpad.W, xpad.C, p->w, p->b, resid ? resid->p : nullptr, y.p, T_out, Ho, Wo,
                      p->cout, p->kt, p->kh, p->kw, stride, pad_h, pad_w, st));
    return y;
  }

  // ---- GroupNorm / SpatialNorm3D + SiLU of x into `out` (same shape; may be the tail of a time-padded buffer)
  // `tail2` (optional): the last two frames of the result are ALSO written there (the next frame batch's conv cache)
  void norm(const Tn& x, const std::string& name, char* out, const Tn* zq, char* tail2 = nullptr) {
    const int64_t tail_from = tail2 ? int64_t(x.T - 2) * x.H * x.W : 0;
    const AetherVaeConfig& c = h->cfg;
    const int G = c.norm_num_groups;
    const bool spatial = zq != nullptr;
    // encoder GroupNorms use config.norm_eps; SpatialNorm3D's norm_layer and encoder.norm_out hard-code 1e-6
    const float eps = (spatial || name == "encoder.norm_out") ? 1e-6f : c.norm_eps;
    char* ws = ar->alloc(gn_workspace_floats(x.C) * 4 + 2 * G * 4 + 256);
    if (!ws) { run(AETHER_ERR_WORKSPACE); return; }
    float* partial = reinterpret_cast<float*>(ws);
    float* mr = reinterpret_cast<float*>(ws + ((gn_workspace_floats(x.C) * 4 + 255) & ~int64_t(255)));
    ++launches;
    if (!dry() && ok()) run(gn_stats(x.p, x.rows(), x.C, G, eps, partial, mr, counter, st));
    auto np = h->norm.find(spatial ? name + ".norm_layer" : name);
    if (np == h->norm.end()) { fprintf(stderr, "[aether_b200] vae: missing norm parameter %s\n", name.c_str()); run(AETHER_ERR_INVALID); ar->release(ws); return; }
    if (!spatial) {
      ++launches;
      if (!dry() && ok())
        run(gn_apply_imap(x.p, out, x.rows(), x.C, G, mr, np->second.g, np->second.b, nullptr, nullptr, 0, nullptr, x.H,
                          x.W, 1, 1, 1, tail2, tail_from, st));
      ar->release(ws);
      return;
    }
    auto yp = h->yb.find(name + ".conv_yb");
    if (yp == h->yb.end()) { run(AETHER_ERR_INVALID); ar->release(ws); return; }
    // conv_y / conv_b are 1x1x1: evaluate them at LATENT resolution (they commute with nearest interpolation) as a GEMM --
    // one for ALL norms of the frame batch when the concatenated weights were supplied (decoder() fills zyb_all)
    const int64_t nz = zq->rows();
    char* zyb = nullptr;
    int zld = 2 * x.C;
    if (zyb_all != nullptr && yp->second.n == 2 * x.C) {
      zld = h->yb_all.n;
    } else {
      zyb = ar->alloc(nz * 2 * x.C * 2);
      if (!zyb) { run(AETHER_ERR_WORKSPACE); ar->release(ws); return; }
      ++launches;
      if (!dry() && ok())
        run(gemm_bf16(zq->p, zq->C, yp->second.w, zq->C, zyb, 2 * x.C, (int)nz, 2 * x.C, zq->C, yp->second.b, 0, nullptr,
                      nullptr, 0, 0, 0, -1, st));
    }
    const char* zy = zyb ? zyb : zyb_all + int64_t(yp->second.col) * 2;
    IMap tmap;
    const int T = x.T, Tz = zq->T;
    if (T > IMap::kMax) { run(AETHER_ERR_INVALID); if (zyb) ar->release(zyb); ar->release(ws); return; }
    if (T > 1 && T % 2 == 1) {       // F.interpolate of the first frame separately from the rest (odd frame counts)
      tmap.v[0] = 0;
      for (int t = 1; t < T; ++t) tmap.v[t] = 1 + ((t - 1) * (Tz - 1)) / (T - 1);
    } else {
      for (int t = 0; t < T; ++t) tmap.v[t] = (t * Tz) / T;
    }
    ++launches;
    if (!dry() && ok())
      run(gn_apply_imap(x.p, out, x.rows(), x.C, G, mr, np->second.g, np->second.b, zy, zy + int64_t(x.C) * 2, zld, &tmap,
                        x.H, x.W, zq->H, zq->W, 1, tail2, tail_from, st));
    if (zyb) ar->release(zyb);
    ar->release(ws);
  }

  // ---- CogVideoXCausalConv3d time padding: [conv_cache] or [first frame] x 2 in front; returns the new cache (last 2).
  // `filled` = a cache buffer the producer of buf's payload has already written (norm's second destination).
  Tn fill_time_pad(const Tn& buf, const Cache& cache, const std::string& name, const Tn* filled = nullptr) {
    const int64_t fb = buf.frame_bytes();
    auto it = cache.find(name);
    if (it != cache.end()) {
      copy(buf.p, it->second.p, 2 * fb);
    } else {
      copy(buf.p, buf.p + 2 * fb, fb);
      copy(buf.p + fb, buf.p + 2 * fb, fb);
    }
    if (filled) return *filled;
    Tn nc = make(2, buf.H, buf.W, buf.C);
    if (nc.p) copy(nc.p, buf.p + int64_t(buf.T - 2) * fb, 2 * fb);
    return nc;
  }

  // causal conv of a plain tensor (conv_in): copy into a padded buffer first.  Does NOT consume x.
  Tn causal_from(const Tn& x, const std::string& name, const Cache& cache, Cache& nc) {
    Tn buf = make(x.T + 2, x.H, x.W, x.C);
    if (!buf.p) return Tn();
    copy(buf.p + 2 * buf.frame_bytes(), x.p, x.bytes());
    nc[name] = fill_time_pad(buf, cache, name);
    Tn y = conv(buf, name, x.T, 1, 1, 1, nullptr, x.H, x.W);
    drop(buf);
    return y;
  }

  // norm -> SiLU -> causal 3x3x3 conv (the norm writes straight into the time-padded conv input).  Does NOT consume x.
  Tn norm_conv(const Tn& x, const std::string& norm_name, const std::string& conv_name, const Cache& cache, Cache& nc,
               const Tn* zq, const Tn* resid, char* dst = nullptr) {
    Tn buf = make(x.T + 2, x.H, x.W, x.C);
    if (!buf.p) return Tn();
    // with two or more frames in the batch the new cache (= the last two payload frames) is written by the norm kernel
    // itself; a one-frame batch's cache also holds a padding frame and is copied out of the padded buffer afterwards
    Tn ncache;
    if (x.T >= 2) {
      ncache = make(2, x.H, x.W, x.C);
      if (!ncache.p) { drop(buf); return Tn(); }
    }
    norm(x, norm_name, buf.p + 2 * buf.frame_bytes(), zq, ncache.p);
    nc[conv_name] = fill_time_pad(buf, cache, conv_name, ncache.p ? &ncache : nullptr);
    Tn y = conv(buf, conv_name, x.T, 1, 1, 1, resid, x.H, x.W, dst);
    drop(buf);
    return y;
  }

  // CogVideoXResnetBlock3D; consumes x
  Tn resnet(Tn x, const std::string& name, const Cache& cache, Cache& nc, const Tn* zq) {
    Tn hmid = norm_conv(x, name + ".norm1", name + ".conv1", cache, nc, zq, nullptr);
    Tn r = x;
    const bool shortcut = h->conv.count(name + ".conv_shortcut") != 0;
    if (shortcut) r = conv(x, name + ".conv_shortcut", x.T, 1, 0, 0, nullptr, x.H, x.W);
    Tn out = norm_conv(hmid, name + ".norm2", name + ".conv2", cache, nc, zq, &r);
    drop(hmid);
    if (shortcut) drop(r);
    drop(x);
    return out;
  }

  // CogVideoXDownsample3D; consumes x
  Tn downsample(Tn x, const std::string& name, bool compress_time) {
    if (compress_time) {
      if (x.T > 2 * IMap::kMax - 1) { run(AETHER_ERR_INVALID); drop(x); return Tn(); }   // frame batch larger than the by-value maps
      IMap ia, ib;
      int To;
      if (x.T % 2 == 1) {
        To = 1 + (x.T - 1) / 2;
        ia.v[0] = 0; ib.v[0] = -1;
        for (int k = 1; k < To; ++k) { ia.v[k] = 2 * k - 1; ib.v[k] = 2 * k; }
      } else {
        To = x.T / 2;
        for (int k = 0; k < To; ++k) { ia.v[k] = 2 * k; ib.v[k] = 2 * k + 1; }
      }
      Tn y = make(To, x.H, x.W, x.C);
      ++launches;
      if (y.p && !dry() && ok()) run(avgpool_time_imap(x.p, y.p, ia, ib, To, int64_t(x.H) * x.W * x.C, st));
      drop(x);
      x = y;
    }
    const int Ho = (x.H + 1 - 3) / 2 + 1, Wo = (x.W + 1 - 3) / 2 + 1;       // F.pad (0,1,0,1) then 3x3 stride 2
    Tn y = conv(x, name + ".downsamplers.0.conv", x.T, 2, 0, 0, nullptr, Ho, Wo);
    drop(x);
    return y;
  }

  // CogVideoXUpsample3D; consumes x
  Tn upsample(Tn x, const std::string& name, bool compress_time) {
    IMap tmap;
    int To;
    if (2 * x.T > IMap::kMax) { run(AETHER_ERR_INVALID); drop(x); return Tn(); }
    if (compress_time && x.T > 1 && x.T % 2 == 1) {
      To = 1 + 2 * (x.T - 1);
      tmap.v[0] = 0;
      for (int t = 1; t < To; ++t) tmap.v[t] = 1 + (t - 1) / 2;
    } else if (compress_time && x.T > 1) {
      To = 2 * x.T;
      for (int t = 0; t < To; ++t) tmap.v[t] = t / 2;
    } else {
      To = x.T;
      for (int t = 0; t < To; ++t) tmap.v[t] = t;
    }
    Tn up = make(To, 2 * x.H, 2 * x.W, x.C);
    ++launches;
    if (up.p && !dry() && ok()) run(upsample_nearest_imap(x.p, up.p, tmap, To, 2 * x.H, 2 * x.W, x.H, x.W, 2, 2, x.C, st));
    drop(x);
    Tn y = conv(up, name + ".upsamplers.0.conv", To, 1, 1, 1, nullptr, up.H, up.W);
    drop(up);
    return y;
  }

  static std::string nm(const char* fmt, int i, int j = 0) {
    char b[128];
    snprintf(b, sizeof b, fmt, i, j);
    return b;
  }

  // one frame batch through the encoder; x [T, H, W, 8] is consumed; output (moments) goes to dst
  Tn encoder(Tn x, const Cache& cache, Cache& nc, char* dst) {
    const AetherVaeConfig& c = h->cfg;
    int tl = 0;
    while ((1 << tl) < c.temporal_compression_ratio) ++tl;
    Tn hcur = causal_from(x, "encoder.conv_in", cache, nc);
    drop(x);
    for (int i = 0; i < c.num_blocks && ok(); ++i) {
      for (int j = 0; j < c.layers_per_block && ok(); ++j)
        hcur = resnet(hcur, nm("encoder.down_blocks.%d.resnets.%d", i, j), cache, nc, nullptr);
      if (i != c.num_blocks - 1 && ok()) hcur = downsample(hcur, nm("encoder.down_blocks.%d", i), i < tl);
    }
    for (int j = 0; j < 2 && ok(); ++j) hcur = resnet(hcur, nm("encoder.mid_block.resnets.%d", j), cache, nc, nullptr);
    if (!ok()) return Tn();
    Tn y = norm_conv(hcur, "encoder.norm_out", "encoder.conv_out", cache, nc, nullptr, nullptr, dst);
    drop(hcur);
    return y;
  }

  // one latent frame batch through the decoder; z is NOT consumed (it is the SpatialNorm conditioning of every norm)
  Tn decoder(const Tn& z, const Cache& cache, Cache& nc, char* dst) {
    const AetherVaeConfig& c = h->cfg;
    int tl = 0;
    while ((1 << tl) < c.temporal_compression_ratio) ++tl;
    if (h->yb_all.w != nullptr) {     // every SpatialNorm's (scale | bias) table of this batch in one GEMM on the latent
      zyb_all = ar->alloc(z.rows() * int64_t(h->yb_all.n) * 2);
      if (!zyb_all) { run(AETHER_ERR_WORKSPACE); return Tn(); }
      ++launches;
      if (!dry() && ok())
        run(gemm_bf16(z.p, z.C, h->yb_all.w, z.C, zyb_all, h->yb_all.n, (int)z.rows(), h->yb_all.n, z.C, h->yb_all.b, 0,
                      nullptr, nullptr, 0, 0, 0, -1, st));
    }
    Tn hcur = causal_from(z, "decoder.conv_in", cache, nc);
    for (int j = 0; j < 2 && ok(); ++j) hcur = resnet(hcur, nm("decoder.mid_block.resnets.%d", j), cache, nc, &z);
    for (int i = 0; i < c.num_blocks && ok(); ++i) {
      for (int j = 0; j < c.layers_per_block + 1 && ok(); ++j)
        hcur = resnet(hcur, nm("decoder.up_blocks.%d.resnets.%d", i, j), cache, nc, &z);
      if (i != c.num_blocks - 1 && ok()) hcur = upsample(hcur, nm("decoder.up_blocks.%d", i), i < tl);
    }
    if (!ok()) return Tn();
    Tn y = norm_conv(hcur, "decoder.norm_out", "decoder.conv_out", cache, nc, &z, nullptr, dst);
    drop(hcur);
    if (zyb_all) { ar->release(zyb_all); zyb_all = nullptr; }
    return y;
  }

  // frame batching of diffusers (remainder folded into the first batch), conv caches carried between batches.
  // `in` is an NCTHW view (element strides sC, sT, sH) of the crop [C, T, Hc, Wc]; returns the tile output.
  Tn run_batched(bool enc, const char* in, int64_t sC, int64_t sT, int64_t sH, int C, int Cp, int T, int Hc, int Wc) {
    const AetherVaeConfig& c = h->cfg;
    const int fbs = enc ? c.num_sample_frames_batch_size : c.num_latent_frames_batch_size;
    const int nb = T / fbs > 0 ? T / fbs : 1;
    const int rem = T % fbs;
    int tcr_log = 0;
    while ((1 << tcr_log) < c.temporal_compression_ratio) ++tcr_log;
    const int scale = 1 << (c.num_blocks - 1);
    // per-batch frame ranges
    std::vector<std::pair<int, int>> ranges;
    for (int i = 0; i < nb; ++i)      // python slicing clips the end of the range to T (T < batch size: one short batch)
      ranges.push_back({fbs * i + (i == 0 ? 0 : rem), std::min(fbs * (i + 1) + rem, T)});
    int T_out = 0;
    std::vector<int> touts;
    for (int i = 0; i < nb; ++i) {
      // the odd-frame ("keep the first frame") rules act on each batch's own frame count, as in diffusers
      const int n = ranges[i].second - ranges[i].first;
      int t = n;
      for (int k = 0; k < tcr_log; ++k) {
        if (enc) t = (t % 2 == 1) ? 1 + (t - 1) / 2 : t / 2;
        else t = (t > 1 && t % 2 == 1) ? 1 + 2 * (t - 1) : (t > 1 ? 2 * t : t);
      }
      touts.push_back(t);
      T_out += t;
    }
    int Ho = Hc, Wo = Wc;
    if (enc) {
      for (int k = 0; k < c.num_blocks - 1; ++k) { Ho = (Ho + 1 - 3) / 2 + 1; Wo = (Wo + 1 - 3) / 2 + 1; }
    } else {
      Ho = Hc * scale; Wo = Wc * scale;
    }
    const int Cout = enc ? ((2 * c.latent_channels + 7) / 8) * 8 : ((c.out_channels + 7) / 8) * 8;
    Tn out = make(T_out, Ho, Wo, Cout);
    if (!out.p) return out;
    Cache cache;
    int t_off = 0;
    for (int i = 0; i < nb && ok(); ++i) {
      const int s = ranges[i].first, n = ranges[i].second - ranges[i].first;
      Tn x = make(n, Hc, Wc, Cp);
      if (!x.p) break;
      ++launches;
      if (!dry() && ok()) run(crop_ncthw_to_thwc(in + int64_t(s) * sT * 2, sC, sT, sH, x.p, C, Cp, n, Hc, Wc, st));
      Cache nc;
      char* dst = out.p + int64_t(t_off) * out.frame_bytes();
      Tn y = enc ? encoder(x, cache, nc, dst) : decoder(x, cache, nc, dst);
      if (!enc) drop(x);
      if (ok() && (y.T != touts[i] || y.H != Ho || y.W != Wo || y.C != Cout)) run(AETHER_ERR_INVALID);
      for (auto& kv : cache) ar->release(kv.second.p);
      cache.swap(nc);
      t_off += touts[i];
    }
    for (auto& kv : cache) ar->release(kv.second.p);
    return out;
  }

  // diffusers tiled_encode / tiled_decode: overlapping tiles, blend_v / blend_h, keep [:lim_h, :lim_w] of every tile
  Tn tiled(bool enc, const char* in, int64_t sC, int64_t sT, int64_t sH, int C, int Cp, int T, int H, int W) {
    const AetherVaeConfig& c = h->cfg;
    const int tile_h = enc ? c.tile_sample_min_height : c.tile_latent_min_height;
    const int tile_w = enc ? c.tile_sample_min_width : c.tile_latent_min_width;
    const int ov_h = enc ? c.enc_overlap_h : c.dec_overlap_h, ov_w = enc ? c.enc_overlap_w : c.dec_overlap_w;
    const int bl_h = enc ? c.enc_blend_h : c.dec_blend_h, bl_w = enc ? c.enc_blend_w : c.dec_blend_w;
    const int lim_h = enc ? c.enc_limit_h : c.dec_limit_h, lim_w = enc ? c.enc_limit_w : c.dec_limit_w;
    if (!(c.use_tiling && (W > tile_w || H > tile_h))) return run_batched(enc, in, sC, sT, sH, C, Cp, T, H, W);
    if (ov_h <= 0 || ov_w <= 0) { run(AETHER_ERR_INVALID); return Tn(); }
    std::vector<std::vector<Tn>> rows;
    for (int i = 0; i < H && ok(); i += ov_h) {
      std::vector<Tn> row;
      for (int j = 0; j < W && ok(); j += ov_w) {
        const int th = (i + tile_h <= H) ? tile_h : H - i, tw = (j + tile_w <= W) ? tile_w : W - j;
        row.push_back(run_batched(enc, in + (int64_t(i) * sH + j) * 2, sC, sT, sH, C, Cp, T, th, tw));
      }
      rows.push_back(row);
    }
    if (!ok()) return Tn();
    // assembled size = sum over rows / cols of min(tile extent, limit)
    int Ho = 0, Wo = 0;
    for (auto& row : rows) Ho += row[0].H < lim_h ? row[0].H : lim_h;
    for (auto& t : rows[0]) Wo += t.W < lim_w ? t.W : lim_w;
    const Tn& t00 = rows[0][0];
    Tn out = make(t00.T, Ho, Wo, t00.C);
    if (!out.p) return out;
    int y0 = 0;
    for (size_t i = 0; i < rows.size() && ok(); ++i) {
      int x0 = 0;
      for (size_t j = 0; j < rows[i].size() && ok(); ++j) {
        Tn& tile = rows[i][j];
        if (i > 0) {
          const Tn& a = rows[i - 1][j];
          const int e = std::min(std::min(a.H, tile.H), bl_h);
          ++launches;
          if (!dry() && ok()) run(tile_blend(a.p, tile.p, tile.T, a.H, a.W, tile.H, tile.W, tile.C, 1, e, st));
        }
        if (j > 0) {
          const Tn& a = rows[i][j - 1];
          const int e = std::min(std::min(a.W, tile.W), bl_w);
          ++launches;
          if (!dry() && ok()) run(tile_blend(a.p, tile.p, tile.T, a.H, a.W, tile.H, tile.W, tile.C, 2, e, st));
        }
        const int hh = tile.H < lim_h ? tile.H : lim_h, ww = tile.W < lim_w ? tile.W : lim_w;
        ++launches;
        if (!dry() && ok()) run(copy_region_cl(tile.p, tile.H, tile.W, out.p, Ho, Wo, tile.T, hh, ww, tile.C, y0, x0, st));
        x0 += ww;
      }
      y0 += rows[i][0].H < lim_h ? rows[i][0].H : lim_h;
    }
    for (auto& row : rows)
      for (auto& t : row) drop(t);
    return out;
  }
};

int check_cfg(const AetherVaeConfig& c) {
  if (c.num_blocks < 2 || c.num_blocks > 8 || c.layers_per_block < 1 || c.norm_num_groups < 1) return AETHER_ERR_INVALID;
  if (c.latent_channels % 8 != 0 || c.in_channels > 8 || c.out_channels > 8) return AETHER_ERR_INVALID;
  if (c.num_latent_frames_batch_size < 1 || c.num_sample_frames_batch_size < 1) return AETHER_ERR_INVALID;
  int tl = 0;
  while ((1 << tl) < c.temporal_compression_ratio) ++tl;
  if ((1 << tl) != c.temporal_compression_ratio || tl > c.num_blocks - 1) return AETHER_ERR_INVALID;
  return AETHER_OK;
}

// the shared body of the size query (dry) and the real calls
int vae_run_impl(const AetherVae* h, bool enc, const void* in, int64_t sC, int64_t sT, int64_t sH, int T, int H, int W, void* out,
            void* workspace, int64_t workspace_bytes, cudaStream_t st, bool dry, int64_t* peak, int64_t* launches,
            int32_t* dims) {
  const AetherVaeConfig& c = h->cfg;
  char* base = dry ? nullptr : reinterpret_cast<char*>((reinterpret_cast<uintptr_t>(workspace) + 1023) & ~uintptr_t(1023));
  const int64_t cap = dry ? 0 : workspace_bytes - (base - reinterpret_cast<char*>(workspace));
  constexpr int64_t kHead = 1024;           // head of the workspace: the GroupNorm statistics' arrival counter
  if (!dry && (!workspace || cap <= kHead)) return AETHER_ERR_WORKSPACE;
  Arena ar(dry ? nullptr : base + kHead, dry ? 0 : cap - kHead, dry);
  Exec ex{h, &ar, st};
  ++ex.launches;
  if (!dry) {
    ex.counter = reinterpret_cast<unsigned*>(base);
    if (cudaMemsetAsync(base, 0, size_t(kHead), st) != cudaSuccess) return AETHER_ERR_CUDA;
  }
  const int C = enc ? c.in_channels : c.latent_channels;
  const int Cp = ((C + 7) / 8) * 8;
  Tn res = ex.tiled(enc, reinterpret_cast<const char*>(in), sC, sT, sH, C, Cp, T, H, W);
  if (ex.ok() && !res.p) ex.run(AETHER_ERR_WORKSPACE);
  if (ex.ok()) {
    if (dims) { dims[0] = res.T; dims[1] = res.H; dims[2] = res.W; dims[3] = res.C; }
    if (enc) {
      ex.copy(out, res.p, res.bytes());     // moments stay channels-last [T', h, w, Cp2] (aether_posterior_sample's input)
    } else {
      ++ex.launches;
      if (!dry) ex.run(thwc_to_ncthw(res.p, out, c.out_channels, res.C, res.rows(), st));
    }
  }
  if (peak) *peak = ar.peak() + kHead + 2048;
  if (launches) *launches = ex.launches;
  if (ar.failed() && ex.rc == AETHER_OK) return AETHER_ERR_WORKSPACE;
  return ex.rc;
}

// The host-side bookkeeping uses std containers; nothing may propagate through the C boundary.
int vae_run(const AetherVae* h, bool enc, const void* in, int64_t sC, int64_t sT, int64_t sH, int T, int H, int W, void* out,
            void* workspace, int64_t workspace_bytes, cudaStream_t st, bool dry, int64_t* peak, int64_t* launches,
            int32_t* dims = nullptr) {
  try {
    return vae_run_impl(h, enc, in, sC, sT, sH, T, H, W, out, workspace, workspace_bytes, st, dry, peak, launches, dims);
  } catch (const std::exception& e) {
    fprintf(stderr, "[aether_b200] vae: host-side failure: %s\n", e.what());
    return AETHER_ERR_INVALID;
  }
}

}  // namespace

extern "C" {

int aether_vae_create(const AetherVaeConfig* cfg, const AetherVaeParam* params, int32_t n_params, AetherVae** out) {
  if (!cfg || !params || n_params <= 0 || !out) return AETHER_ERR_INVALID;
  if (check_cfg(*cfg) != AETHER_OK) {
    fprintf(stderr, "[aether_b200] vae_create: unsupported configuration\n");
    return AETHER_ERR_INVALID;
  }
  AetherVae* h = new (std::nothrow) AetherVae;
  if (!h) return AETHER_ERR_INVALID;
  h->cfg = *cfg;
  try {
    for (int i = 0; i < n_params; ++i) {
      const AetherVaeParam& p = params[i];
      if (!p.name || !p.data) { delete h; return AETHER_ERR_INVALID; }
      const std::string name(p.name);
      if (p.kind == 0) h->conv[name] = ConvP{p.data, p.bias, p.kt, p.kh, p.kw, p.cin, p.cout};
      else if (p.kind == 1) h->norm[name] = NormP{reinterpret_cast<const float*>(p.data), p.bias};
      else if (p.kind == 2) h->yb[name] = YbP{p.data, p.bias, p.cin, p.cout};
      else if (p.kind == 3) h->yb_all = YbP{p.data, p.bias, 0, p.cout};
      else { delete h; return AETHER_ERR_INVALID; }
    }
  } catch (const std::exception&) {
    delete h;
    return AETHER_ERR_INVALID;
  }
  *out = h;
  return AETHER_OK;
}

void aether_vae_destroy(AetherVae* h) { delete h; }

int64_t aether_vae_workspace_bytes(const AetherVae* h, int32_t op, int32_t T, int32_t H, int32_t W) {
  if (!h || T <= 0 || H <= 0 || W <= 0 || (op != 0 && op != 1)) return -1;
  int64_t peak = 0;
  const int rc = vae_run(h, op == 0, reinterpret_cast<const void*>(uintptr_t(1) << 41), int64_t(T) * H * W, int64_t(H) * W, W, T,
                         H, W, nullptr, nullptr, 0, nullptr, true, &peak, nullptr);
  return rc == AETHER_OK ? peak : -1;
}

int aether_vae_output_shape(const AetherVae* h, int32_t op, int32_t T, int32_t H, int32_t W, int32_t* dims4) {
  if (!h || !dims4 || T <= 0 || H <= 0 || W <= 0 || (op != 0 && op != 1)) return AETHER_ERR_INVALID;
  return vae_run(h, op == 0, reinterpret_cast<const void*>(uintptr_t(1) << 41), int64_t(T) * H * W, int64_t(H) * W, W, T, H, W,
                 nullptr, nullptr, 0, nullptr, true, nullptr, nullptr, dims4);
}

int64_t aether_vae_launch_count(const AetherVae* h, int32_t op, int32_t T, int32_t H, int32_t W) {
  if (!h || T <= 0 || H <= 0 || W <= 0 || (op != 0 && op != 1)) return -1;
  int64_t peak = 0, n = 0;
  const int rc = vae_run(h, op == 0, reinterpret_cast<const void*>(uintptr_t(1) << 41), int64_t(T) * H * W, int64_t(H) * W, W, T,
                         H, W, nullptr, nullptr, 0, nullptr, true, &peak, &n);
  return rc == AETHER_OK ? n : -1;
}

int aether_vae_encode(const AetherVae* h, const void* x, int64_t stride_c, int64_t stride_t, int64_t stride_h, int32_t T,
                      int32_t H, int32_t W, void* moments, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h || !x || !moments || T <= 0 || H <= 0 || W <= 0) return AETHER_ERR_INVALID;
  return vae_run(h, true, x, stride_c, stride_t, stride_h, T, H, W, moments, workspace, workspace_bytes,
                 reinterpret_cast<cudaStream_t>(stream), false, nullptr, nullptr);
}

int aether_vae_decode(const AetherVae* h, const void* z, int64_t stride_c, int64_t stride_t, int64_t stride_h, int32_t T,
                      int32_t H, int32_t W, void* sample, void* workspace, int64_t workspace_bytes, void* stream) {
  if (!h || !z || !sample || T <= 0 || H <= 0 || W <= 0) return AETHER_ERR_INVALID;
  return vae_run(h, false, z, stride_c, stride_t, stride_h, T, H, W, sample, workspace, workspace_bytes,
                 reinterpret_cast<cudaStream_t>(stream), false, nullptr, nullptr);
}
}
