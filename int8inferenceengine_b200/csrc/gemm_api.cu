// C-ABI entry points for the GEMM-shaped ops: conv2d plan/run and fully-connected.
// Dispatch is by SHAPE only (alignment / size), never by backend: both kernels are sm_100a
// CUDA; the tcgen05 kernel takes every shape whose K blocks are TMA/UMMA-aligned.
#include <cstdlib>
#include <mutex>
#include <new>
#include <vector>

#include "gemm_api.cuh"

using namespace i8ie;

namespace {

struct MapCacheEntry {
  const void* ptr;
  long long a, b, c, d;
  CUtensorMap map;
};

// small pointer-keyed cache: encoding a CUtensorMap costs microseconds on the host and the
// activation buffers of a served model are recycled (CUDA-graph replays reuse them exactly)
class MapCache {
 public:
  template <typename MakeFn>
  int get(const void* ptr, long long a, long long b, long long c, long long d, CUtensorMap* out, MakeFn make) {
    std::lock_guard<std::mutex> lk(mu_);
    for (auto& e : entries_)
      if (e.ptr == ptr && e.a == a && e.b == b && e.c == c && e.d == d) { *out = e.map; return I8IE_OK; }
    MapCacheEntry e{ptr, a, b, c, d, {}};
    int rc = make(&e.map);
    if (rc != I8IE_OK) return rc;
    if (entries_.size() >= 256) entries_.erase(entries_.begin());
    entries_.push_back(e);
    *out = e.map;
    return I8IE_OK;
  }

 private:
  std::mutex mu_;
  std::vector<MapCacheEntry> entries_;
};

MapCache g_fc_maps;

// fc weight buffers that have a tiled, pre-swizzled copy (i8ie_fc_weight_tiled_attach): the owner of both
// buffers registers the pair and removes it before freeing either, so a hit is never stale.
struct TiledWeight {
  const int8_t* w;
  const int8_t* wt;
  int n_pad, ldw;
};
std::mutex g_tiled_mu;
std::vector<TiledWeight> g_tiled;

const int8_t* tiled_copy_of(const int8_t* w, int n_pad, int ldw) {
  std::lock_guard<std::mutex> lk(g_tiled_mu);
  for (const auto& e : g_tiled)
    if (e.w == w && e.n_pad == n_pad && e.ldw == ldw) return e.wt;
  return nullptr;
}

bool tc_disabled() {
  static const bool off = std::getenv("I8IE_DISABLE_TC") != nullptr;
  return off;
}

}  // namespace

struct i8ie_conv_plan {
  GemmGeom g;
  const int8_t* w_packed;
  int impl;  // 1 = SIMT dp4a, 2 = tcgen05 im2col, 3 = tcgen05 stem (small-C strided first layer)
  int c;     // real input channels
  // tcgen05 state
  int bk, bn;
  int pair_mt;          // pair kernel: 256-row accumulators per tile (1 or 2)
  int cluster;          // CTAs per cluster (weight tile multicast); tmB's box holds bn / cluster rows
  CUtensorMap tmB;
  int32_t* border_tab;  // device, owned
  MapCache amaps;
  // stem state (all device buffers owned by the plan)
  StemGeom stem;
  uint8_t* stem_x;      // bordered superpixel image, rewritten by every call
  int8_t* stem_w;       // [kc_pad][kh][64]
  int8_t* stem_w48;     // [kc_pad][k48]: K-packed weights of the stem2 kernel (kw <= 12), or nullptr
  CUtensorMap tmB48;
  CUtensorMap tmA_stem;
  bool stem2;           // smem-resident-row kernel (stride 4, narrow rows)
  // row mode (impl 4): physically padded input, plan-owned [kc_pad][kh][kr] weights
  int kr;
  int8_t* row_w;
  GemmGeom gv;          // the virtual geometry the pair kernel runs with
  // F4 extension: per-output-channel weight scales (device [kc], owned by the caller) or nullptr
  const float* sb_vec;
  float sb_min, sb_max;
};

static void apply_channel_scales(const i8ie_conv_plan* plan, EpiParams* ep) {
  if (plan->sb_vec != nullptr) { ep->sb_vec = plan->sb_vec; ep->sb_min = plan->sb_min; ep->sb_max = plan->sb_max; }
}

static void plan_free(i8ie_conv_plan* p) {
  if (!p) return;
  if (p->border_tab) cudaFree(p->border_tab);
  if (p->stem_x) cudaFree(p->stem_x);
  if (p->stem_w) cudaFree(p->stem_w);
  if (p->stem_w48) cudaFree(p->stem_w48);
  if (p->row_w) cudaFree(p->row_w);
  delete p;
}

extern "C" {

i8ie_conv_plan* i8ie_conv2d_plan_create(int n, int c, int h, int w, int cp, int kc, int kh, int kw,
                                        int stride, int pad, int out_cp, const int8_t* w_packed,
                                        int kc_pad, int impl) {
  if (n <= 0 || c <= 0 || h <= 0 || w <= 0 || kc <= 0 || kh <= 0 || kw <= 0 || stride <= 0 || pad < 0 ||
      cp % 16 != 0 || cp < c || out_cp % 16 != 0 || out_cp < kc || kc_pad < kc || w_packed == nullptr) {
    set_error("conv2d_plan_create: bad geometry (n=%d c=%d h=%d w=%d cp=%d kc=%d k=%dx%d s=%d p=%d out_cp=%d kc_pad=%d)",
              n, c, h, w, cp, kc, kh, kw, stride, pad, out_cp, kc_pad);
    return nullptr;
  }
  if (h + 2 * pad < kh || w + 2 * pad < kw) {
    set_error("conv2d_plan_create: kernel larger than padded input");
    return nullptr;
  }
  tc_error_sink_init();   // plans are created eagerly (never under stream capture)
  i8ie_conv_plan* p = new (std::nothrow) i8ie_conv_plan();
  if (!p) { set_error("conv2d_plan_create: out of host memory"); return nullptr; }
  GemmGeom& g = p->g;
  g.n = n; g.h = h; g.w = w; g.cp = cp;
  g.kh = kh; g.kw = kw; g.stride = stride; g.pad = pad;
  g.oh = (h - kh + 2 * pad) / stride + 1;   // conv2d.cc:108
  g.ow = (w - kw + 2 * pad) / stride + 1;   // conv2d.cc:109
  g.M = n * g.oh * g.ow; g.N = kc; g.n_pad = kc_pad;
  g.ldw = kh * kw * cp; g.out_cp = out_cp;
  p->w_packed = w_packed;
  p->border_tab = nullptr;
  p->stem_x = nullptr;
  p->stem_w = nullptr;
  p->stem_w48 = nullptr;
  p->row_w = nullptr;
  p->kr = 0;
  p->sb_vec = nullptr;
  p->sb_min = p->sb_max = 0.f;
  p->c = c;
  p->cluster = 1;
  p->pair_mt = 1;
  if (impl == 4) {
    // row mode: x will be the physically padded tensor [n][h + 2p][w + 2p][cp]
    if (tc_disabled() || !tc_row_mode_ok(c, kh, kw, stride, pad, out_cp) || cp != (c + 15) / 16 * 16) {
      set_error("conv2d_plan_create: geometry not eligible for the row-mode kernel (c=%d cp=%d k=%dx%d s=%d p=%d out_cp=%d)",
                c, cp, kh, kw, stride, pad, out_cp);
      plan_free(p);
      return nullptr;
    }
    p->impl = 4;
    p->kr = tc_row_mode_kr(cp, kw);
    p->gv = tc_row_mode_geom(g, p->kr);
    p->bk = 128;
    p->bn = tc_pick_bn_pair(p->gv, 128, &p->pair_mt);
    p->cluster = tc_conv_cluster(128, p->bn);
    int rc4 = I8IE_OK;
    if (cudaMalloc(&p->row_w, (size_t)kc_pad * kh * p->kr) != cudaSuccess) {
      set_error("conv2d_plan_create: cudaMalloc of the row-mode weights failed");
      rc4 = I8IE_ECUDA;
    }
    if (rc4 == I8IE_OK) rc4 = tc_row_pack_weights(g, w_packed, p->row_w, p->kr, 0);
    if (rc4 == I8IE_OK && cudaStreamSynchronize(0) != cudaSuccess) {
      set_error("conv2d_plan_create: row-mode weight kernel failed");
      rc4 = I8IE_ECUDA;
    }
    if (rc4 == I8IE_OK)
      rc4 = tc_encode_weight_map(&p->tmB, p->row_w, kc_pad, kh * p->kr, 128, p->cluster > 1 ? tc_pair_box_rows(p->bn) : p->bn);
    if (rc4 != I8IE_OK) { plan_free(p); return nullptr; }
    return p;
  }
  const bool stem_ok = tc_stem_eligible(g, c) && !tc_disabled();
  const bool eligible = tc_conv_eligible(g) && !tc_disabled();
  if (impl == 2 && !eligible && !stem_ok) {
    set_error("conv2d_plan_create: geometry not eligible for a tcgen05 kernel (needs cp %% 32 == 0 and pad <= 3, "
              "or c <= 4 with stride 4/8); cp=%d c=%d stride=%d", cp, c, stride);
    plan_free(p);
    return nullptr;
  }
  p->impl = (impl == 1 || (!eligible && !stem_ok)) ? 1 : (stem_ok ? 3 : 2);
  int rc = I8IE_OK;
  if (p->impl == 3) {
    p->stem = tc_stem_geom(g, c);
    p->stem2 = tc_stem2_eligible(g, c) && std::getenv("I8IE_NO_STEM2") == nullptr;
    p->bk = 64;
    p->bn = tc_pick_bn(g.out_cp);   // every lane of the padded output pitch is written
    // stem2 covers all real channels with one N tile and fills the pad lanes with plain stores
    if (p->stem2 && tc_pick_bn(kc) < p->bn && tc_pick_bn(kc) % 32 == 0 && tc_pick_bn(kc) >= kc && kc_pad >= tc_pick_bn(kc))
      p->bn = tc_pick_bn(kc);
    if (cudaMalloc(&p->stem_x, (size_t)p->stem.bytes) != cudaSuccess ||
        cudaMalloc(&p->stem_w, (size_t)kc_pad * kh * 64) != cudaSuccess) {
      set_error("conv2d_plan_create: cudaMalloc of the stem buffers failed");
      rc = I8IE_ECUDA;
    }
    if (rc == I8IE_OK) rc = tc_stem_pack_weights(g, c, w_packed, p->stem_w, 0);
    if (rc == I8IE_OK && cudaStreamSynchronize(0) != cudaSuccess) {
      set_error("conv2d_plan_create: stem weight kernel failed");
      rc = I8IE_ECUDA;
    }
    if (rc == I8IE_OK) rc = tc_encode_weight_map(&p->tmB, p->stem_w, kc_pad, kh * 64, 64, p->bn);
    const int k48 = p->stem2 ? tc_stem_k48_bytes(g, c) : 0;
    if (rc == I8IE_OK && k48 > 0) {
      if (cudaMalloc(&p->stem_w48, (size_t)kc_pad * k48) != cudaSuccess) {
        set_error("conv2d_plan_create: cudaMalloc of the packed stem weights failed");
        rc = I8IE_ECUDA;
      }
      if (rc == I8IE_OK) rc = tc_stem_pack_weights48(g, c, w_packed, p->stem_w48, 0);
      if (rc == I8IE_OK && cudaStreamSynchronize(0) != cudaSuccess) {
        set_error("conv2d_plan_create: stem weight kernel failed");
        rc = I8IE_ECUDA;
      }
      if (rc == I8IE_OK) rc = tc_encode_weight_map(&p->tmB48, p->stem_w48, kc_pad, k48, 32, p->bn);
    }
    if (rc == I8IE_OK) rc = tc_encode_stem_act_map(&p->tmA_stem, p->stem_x, g, p->stem);
    if (rc != I8IE_OK && impl != 2) {
      // the overlapping-window tensor map was refused: serve the layer with the SIMT kernel
      cudaFree(p->stem_x); cudaFree(p->stem_w);
      p->stem_x = nullptr; p->stem_w = nullptr;
      p->impl = eligible ? 2 : 1;
      rc = I8IE_OK;
    }
  }
  if (rc == I8IE_OK && p->impl == 2) {
    p->bk = tc_conv_bk(g);
    p->bn = tc_pick_bn_pair(g, p->bk, &p->pair_mt);   // every lane of the padded output pitch is written
    p->cluster = tc_conv_cluster(p->bk, p->bn);
    rc = tc_encode_weight_map(&p->tmB, w_packed, kc_pad, g.ldw, p->bk, p->cluster > 1 ? tc_pair_box_rows(p->bn) : p->bn);
    const int tab = tc_border_table_size(g);
    if (rc == I8IE_OK && tab > 0) {
      if (cudaMalloc(&p->border_tab, sizeof(int32_t) * (size_t)tab) != cudaSuccess) {
        set_error("conv2d_plan_create: cudaMalloc of the border table failed");
        rc = I8IE_ECUDA;
      } else {
        rc = tc_build_border_table(g, w_packed, p->border_tab, 0);
        if (rc == I8IE_OK && cudaStreamSynchronize(0) != cudaSuccess) {
          set_error("conv2d_plan_create: border table kernel failed");
          rc = I8IE_ECUDA;
        }
      }
    }
  }
  if (rc != I8IE_OK) {
    plan_free(p);
    return nullptr;
  }
  return p;
}

void i8ie_conv2d_plan_destroy(i8ie_conv_plan* plan) { plan_free(plan); }

int i8ie_conv2d_plan_impl(const i8ie_conv_plan* plan) { return plan ? plan->impl : 0; }

int i8ie_conv2d_plan_set_channel_scales(i8ie_conv_plan* plan, const float* sb_vec, float sb_min, float sb_max) {
  I8IE_REQUIRE(plan != nullptr, "conv2d_plan_set_channel_scales: null plan");
  I8IE_REQUIRE(sb_vec == nullptr || (sb_min > 0.f && sb_max >= sb_min), "conv2d_plan_set_channel_scales: bad bounds");
  plan->sb_vec = sb_vec;
  plan->sb_min = sb_min;
  plan->sb_max = sb_max;
  return I8IE_OK;
}

int i8ie_conv2d_row_mode_cp(int c, int cp_plain, int kh, int kw, int stride, int pad, int out_cp) {
  if (tc_disabled()) return 0;
  return tc_row_mode_cp(c, cp_plain, kh, kw, stride, pad, out_cp);
}

int i8ie_conv2d_u8(i8ie_conv_plan* plan, const uint8_t* x, uint8_t* y, const int32_t* oc, float sa,
                   float sb, float sc, int zp_in, int zp_out, int flags, int32_t* acc_out, void* stream) {
  I8IE_REQUIRE(plan && x && y && oc, "conv2d_u8: null argument");
  I8IE_REQUIRE(zp_in >= 0 && zp_in <= 255 && zp_out >= 0 && zp_out <= 255, "conv2d_u8: zero point out of range");
  EpiParams ep{oc, nullptr, sa, sb, sc, zp_out, (flags & I8IE_EPI_RELU) ? 1 : 0, acc_out};
  apply_channel_scales(plan, &ep);
  tc_error_sink_touch((cudaStream_t)stream);
  if (plan->impl == 3) {
    int rc = tc_stem_pack_input(plan->g, plan->stem, x, plan->stem_x, zp_in, (cudaStream_t)stream);
    if (rc != I8IE_OK) return rc;
    if (plan->stem2)
      return launch_tc_stem2(plan->g, plan->stem, plan->stem_x, nullptr, plan->stem_w48 ? plan->tmB48 : plan->tmB, plan->bn, y, ep,
                           (cudaStream_t)stream, plan->stem_w48 != nullptr);
    return launch_tc_stem(plan->g, plan->tmA_stem, plan->tmB, plan->bn, y, ep, (cudaStream_t)stream);
  }
  if (plan->impl == 4) {   // x = physically padded input (border = zp_in): no border table, exact by construction
    CUtensorMap tmA;
    int rc = plan->amaps.get(x, 4, 0, 0, 0, &tmA,
                             [&](CUtensorMap* m) { return tc_encode_act_map_row_mode(m, x, plan->g, plan->kr); });
    if (rc != I8IE_OK) return rc;
    return launch_tc_conv(plan->gv, tmA, plan->tmB, 128, plan->bn, plan->cluster, nullptr, y, ep, zp_in,
                          (cudaStream_t)stream, plan->pair_mt);
  }
  if (plan->impl == 2) {
    CUtensorMap tmA;
    int rc = plan->amaps.get(x, 0, 0, 0, 0, &tmA,
                             [&](CUtensorMap* m) { return tc_encode_act_map_im2col(m, x, plan->g, plan->bk); });
    if (rc != I8IE_OK) return rc;
    return launch_tc_conv(plan->g, tmA, plan->tmB, plan->bk, plan->bn, plan->cluster, plan->border_tab, y, ep, zp_in,
                          (cudaStream_t)stream, plan->pair_mt);
  }
  return launch_simt_igemm(plan->g, x, plan->w_packed, y, ep, zp_in, (cudaStream_t)stream);
}

static int conv2d_f32_u8(i8ie_conv_plan* plan, const float* x_nchw, const float* const* x_slot, float in_scale,
                        int in_zp, uint8_t* y, const int32_t* oc, float sb, float sc, int zp_out, int flags,
                        int32_t* acc_out, void* stream) {
  I8IE_REQUIRE(plan && (x_nchw || x_slot) && y && oc, "conv2d_f32_u8: null argument");
  I8IE_REQUIRE(plan->impl == 3, "conv2d_f32_u8: only stem plans fuse the input quantise (plan impl=%d)", plan->impl);
  I8IE_REQUIRE(in_zp >= 0 && in_zp <= 255 && zp_out >= 0 && zp_out <= 255, "conv2d_f32_u8: zero point out of range");
  EpiParams ep{oc, nullptr, in_scale, sb, sc, zp_out, (flags & I8IE_EPI_RELU) ? 1 : 0, acc_out};
  apply_channel_scales(plan, &ep);
  tc_error_sink_touch((cudaStream_t)stream);
  if (plan->stem2 && tc_stem2_can_fuse_quantize(plan->g, plan->c, x_nchw, x_slot, in_scale)) {
    // the stem kernel's producer warps quantise the fp32 image straight into its operand ring
    const StemF32Src src{x_nchw, x_slot, in_scale, in_zp};
    return launch_tc_stem2(plan->g, plan->stem, plan->stem_x, &src, plan->stem_w48 ? plan->tmB48 : plan->tmB, plan->bn, y, ep,
                           (cudaStream_t)stream, plan->stem_w48 != nullptr);
  }
  int rc = tc_stem_quantize_input(plan->g, plan->stem, x_nchw, x_slot, plan->stem_x, in_scale, in_zp,
                                  (cudaStream_t)stream);
  if (rc != I8IE_OK) return rc;
  if (plan->stem2)
    return launch_tc_stem2(plan->g, plan->stem, plan->stem_x, nullptr, plan->stem_w48 ? plan->tmB48 : plan->tmB, plan->bn, y, ep,
                           (cudaStream_t)stream, plan->stem_w48 != nullptr);
  return launch_tc_stem(plan->g, plan->tmA_stem, plan->tmB, plan->bn, y, ep, (cudaStream_t)stream);
}

int i8ie_conv2d_f32_u8(i8ie_conv_plan* plan, const float* x_nchw, float in_scale, int in_zp, uint8_t* y,
                       const int32_t* oc, float sb, float sc, int zp_out, int flags, int32_t* acc_out,
                       void* stream) {
  I8IE_REQUIRE(x_nchw != nullptr, "conv2d_f32_u8: null input");
  return conv2d_f32_u8(plan, x_nchw, nullptr, in_scale, in_zp, y, oc, sb, sc, zp_out, flags, acc_out, stream);
}

int i8ie_conv2d_f32_u8_indirect(i8ie_conv_plan* plan, const float* const* x_slot, float in_scale, int in_zp,
                                uint8_t* y, const int32_t* oc, float sb, float sc, int zp_out, int flags,
                                int32_t* acc_out, void* stream) {
  I8IE_REQUIRE(x_slot != nullptr, "conv2d_f32_u8_indirect: null slot");
  return conv2d_f32_u8(plan, nullptr, x_slot, in_scale, in_zp, y, oc, sb, sc, zp_out, flags, acc_out, stream);
}

static int fc_u8(const uint8_t* x, int ldx, const int8_t* w, int ldw, int n_pad, uint8_t* y, int ldy,
                 int m, int n, int k, const int32_t* oc, const float* bias_f, float sa, float sb,
                 const float* sb_vec, float sb_min, float sb_max, float sc,
                 int zp_out, int flags, int32_t* acc_out, int impl, void* stream, float* deq_out = nullptr) {
  I8IE_REQUIRE(x && w && y && oc && bias_f, "fc_u8: null argument");
  I8IE_REQUIRE(m > 0 && n > 0 && k > 0 && ldx % 16 == 0 && ldw % 16 == 0 && ldy % 16 == 0 && ldx >= k &&
                   ldw >= ldx && ldy >= n && n_pad >= n,
               "fc_u8: bad shape/pitch (m=%d n=%d k=%d ldx=%d ldw=%d ldy=%d n_pad=%d)", m, n, k, ldx, ldw, ldy, n_pad);
  I8IE_REQUIRE(zp_out >= 0 && zp_out <= 255, "fc_u8: zero point out of range");
  EpiParams ep{oc, bias_f, sa, sb, sc, zp_out, (flags & I8IE_EPI_RELU) ? 1 : 0, acc_out};
  if (sb_vec != nullptr) { ep.sb_vec = sb_vec; ep.sb_min = sb_min; ep.sb_max = sb_max; }
  tc_error_sink_touch((cudaStream_t)stream);
  // shape dispatch: the tensor-core kernel needs at least one full 32-byte K step to be worthwhile
  const bool eligible = !tc_disabled() && k >= 32;
  I8IE_REQUIRE(!(impl == 2 && !eligible), "fc_u8: shape not eligible for the tcgen05 kernel");
  // shape dispatch: a classifier head (<= 16 outputs) is one warp per row, not a 128-row MMA tile
  if ((impl == 0 || impl == 3) && fc_head_eligible(n_pad, ldx, ldw, ldy, x, w)) {
    ep.deq_out = deq_out;   // the head kernel writes the dequantised logits itself
    return launch_fc_head(x, ldx, w, ldw, y, ldy, m, n, k, ep, (cudaStream_t)stream);
  }
  // every other kernel: u8 result first, then the standalone dequantise (same values, one more launch)
  auto then_dequantize = [&](int rc) {
    if (rc != I8IE_OK || deq_out == nullptr) return rc;
    return i8ie_dequantize_rows_u8_f32(y, deq_out, m, n, ldy, sc, zp_out, stream);
  };
  I8IE_REQUIRE(impl != 3, "fc_u8: shape not eligible for the head kernel (needs n_pad == ldy == 16)");
  if (impl != 1 && eligible) {
    int bn, splits, kb_per, cluster;
    tc_fc_config(m, ldy, k, &bn, &splits, &kb_per, &cluster);
    CUtensorMap tmA, tmB;
    int rc = g_fc_maps.get(x, m, k, ldx, 1, &tmA, [&](CUtensorMap* mp) { return tc_encode_act_map_rows(mp, x, m, k, ldx); });
    if (rc != I8IE_OK) return rc;
    rc = g_fc_maps.get(w, n_pad, ldw, bn, 2, &tmB,
                       [&](CUtensorMap* mp) { return tc_encode_weight_map(mp, w, n_pad, ldw, 128, bn); });
    if (rc != I8IE_OK) return rc;
    static const bool no_tiled = std::getenv("I8IE_NO_FC_TILED") != nullptr;
    const int8_t* wt = (no_tiled || n_pad % bn != 0) ? nullptr : tiled_copy_of(w, n_pad, ldw);   // whole N tiles only
    return then_dequantize(launch_tc_fc(m, n, k, ldy, tmA, tmB, bn, splits, kb_per, y, ep, (cudaStream_t)stream, wt, ldw, cluster));
  }
  GemmGeom g;
  g.n = m; g.h = 1; g.w = 1; g.cp = ldx;
  g.kh = 1; g.kw = 1; g.stride = 1; g.pad = 0;
  g.oh = 1; g.ow = 1; g.M = m; g.N = n; g.n_pad = n_pad; g.ldw = ldw; g.out_cp = ldy;
  return then_dequantize(launch_simt_igemm(g, x, w, y, ep, 0, (cudaStream_t)stream));
}

int i8ie_fc_u8(const uint8_t* x, int ldx, const int8_t* w, int ldw, int n_pad, uint8_t* y, int ldy,
               int m, int n, int k, const int32_t* oc, const float* bias_f, float sa, float sb, float sc,
               int zp_out, int flags, int32_t* acc_out, int impl, void* stream) {
  return fc_u8(x, ldx, w, ldw, n_pad, y, ldy, m, n, k, oc, bias_f, sa, sb, nullptr, 0.f, 0.f, sc, zp_out, flags, acc_out,
               impl, stream);
}

int64_t i8ie_fc_weight_tiled_bytes(int n_pad, int ldw) {
  if (n_pad <= 0 || ldw <= 0 || n_pad % 128 != 0 || ldw % 128 != 0) return 0;   // shape has no tiled form
  return (int64_t)n_pad * ldw;
}

int i8ie_fc_weight_tiled_attach(const int8_t* w, int n_pad, int ldw, int8_t* w_tiled, void* stream) {
  I8IE_REQUIRE(w && w_tiled && i8ie_fc_weight_tiled_bytes(n_pad, ldw) > 0, "fc_weight_tiled_attach: bad arguments");
  int rc = tc_fc_tile_weights(w, n_pad, ldw, w_tiled, (cudaStream_t)stream);
  if (rc != I8IE_OK) return rc;
  // The fc kernels request their first weight stages BEFORE griddepcontrol.wait (the weights do not depend on the
  // previous kernel of a forward) — so the copy must be complete, not merely ordered on the stream, before the pair
  // is registered. Once per model, never inside a capture.
  cudaStreamCaptureStatus cst = cudaStreamCaptureStatusNone;
  if (cudaStreamIsCapturing((cudaStream_t)stream, &cst) == cudaSuccess && cst == cudaStreamCaptureStatusNone)
    I8IE_CUDA_OK(cudaStreamSynchronize((cudaStream_t)stream));
  else
    return I8IE_OK;   // capturing: leave the pair unregistered (the K-major weights are used)
  std::lock_guard<std::mutex> lk(g_tiled_mu);
  for (auto& e : g_tiled)
    if (e.w == w) { e = TiledWeight{w, w_tiled, n_pad, ldw}; return I8IE_OK; }
  g_tiled.push_back(TiledWeight{w, w_tiled, n_pad, ldw});
  return I8IE_OK;
}

int i8ie_fc_weight_tiled_detach(const int8_t* w) {
  std::lock_guard<std::mutex> lk(g_tiled_mu);
  for (size_t i = 0; i < g_tiled.size(); ++i)
    if (g_tiled[i].w == w) { g_tiled.erase(g_tiled.begin() + (long)i); break; }
  return I8IE_OK;
}

int i8ie_fc_u8_deq(const uint8_t* x, int ldx, const int8_t* w, int ldw, int n_pad, uint8_t* y, int ldy,
                   int m, int n, int k, const int32_t* oc, const float* bias_f, float sa, float sb,
                   const float* sb_vec, float sb_min, float sb_max, float sc, int zp_out, int flags, int impl,
                   float* deq_out, void* stream) {
  I8IE_REQUIRE(deq_out != nullptr, "fc_u8_deq: null output");
  I8IE_REQUIRE(sb_vec == nullptr || (sb_min > 0.f && sb_max >= sb_min), "fc_u8_deq: bad per-channel scales");
  return fc_u8(x, ldx, w, ldw, n_pad, y, ldy, m, n, k, oc, bias_f, sa, sb_vec ? sb_min : sb, sb_vec, sb_min, sb_max, sc,
               zp_out, flags, nullptr, impl, stream, deq_out);
}

int i8ie_fc_u8_pc(const uint8_t* x, int ldx, const int8_t* w, int ldw, int n_pad, uint8_t* y, int ldy,
                  int m, int n, int k, const int32_t* oc, const float* bias_f, float sa, const float* sb_vec,
                  float sb_min, float sb_max, float sc, int zp_out, int flags, int32_t* acc_out, int impl,
                  void* stream) {
  I8IE_REQUIRE(sb_vec != nullptr && sb_min > 0.f && sb_max >= sb_min, "fc_u8_pc: bad per-channel scales");
  return fc_u8(x, ldx, w, ldw, n_pad, y, ldy, m, n, k, oc, bias_f, sa, sb_min, sb_vec, sb_min, sb_max, sc, zp_out, flags,
               acc_out, impl, stream);
}

// Debug hook: first protocol error (timeout) recorded by a tensor-core kernel, 0 if none.
// Synchronises the device. Not part of the reference-facing surface.
int i8ie_debug_tc_error(int reset) {
  if (cudaDeviceSynchronize() != cudaSuccess) {
    set_error("debug_tc_error: %s", cudaGetErrorString(cudaGetLastError()));
    return I8IE_ECUDA;
  }
  int v = 0;
  int rc = tc_read_error(&v, reset != 0);
  return rc != I8IE_OK ? rc : v;
}

// Non-synchronising check used by the product paths (numpy(), dequantise read-back, bench, smoke):
// first protocol error mirrored into host-mapped memory on the current device (0 = none).
// Meaningful once the work in question has been synchronised (e.g. right after a D2H copy).
int i8ie_tc_error_poll(int reset) { return tc_error_poll(reset != 0); }

}  // extern "C"
