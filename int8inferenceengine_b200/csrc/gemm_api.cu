// C-ABI entry points for the GEMM-shaped ops: conv2d plan/run and fully-connected.
#include <new>

#include "gemm_api.cuh"

using namespace i8ie;

struct i8ie_conv_plan {
  GemmGeom g;
  const int8_t* w_packed;
  int impl;  // 1 = SIMT dp4a, 2 = tcgen05
};

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
  if (impl == 2) {
    set_error("conv2d_plan_create: tcgen05 path not available for this geometry");
    delete p;
    return nullptr;
  }
  p->impl = 1;
  return p;
}

void i8ie_conv2d_plan_destroy(i8ie_conv_plan* plan) { delete plan; }

int i8ie_conv2d_plan_impl(const i8ie_conv_plan* plan) { return plan ? plan->impl : 0; }

int i8ie_conv2d_u8(i8ie_conv_plan* plan, const uint8_t* x, uint8_t* y, const int32_t* oc, float sa,
                   float sb, float sc, int zp_in, int zp_out, int flags, int32_t* acc_out, void* stream) {
  I8IE_REQUIRE(plan && x && y && oc, "conv2d_u8: null argument");
  I8IE_REQUIRE(zp_in >= 0 && zp_in <= 255 && zp_out >= 0 && zp_out <= 255, "conv2d_u8: zero point out of range");
  EpiParams ep{oc, nullptr, sa, sb, sc, zp_out, (flags & I8IE_EPI_RELU) ? 1 : 0, acc_out};
  return launch_simt_igemm(plan->g, x, plan->w_packed, y, ep, zp_in, (cudaStream_t)stream);
}

int i8ie_fc_u8(const uint8_t* x, int ldx, const int8_t* w, int ldw, int n_pad, uint8_t* y, int ldy,
               int m, int n, int k, const int32_t* oc, const float* bias_f, float sa, float sb, float sc,
               int zp_out, int flags, int32_t* acc_out, int impl, void* stream) {
  I8IE_REQUIRE(x && w && y && oc && bias_f, "fc_u8: null argument");
  I8IE_REQUIRE(m > 0 && n > 0 && k > 0 && ldx % 16 == 0 && ldw % 16 == 0 && ldy % 16 == 0 && ldx >= k &&
                   ldw >= ldx && ldy >= n && n_pad >= n,
               "fc_u8: bad shape/pitch (m=%d n=%d k=%d ldx=%d ldw=%d ldy=%d n_pad=%d)", m, n, k, ldx, ldw, ldy, n_pad);
  I8IE_REQUIRE(zp_out >= 0 && zp_out <= 255, "fc_u8: zero point out of range");
  I8IE_REQUIRE(impl != 2, "fc_u8: tcgen05 path not available for this shape");
  GemmGeom g;
  g.n = m; g.h = 1; g.w = 1; g.cp = ldx;
  g.kh = 1; g.kw = 1; g.stride = 1; g.pad = 0;
  g.oh = 1; g.ow = 1; g.M = m; g.N = n; g.n_pad = n_pad; g.ldw = ldw; g.out_cp = ldy;
  EpiParams ep{oc, bias_f, sa, sb, sc, zp_out, (flags & I8IE_EPI_RELU) ? 1 : 0, acc_out};
  return launch_simt_igemm(g, x, w, y, ep, 0, (cudaStream_t)stream);
}

}  // extern "C"
