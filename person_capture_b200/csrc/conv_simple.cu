// CUDA-core direct convolution over the same P-layout / packed weights as conv_tc.cu.
// Validation kernel only (pcb_set_conv_impl(ctx, 1) in tests): it lets the parity suite tell a
// tensor-core descriptor bug from a graph/weight bug.  Never selected by the product path.
#include "pcb_common.cuh"

namespace {

struct SimpleParams {
  const __half* in;
  const __half* w;
  const float* scale;
  const float* bias;
  const float* slope;
  const __half* residual;
  const float* residual_f32;
  __half* out;
  float* out_s32;
  __half* out2;
  const float* scale2;
  const float* bias2;
  int cp_out2;
  float* out_f32;
  int n, h, w_, cp_in;       // input logical dims
  int ho, wo, cp_out, res_cp;
  int cin_eff;               // channels to reduce over (<= cp_in)
  int cout_store;            // channels to write
  int taps, cin_w, stride, act, dense, out_f32_stride, out_f32_cols;
  int iplo, ipad, oplo, opad;   // input / output tensor padding
};

__global__ void conv_simple_kernel(const SimpleParams p) {
  const long long total = p.dense ? (long long)p.n * p.cout_store : (long long)p.n * p.ho * p.wo * p.cout_store;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total;
       idx += (long long)gridDim.x * blockDim.x) {
    const int co = (int)(idx % p.cout_store);
    long long pix = idx / p.cout_store;
    float acc = 0.f;
    long long orow;
    if (p.dense) {
      const __half* a = p.in + pix * p.cp_in;
      const __half* wr = p.w + (long long)co * p.cin_w;
      for (int ci = 0; ci < p.cin_eff; ++ci) acc = fmaf(__half2float(a[ci]), __half2float(wr[ci]), acc);
      orow = pix;
    } else {
      const int xo = (int)(pix % p.wo);
      const int yo = (int)((pix / p.wo) % p.ho);
      const int img = (int)(pix / ((long long)p.wo * p.ho));
      const int yc = yo * p.stride, xc = xo * p.stride;  // centre in image coords
      for (int t = 0; t < p.taps; ++t) {
        const int dy = p.taps == 9 ? t / 3 - 1 : 0, dx = p.taps == 9 ? t % 3 - 1 : 0;
        if (yc + dy < 0 || yc + dy >= p.h || xc + dx < 0 || xc + dx >= p.w_) continue;   // zero padding
        const __half* a = p.in + pcb_prow_l(img, yc + dy, xc + dx, p.h, p.w_, p.iplo, p.ipad) * p.cp_in;
        const __half* wr = p.w + ((long long)co * p.taps + t) * p.cin_w;
        for (int ci = 0; ci < p.cin_eff; ++ci) acc = fmaf(__half2float(a[ci]), __half2float(wr[ci]), acc);
      }
      orow = pcb_prow_l(img, yo, xo, p.ho, p.wo, p.oplo, p.opad);
    }
    float y = fmaf(acc, p.scale[co], p.bias[co]);
    if (p.out_f32) {
      if (co < p.out_f32_cols) p.out_f32[orow * p.out_f32_stride + co] = y;
      continue;
    }
    if (p.residual_f32) y += p.residual_f32[orow * p.res_cp + co];
    else if (p.residual) y += __half2float(p.residual[orow * p.res_cp + co]);
    if (p.act == PCB_ACT_RELU) y = fmaxf(y, 0.f);
    else if (p.act == PCB_ACT_PRELU) y = y >= 0.f ? y : y * p.slope[co];
    if (p.out_s32) p.out_s32[orow * p.cp_out + co] = y;
    else p.out[orow * p.cp_out + co] = __float2half_rn(y);
    if (p.out2) p.out2[orow * p.cp_out2 + co] = __float2half_rn(fmaf(y, p.scale2[co], p.bias2[co]));
  }
}

}  // namespace

int pcb_conv_simple(pcb_ctx* c, const ConvArgs& a) {
  const PTensor& in = *a.in;
  const ConvWeights& w = *a.w;
  SimpleParams p{};
  p.in = in.data;
  p.w = w.w;
  p.scale = w.scale;
  p.bias = w.bias;
  p.slope = w.slope;
  p.n = in.n;
  p.h = in.h;
  p.w_ = in.w;
  p.cp_in = in.cp;
  p.iplo = in.pad_lo; p.ipad = in.pad;
  p.oplo = a.out ? a.out->pad_lo : kPadLo; p.opad = a.out ? a.out->pad : kPad;
  p.cin_eff = in.cp < w.cin_w ? in.cp : w.cin_w;
  p.taps = w.taps;
  p.cin_w = w.cin_w;
  p.stride = in.dense ? 1 : a.stride;
  p.act = a.act;
  p.dense = in.dense ? 1 : 0;
  if (in.dense) p.cin_w = w.cin_w;
  if (a.out_f32) {
    p.out_f32 = a.out_f32;
    p.out_f32_stride = a.out_f32_stride;
    p.out_f32_cols = w.cout;
    p.cout_store = w.cout;
    p.ho = p.wo = 1;
  } else {
    if (a.out->f32) p.out_s32 = (float*)a.out->data;
    else p.out = a.out->data;
    p.ho = a.out->h;
    p.wo = a.out->w;
    p.cp_out = a.out->cp;
    p.cout_store = a.out->cp;
    if (a.residual) {
      if (a.residual->f32) p.residual_f32 = (const float*)a.residual->data;
      else p.residual = a.residual->data;
      p.res_cp = a.residual->cp;
    }
    if (a.out2) {
      p.out2 = a.out2->data;
      p.cp_out2 = a.out2->cp;
      p.scale2 = a.scale2;
      p.bias2 = a.bias2;
    }
  }
  const long long total = p.dense ? (long long)p.n * p.cout_store : (long long)p.n * p.ho * p.wo * p.cout_store;
  long long blocks = (total + 255) / 256;
  if (blocks > 148LL * 32) blocks = 148LL * 32;
  if (blocks < 1) blocks = 1;
  conv_simple_kernel<<<(int)blocks, 256, 0, c->stream>>>(p);
  PCB_LAUNCH_CHECK(c, "conv_simple_kernel");
  return PCB_OK;
}
