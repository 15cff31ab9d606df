// Elementwise / pooling ops of the two graphs over the P-layout (HBM-bound, 16-byte vector
// accesses, 8 channels per thread).  They replace the BatchNorm / MaxPool / AveragePool /
// Resize(nearest)+Add / Flatten nodes that ONNX Runtime executes for the reference
// (face_embedder.py:1102-1107, 1369).  Only image pixels are written (zero pad invariant).
#include "pcb_common.cuh"

namespace {

struct Geo {
  int n, h, w, cp;      // output tensor logical dims + channel stride
  int ih, iw, icp;      // input dims
  int plo, pad;         // output tensor padding (PTensor::pad_lo / pad)
  int iplo, ipad;       // input tensor padding
};

// output-tensor row / input-tensor row of pixel (y, x)
__device__ __forceinline__ long long orow(const Geo& g, int img, int y, int x) { return pcb_prow_l(img, y, x, g.h, g.w, g.plo, g.pad); }
__device__ __forceinline__ long long irow(const Geo& g, int img, int y, int x) { return pcb_prow_l(img, y, x, g.ih, g.iw, g.iplo, g.ipad); }

__device__ __forceinline__ void unpack8(const uint4& v, float (&f)[8]) {
  const __half2* h = (const __half2*)&v;
#pragma unroll
  for (int j = 0; j < 4; ++j) {
    float2 t = __half22float2(h[j]);
    f[2 * j] = t.x;
    f[2 * j + 1] = t.y;
  }
}
__device__ __forceinline__ uint4 pack8(const float (&f)[8]) {
  uint4 v;
  __half2* h = (__half2*)&v;
#pragma unroll
  for (int j = 0; j < 4; ++j) h[j] = __floats2half2_rn(f[2 * j], f[2 * j + 1]);
  return v;
}

// decode a flat work index into (img, y, x, c8) over the output interior
#define PCB_DECODE_IDX(idx, g)                              \
  const int c8 = (int)((idx) % ((g).cp / 8));               \
  long long _t = (idx) / ((g).cp / 8);                      \
  const int x = (int)(_t % (g).w);                          \
  _t /= (g).w;                                              \
  const int y = (int)(_t % (g).h);                          \
  const int img = (int)(_t / (g).h);

__global__ void affine_kernel(const __half* __restrict__ in, __half* __restrict__ out, const float* __restrict__ scale,
                              const float* __restrict__ bias, Geo g, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    PCB_DECODE_IDX(idx, g)
    float f[8];
    unpack8(*(const uint4*)(in + irow(g, img, y, x) * g.icp + c8 * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], __ldg(scale + c8 * 8 + j), __ldg(bias + c8 * 8 + j));
    *(uint4*)(out + orow(g, img, y, x) * g.cp + c8 * 8) = pack8(f);
  }
}

__global__ void affine_flatten_kernel(const __half* __restrict__ in, __half* __restrict__ out, const float* __restrict__ scale,
                                      const float* __restrict__ bias, Geo g, long long total) {
  // out: dense [n][h*w*cp] in (y, x, c) order
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    PCB_DECODE_IDX(idx, g)
    float f[8];
    unpack8(*(const uint4*)(in + irow(g, img, y, x) * g.icp + c8 * 8), f);
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = fmaf(f[j], __ldg(scale + c8 * 8 + j), __ldg(bias + c8 * 8 + j));
    *(uint4*)(out + (((long long)img * g.h + y) * g.w + x) * g.cp + c8 * 8) = pack8(f);
  }
}

__global__ void maxpool3s2_kernel(const __half* __restrict__ in, __half* __restrict__ out, Geo g, long long total) {
  // 3x3, stride 2, pad 1 (pad value -inf): window rows 2y-1..2y+1 of the input.  fp16 max is exact, so the window is reduced in
  // packed half2 registers (no fp32 round trip): 9 16-byte loads + 36 HMNMX2 per 8 channels -- the kernel is bound by L2 reads
  // (every input pixel is touched 2.25 times), not by arithmetic.
  const __half2 lowest = __half2half2(__ushort_as_half((unsigned short)0xFBFF));      // -65504
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    PCB_DECODE_IDX(idx, g)
    __half2 m[4] = {lowest, lowest, lowest, lowest};
#pragma unroll
    for (int dy = -1; dy <= 1; ++dy) {
      const int yy = 2 * y + dy;
      if (yy < 0 || yy >= g.ih) continue;
#pragma unroll
      for (int dx = -1; dx <= 1; ++dx) {
        const int xx = 2 * x + dx;
        if (xx < 0 || xx >= g.iw) continue;
        const uint4 v = __ldg((const uint4*)(in + irow(g, img, yy, xx) * g.icp + c8 * 8));
        const __half2* h = (const __half2*)&v;
#pragma unroll
        for (int j = 0; j < 4; ++j) m[j] = __hmax2(m[j], h[j]);
      }
    }
    *(uint4*)(out + orow(g, img, y, x) * g.cp + c8 * 8) = *(const uint4*)m;
  }
}

__global__ void avgpool2_kernel(const __half* __restrict__ in, __half* __restrict__ out, Geo g, long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    PCB_DECODE_IDX(idx, g)
    float s[8] = {0, 0, 0, 0, 0, 0, 0, 0};
#pragma unroll
    for (int dy = 0; dy < 2; ++dy)
#pragma unroll
      for (int dx = 0; dx < 2; ++dx) {
        float f[8];
        unpack8(*(const uint4*)(in + irow(g, img, 2 * y + dy, 2 * x + dx) * g.icp + c8 * 8), f);
#pragma unroll
        for (int j = 0; j < 8; ++j) s[j] += f[j];
      }
#pragma unroll
    for (int j = 0; j < 8; ++j) s[j] *= 0.25f;
    *(uint4*)(out + orow(g, img, y, x) * g.cp + c8 * 8) = pack8(s);
  }
}

// out = a + (up ? nearest_up2(b) : b)
__global__ void add_kernel(const __half* __restrict__ a, const __half* __restrict__ b, __half* __restrict__ out, Geo g, int up,
                           long long total) {
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    PCB_DECODE_IDX(idx, g)
    const long long r = orow(g, img, y, x);       // `a` has the output's geometry and layout
    const long long rb = up ? irow(g, img, y >> 1, x >> 1) : irow(g, img, y, x);
    float fa[8], fb[8];
    unpack8(*(const uint4*)(a + r * g.cp + c8 * 8), fa);
    unpack8(*(const uint4*)(b + rb * g.icp + c8 * 8), fb);
#pragma unroll
    for (int j = 0; j < 8; ++j) fa[j] += fb[j];
    *(uint4*)(out + r * g.cp + c8 * 8) = pack8(fa);
  }
}

inline int grid_for(long long total, pcb_ctx* c) {
  long long b = (total + 255) / 256;
  long long cap = (long long)c->num_sms * 16;
  if (b > cap) b = cap;
  return (int)(b < 1 ? 1 : b);
}

inline Geo geo(const PTensor& in, const PTensor& out) {
  Geo g;
  g.n = out.n; g.h = out.h; g.w = out.w; g.cp = out.cp;
  g.ih = in.h; g.iw = in.w; g.icp = in.cp;
  g.plo = out.pad_lo; g.pad = out.pad; g.iplo = in.pad_lo; g.ipad = in.pad;
  return g;
}

}  // namespace

int pcb_op_affine(pcb_ctx* c, const PTensor& in, const PTensor& out, const float* scale, const float* bias) {
  Geo g = geo(in, out);
  long long total = (long long)out.n * out.h * out.w * (out.cp / 8);
  affine_kernel<<<grid_for(total, c), 256, 0, c->stream>>>(in.data, out.data, scale, bias, g, total);
  PCB_LAUNCH_CHECK(c, "affine_kernel");
  return PCB_OK;
}

int pcb_op_affine_flatten(pcb_ctx* c, const PTensor& in, const PTensor& out, const float* scale, const float* bias) {
  Geo g = geo(in, in);
  long long total = (long long)in.n * in.h * in.w * (in.cp / 8);
  affine_flatten_kernel<<<grid_for(total, c), 256, 0, c->stream>>>(in.data, out.data, scale, bias, g, total);
  PCB_LAUNCH_CHECK(c, "affine_flatten_kernel");
  return PCB_OK;
}

int pcb_op_maxpool3s2(pcb_ctx* c, const PTensor& in, const PTensor& out) {
  Geo g = geo(in, out);
  long long total = (long long)out.n * out.h * out.w * (out.cp / 8);
  maxpool3s2_kernel<<<grid_for(total, c), 256, 0, c->stream>>>(in.data, out.data, g, total);
  PCB_LAUNCH_CHECK(c, "maxpool3s2_kernel");
  return PCB_OK;
}

int pcb_op_avgpool2(pcb_ctx* c, const PTensor& in, const PTensor& out) {
  Geo g = geo(in, out);
  long long total = (long long)out.n * out.h * out.w * (out.cp / 8);
  avgpool2_kernel<<<grid_for(total, c), 256, 0, c->stream>>>(in.data, out.data, g, total);
  PCB_LAUNCH_CHECK(c, "avgpool2_kernel");
  return PCB_OK;
}

int pcb_op_upsample_add(pcb_ctx* c, const PTensor& big, const PTensor& small, const PTensor& out) {
  Geo g = geo(small, out);
  long long total = (long long)out.n * out.h * out.w * (out.cp / 8);
  add_kernel<<<grid_for(total, c), 256, 0, c->stream>>>(big.data, small.data, out.data, g, 1, total);
  PCB_LAUNCH_CHECK(c, "upsample_add_kernel");
  return PCB_OK;
}

int pcb_op_add(pcb_ctx* c, const PTensor& a, const PTensor& b, const PTensor& out) {
  Geo g = geo(b, out);
  long long total = (long long)out.n * out.h * out.w * (out.cp / 8);
  add_kernel<<<grid_for(total, c), 256, 0, c->stream>>>(a.data, b.data, out.data, g, 0, total);
  PCB_LAUNCH_CHECK(c, "add_kernel");
  return PCB_OK;
}
