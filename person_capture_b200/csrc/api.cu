// C ABI (include/pcb200.h): context, graph loading (weight repacking for the implicit GEMM),
// plan executor, and the detect / embed entry points that chain K1 -> K2 -> K3 and
// chip-patch -> K2 on the context's stream.
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "pcb_common.cuh"

// preproc.cu / detect.cu internals
int pcb_letterbox_impl(pcb_ctx* c, const uint8_t* frames, int n, int h, int w, int S, int rot, int pad, __half* out,
                       uint8_t* det_img, double* det_scale_out);
int pcb_chip_patch_impl(pcb_ctx* c, const uint8_t* chips, int f, int with_flip, __half* out);
int pcb_decode_nms_impl(pcb_ctx* c, const float* h8, const float* h16, const float* h32, const float* reg_scale3,
                        const pcb_detect_args* a, float det_scale);

struct Model {
  std::vector<pcb_op> ops;
  std::vector<ConvWeights> conv;     // per op (unused entries empty)
  std::vector<float*> aff_scale, aff_bias;   // AFFINE ops, and scale2/bias2 of CONV ops with out2
  int n_tensors = 0;
  std::vector<int> outputs;
  float reg_scale[3] = {1.f, 1.f, 1.f};
  bool has_reg_scale = false;
  int small_pad_max = 0;   // maps up to this many pixels wide use the trailing-pad layout (0: ring everywhere)
  struct Run {
    int n = 0, h = 0, w = 0;
    std::vector<PTensor> t;
    float* fc_out = nullptr;   // [n][512] for ArcFace
  };
  std::map<std::vector<int>, Run> runs;
  Run* last = nullptr;
  unsigned long long generation = 0;   // bumped whenever an activation set is (re)allocated: captured graphs hold its pointers
};

int pcb_fail(pcb_ctx* c, int code, const char* what, cudaError_t e) {
  if (c) {
    char buf[512];
    if (e != cudaSuccess) snprintf(buf, sizeof buf, "%s: %s", what, cudaGetErrorString(e));
    else snprintf(buf, sizeof buf, "%s", what);
    c->last_error = buf;
  }
  return code;
}

void* pcb_dev_alloc(pcb_ctx* c, size_t bytes, bool zero) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 16) != cudaSuccess) {
    cudaGetLastError();
    return nullptr;
  }
  if (zero && cudaMemsetAsync(p, 0, bytes, c->stream) != cudaSuccess) {
    cudaFree(p);
    return nullptr;
  }
  c->allocs.push_back(p);
  return p;
}

extern "C" pcb_ctx* pcb_create(int device, void* cuda_stream) {
  if (cudaSetDevice(device) != cudaSuccess) return nullptr;
  pcb_ctx* c = new pcb_ctx();
  c->device = device;
  if (cuda_stream) {
    c->stream = (cudaStream_t)cuda_stream;
  } else {
    if (cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking) != cudaSuccess) { delete c; return nullptr; }
    c->own_stream = true;
  }
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) { delete c; return nullptr; }
  c->num_sms = prop.multiProcessorCount;
  if (prop.major != 10) {
    // sm_100a cubins only: refuse to run anywhere else instead of failing at the first launch
    fprintf(stderr, "libpcb200: device %d is sm_%d%d; this library is built for sm_100a (B200) only\n", device, prop.major, prop.minor);
    delete c;
    return nullptr;
  }
  c->d_err = (int*)pcb_dev_alloc(c, 256, true);
  if (!c->d_err || cudaMallocHost((void**)&c->h_err, sizeof(int)) != cudaSuccess) { delete c; return nullptr; }
  *c->h_err = 0;
  cudaStreamSynchronize(c->stream);
  return c;
}

extern "C" void pcb_destroy(pcb_ctx* c) {
  if (!c) return;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (void* p : c->allocs) cudaFree(p);
  for (cudaEvent_t e : c->ev_pool) cudaEventDestroy(e);
  if (c->h_err) cudaFreeHost(c->h_err);
  if (c->bank_stage) cudaFreeHost(c->bank_stage);
  if (c->live_sim_host) cudaFreeHost(c->live_sim_host);
  if (c->live_row_stage) cudaFreeHost(c->live_row_stage);
  if (c->bank_ev) cudaEventDestroy(c->bank_ev);
  for (auto& kv : c->embed_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
  for (int i = 0; i < 4; ++i) delete c->models[i];
  if (c->own_stream) cudaStreamDestroy(c->stream);
  delete c;
}

extern "C" const char* pcb_last_error(pcb_ctx* c) { return c ? c->last_error.c_str() : "null context"; }

extern "C" int pcb_sync(pcb_ctx* c) {
  PCB_ENTER(c);
  PCB_CUDA(c, cudaMemcpyAsync(c->h_err, c->d_err, sizeof(int), cudaMemcpyDeviceToHost, c->stream));
  PCB_CUDA(c, cudaStreamSynchronize(c->stream));
  if (*c->h_err != 0) {
    const int code = *c->h_err;
    *c->h_err = 0;
    cudaMemsetAsync(c->d_err, 0, sizeof(int), c->stream);
    cudaStreamSynchronize(c->stream);
    char buf[160];
    const char* what = code >= 100 && code < 200 ? "conv_tc watchdog: mbarrier wait timed out"
                     : code == 201             ? "decode: more than 8192 candidates above threshold in one frame"
                     : code == 202             ? "nms: detections exceed max_det"
                     : code == 301             ? "align: faces exceed max_faces"
                     : code == 302             ? "align: eye-roll scratch exhausted"
                                               : "device error";
    snprintf(buf, sizeof buf, "%s (device code %d)", what, code);
    return pcb_fail(c, PCB_ERR_KERNEL, buf);
  }
  return PCB_OK;
}

extern "C" void pcb_layout_pad(int* pad_lo, int* pad) {
  if (pad_lo) *pad_lo = kPadLo;
  if (pad) *pad = kPad;
}

extern "C" int pcb_set_conv_impl(pcb_ctx* c, int impl) {
  if (impl < 0 || impl > 2) return pcb_fail(c, PCB_ERR_ARG, "conv impl must be 0..2");
#ifndef PCB_VALIDATION_KERNEL
  if (impl == 1) return pcb_fail(c, PCB_ERR_ARG, "the CUDA-core validation kernel is not part of the product library (load libpcb200_val.so)");
#endif
  c->conv_impl = impl;
  return PCB_OK;
}
PcbConvTimer::PcbConvTimer(pcb_ctx* ctx, double flops, const char* desc) : c(ctx) {
  if (!c->profile) return;
  c->ev_desc.push_back(desc ? desc : "");
  while (c->ev_pool.size() < c->ev_used + 2) {
    cudaEvent_t e;
    if (cudaEventCreate(&e) != cudaSuccess) return;
    c->ev_pool.push_back(e);
  }
  cudaEvent_t e0 = c->ev_pool[c->ev_used];
  e1 = c->ev_pool[c->ev_used + 1];
  c->ev_used += 2;
  c->ev_flops.push_back(flops);
  cudaEventRecord(e0, c->stream);
}
PcbConvTimer::~PcbConvTimer() {
  if (e1) cudaEventRecord(e1, c->stream);
}

extern "C" int pcb_set_profile(pcb_ctx* c, int on) {
  c->profile = on != 0;
  return PCB_OK;
}

// Folds all recorded conv launches into the running totals and returns them (synchronises).
extern "C" int pcb_profile_read(pcb_ctx* c, double* conv_ms, double* conv_flops, long long* conv_launches, int reset) {
  PCB_ENTER(c);
  PCB_CUDA(c, cudaStreamSynchronize(c->stream));
  FILE* dump = nullptr;
  if (const char* path = getenv("PCB_PROFILE_DUMP")) dump = fopen(path, "a");
  for (size_t i = 0; i + 1 < c->ev_used; i += 2) {
    float ms = 0.f;
    if (cudaEventElapsedTime(&ms, c->ev_pool[i], c->ev_pool[i + 1]) == cudaSuccess) {
      c->prof_ms += ms;
      c->prof_flops += c->ev_flops[i / 2];
      c->prof_launches++;
      if (dump) fprintf(dump, "%s,%.6f,%.0f\n", c->ev_desc[i / 2].c_str(), ms, c->ev_flops[i / 2]);
    }
  }
  if (dump) fclose(dump);
  c->ev_used = 0;
  c->ev_flops.clear();
  c->ev_desc.clear();
  if (conv_ms) *conv_ms = c->prof_ms;
  if (conv_flops) *conv_flops = c->prof_flops;
  if (conv_launches) *conv_launches = c->prof_launches;
  if (reset) { c->prof_ms = 0.0; c->prof_flops = 0.0; c->prof_launches = 0; }
  return PCB_OK;
}

extern "C" long long pcb_launch_count(pcb_ctx* c) { return c->launches; }
extern "C" void pcb_reset_launch_count(pcb_ctx* c) { c->launches = 0; }

// ---------------------------------------------------------------------------------------
// graph loading
// ---------------------------------------------------------------------------------------
static float* upload_f32_padded(pcb_ctx* c, const float* src, int n, int npad, float fill) {
  std::vector<float> h(npad, 0.f);
  for (int i = 0; i < n; ++i) h[i] = src ? src[i] : fill;
  float* d = (float*)pcb_dev_alloc(c, (size_t)npad * sizeof(float), false);
  if (!d) return nullptr;
  if (cudaMemcpyAsync(d, h.data(), (size_t)npad * sizeof(float), cudaMemcpyHostToDevice, c->stream) != cudaSuccess) return nullptr;
  cudaStreamSynchronize(c->stream);
  return d;
}

extern "C" int pcb_model_load(pcb_ctx* c, int slot, const pcb_op* ops, int n_ops, int n_tensors, const void* blob_host,
                              size_t blob_bytes, const int32_t* outputs, int n_outputs, const float* reg_scale3) {
  PCB_ENTER(c);
  if (slot < 0 || slot >= 4 || !ops || n_ops <= 0 || !blob_host) return pcb_fail(c, PCB_ERR_ARG, "model_load: bad arguments");
  Model* m = new Model();
  m->ops.assign(ops, ops + n_ops);
  m->conv.resize(n_ops);
  m->aff_scale.assign(n_ops, nullptr);
  m->aff_bias.assign(n_ops, nullptr);
  m->n_tensors = n_tensors;
  m->outputs.assign(outputs, outputs + n_outputs);
  if (reg_scale3) { memcpy(m->reg_scale, reg_scale3, sizeof(float) * 3); m->has_reg_scale = true; }
  // iResNet: the 14x14 and 7x7 stages (half of the FLOPs) use the trailing-pad layout; PCB_SMALL_PAD_MAX=0 keeps the ring everywhere
  // (A/B), any other value moves the boundary.  SCRFD keeps the ring: its head maps are read by K3 with the ring geometry.
  if (slot == PCB_MODEL_ARCFACE) m->small_pad_max = getenv("PCB_SMALL_PAD_MAX") ? atoi(getenv("PCB_SMALL_PAD_MAX")) : 16;
  const uint8_t* blob = (const uint8_t*)blob_host;
  for (int i = 0; i < n_ops; ++i) {
    const pcb_op& op = m->ops[i];
    auto need = [&](int64_t off, size_t bytes) { return off >= 0 && (size_t)off + bytes <= blob_bytes; };
    if (op.kind == PCB_OP_CONV || op.kind == PCB_OP_FC) {
      ConvWeights& w = m->conv[i];
      const bool stem = (op.kind == PCB_OP_CONV && op.in0 == 0);
      const bool fc = op.kind == PCB_OP_FC;
      w.cin = op.cin; w.cout = op.cout; w.k = fc ? 1 : op.k;
      const int kk = w.k * w.k;
      if (stem && !(op.k == 3 && op.cin == 3)) { delete m; return pcb_fail(c, PCB_ERR_ARG, "model_load: stem must be 3x3 over 3 channels"); }
      w.taps = stem ? 1 : kk;
      const int cin_eff = stem ? 27 : op.cin;
      w.cin_w = pcb_round_up(cin_eff, 64);
      w.npad = op.cout <= 256 ? pcb_round_up(op.cout, 16) : pcb_round_up(op.cout, 256);
      w.n_tile = w.npad <= 256 ? w.npad : 256;
      w.rows_alloc = pcb_round_up(w.npad, 128);   // zero rows up to a multiple of 128 (TMA boxes of the weight map never run past the allocation)
      const size_t wbytes = (size_t)op.cout * op.cin * kk * sizeof(__half);
      if (!need(op.w_off, wbytes) || !need(op.scale_off, op.cout * 4) || !need(op.bias_off, op.cout * 4)) {
        delete m;
        return pcb_fail(c, PCB_ERR_ARG, "model_load: blob offsets out of range");
      }
      const __half* src = (const __half*)(blob + op.w_off);
      std::vector<__half> packed((size_t)w.rows_alloc * w.taps * w.cin_w, __float2half(0.f));
      for (int co = 0; co < op.cout; ++co)
        for (int ci = 0; ci < op.cin; ++ci)
          for (int t = 0; t < kk; ++t) {
            const __half v = src[((size_t)co * op.cin + ci) * kk + t];
            size_t dst;
            if (stem) dst = (size_t)co * w.cin_w + (size_t)t * 3 + ci;            // K index = (ky*3+kx)*3 + c
            else dst = ((size_t)co * w.taps + t) * w.cin_w + ci;
            packed[dst] = v;
          }
      w.w = (__half*)pcb_dev_alloc(c, packed.size() * sizeof(__half), false);
      if (!w.w) { delete m; return pcb_fail(c, PCB_ERR_CUDA, "model_load: weight alloc failed"); }
      PCB_CUDA(c, cudaMemcpyAsync(w.w, packed.data(), packed.size() * sizeof(__half), cudaMemcpyHostToDevice, c->stream));
      PCB_CUDA(c, cudaStreamSynchronize(c->stream));
      w.scale = upload_f32_padded(c, (const float*)(blob + op.scale_off), op.cout, w.rows_alloc, 0.f);
      w.bias = upload_f32_padded(c, (const float*)(blob + op.bias_off), op.cout, w.rows_alloc, 0.f);
      if (op.act == PCB_ACT_PRELU) {
        if (!need(op.slope_off, op.cout * 4)) { delete m; return pcb_fail(c, PCB_ERR_ARG, "model_load: slope offset"); }
        w.slope = upload_f32_padded(c, (const float*)(blob + op.slope_off), op.cout, w.rows_alloc, 0.f);
      }
      if (!w.scale || !w.bias) { delete m; return pcb_fail(c, PCB_ERR_CUDA, "model_load: param upload failed"); }
      if (op.kind == PCB_OP_CONV && op.out2 >= 0) {
        if (!need(op.scale2_off, op.cout * 4) || !need(op.bias2_off, op.cout * 4)) { delete m; return pcb_fail(c, PCB_ERR_ARG, "model_load: out2 offsets"); }
        m->aff_scale[i] = upload_f32_padded(c, (const float*)(blob + op.scale2_off), op.cout, w.rows_alloc, 0.f);
        m->aff_bias[i] = upload_f32_padded(c, (const float*)(blob + op.bias2_off), op.cout, w.rows_alloc, 0.f);
        if (!m->aff_scale[i] || !m->aff_bias[i]) { delete m; return pcb_fail(c, PCB_ERR_CUDA, "model_load: param upload failed"); }
      }
    } else if (op.kind == PCB_OP_AFFINE || op.kind == PCB_OP_AFFINE_FLATTEN) {
      if (!need(op.scale_off, op.cout * 4) || !need(op.bias_off, op.cout * 4)) { delete m; return pcb_fail(c, PCB_ERR_ARG, "model_load: affine offsets"); }
      const int cp = pcb_round_up(op.cout, 8);
      m->aff_scale[i] = upload_f32_padded(c, (const float*)(blob + op.scale_off), op.cout, cp, 0.f);
      m->aff_bias[i] = upload_f32_padded(c, (const float*)(blob + op.bias_off), op.cout, cp, 0.f);
      if (!m->aff_scale[i] || !m->aff_bias[i]) { delete m; return pcb_fail(c, PCB_ERR_CUDA, "model_load: param upload failed"); }
    }
  }
  if (slot == PCB_MODEL_ARCFACE) {            // captured small-call graphs hold pointers into the old model
    for (auto& kv : c->embed_graphs) if (kv.second.exec) cudaGraphExecDestroy(kv.second.exec);
    c->embed_graphs.clear();
  }
  if (Model* old = c->models[slot]) {
    // the replaced graph's weights, vectors and activation sets go back to the device now, not at pcb_destroy
    PCB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto& w : old->conv) { pcb_dev_free(c, w.w); pcb_dev_free(c, w.scale); pcb_dev_free(c, w.bias); pcb_dev_free(c, w.slope); }
    for (float* q : old->aff_scale) pcb_dev_free(c, q);
    for (float* q : old->aff_bias) pcb_dev_free(c, q);
    for (auto& kv : old->runs) {
      for (auto& t : kv.second.t) pcb_dev_free(c, t.data);
      pcb_dev_free(c, kv.second.fc_out);
    }
    delete old;
  }
  c->models[slot] = m;
  return PCB_OK;
}

// ---------------------------------------------------------------------------------------
// plan executor
// ---------------------------------------------------------------------------------------
static int alloc_tensor(pcb_ctx* c, PTensor& t) {
  t.data = (__half*)pcb_dev_alloc(c, t.bytes(), true);   // zero ring (and zero pad channels) once
  return t.data ? PCB_OK : pcb_fail(c, PCB_ERR_CUDA, "activation alloc failed (batch too large for HBM?)");
}

void pcb_dev_free(pcb_ctx* c, void* p) {
  if (!p) return;
  for (size_t i = 0; i < c->allocs.size(); ++i)
    if (c->allocs[i] == p) { c->allocs.erase(c->allocs.begin() + i); break; }
  cudaFree(p);
}

// Builds (or fetches) the activation set for input patch-tensor dims h x w with capacity >= n
// images.  Buffers are image-major, so a smaller batch simply uses a prefix of each buffer.
static int model_prepare(pcb_ctx* c, Model* m, int n, int h, int w, Model::Run** out) {
  std::vector<int> key = {h, w};
  auto it = m->runs.find(key);
  if (it != m->runs.end() && it->second.n >= n) {
    for (auto& t : it->second.t) if (t.data) t.n = n;
    *out = &it->second;
    return PCB_OK;
  }
  int old_cap = 0;
  if (it != m->runs.end()) {   // grow: release the smaller set first
    old_cap = it->second.n;
    PCB_CUDA(c, cudaStreamSynchronize(c->stream));
    for (auto& t : it->second.t) pcb_dev_free(c, t.data);
    pcb_dev_free(c, it->second.fc_out);
    if (m->last == &it->second) m->last = nullptr;
    m->runs.erase(it);
  }
  // geometric growth: a caller that ramps its batch up (first small ArcFace runs of a host-resident pre-scan) pays for at most
  // log2 reallocations, each of which synchronises the stream
  const int cap = pcb_round_up(n > 2 * old_cap ? n : 2 * old_cap, 8);
  Model::Run r;
  r.n = cap; r.h = h; r.w = w;
  r.t.resize(m->n_tensors);
  PTensor& in = r.t[0];
  in.n = cap; in.h = h; in.w = w; in.c = 27; in.cp = 32;
  int rc = alloc_tensor(c, in);
  if (rc) return rc;
  for (size_t i = 0; i < m->ops.size(); ++i) {
    const pcb_op& op = m->ops[i];
    if (op.kind == PCB_OP_FC) {
      r.fc_out = (float*)pcb_dev_alloc(c, (size_t)cap * op.cout * sizeof(float), true);
      if (!r.fc_out) return pcb_fail(c, PCB_ERR_CUDA, "fc output alloc failed");
      continue;
    }
    const PTensor& a = r.t[op.in0];
    PTensor& o = r.t[op.out];
    if (o.data) return pcb_fail(c, PCB_ERR_ARG, "graph writes a tensor twice");
    o.n = cap;
    auto layout_for = [&](PTensor& t) {
      if (!t.dense && t.w <= m->small_pad_max && t.h <= m->small_pad_max) { t.pad_lo = 0; t.pad = 1; }
      else { t.pad_lo = kPadLo; t.pad = kPad; }
    };
    switch (op.kind) {
      case PCB_OP_CONV: {
        const bool stem = op.in0 == 0;
        const int s = stem ? 1 : op.stride;
        o.h = a.h / s; o.w = a.w / s; o.c = op.cout; o.cp = pcb_round_up(op.cout, 8);
        o.f32 = (op.flags & PCB_OPF_OUT_F32) != 0;
        layout_for(o);
        if (op.out2 >= 0) {
          PTensor& o2 = r.t[op.out2];
          if (o2.data) return pcb_fail(c, PCB_ERR_ARG, "graph writes a tensor twice");
          o2 = o;
          o2.f32 = false;
          o2.data = nullptr;
          rc = alloc_tensor(c, o2);
          if (rc) return rc;
        }
        break;
      }
      case PCB_OP_AFFINE: o.h = a.h; o.w = a.w; o.c = a.c; o.cp = a.cp; layout_for(o); break;
      case PCB_OP_MAXPOOL3S2:
      case PCB_OP_AVGPOOL2: o.h = a.h / 2; o.w = a.w / 2; o.c = a.c; o.cp = a.cp; layout_for(o); break;
      case PCB_OP_UPSAMPLE_ADD:
      case PCB_OP_ADD: o.h = a.h; o.w = a.w; o.c = a.c; o.cp = a.cp; layout_for(o); break;
      case PCB_OP_AFFINE_FLATTEN: o.dense = true; o.h = 1; o.w = 1; o.c = a.h * a.w * a.cp; o.cp = o.c; break;
      default: return pcb_fail(c, PCB_ERR_ARG, "unknown op kind");
    }
    rc = alloc_tensor(c, o);
    if (rc) return rc;
  }
  m->generation++;
  auto ins = m->runs.emplace(key, std::move(r));
  for (auto& t : ins.first->second.t) if (t.data) t.n = n;
  *out = &ins.first->second;
  return PCB_OK;
}

static int pick_n_tile(const pcb_ctx* c, const ConvWeights& w, size_t rows) {
  if (w.npad <= 128) return w.npad;
  const long long m_tiles = (long long)((rows + 127) / 128);
  const int cands[3] = {256, 128, 64};
  static const int nt_max = getenv("PCB_CONV_NTILE_MAX") ? atoi(getenv("PCB_CONV_NTILE_MAX")) : 256;   // A/B knob
  for (int i = 0; i < 3; ++i) {
    const int nt = cands[i];
    if (nt > w.npad || w.npad % nt || nt > nt_max) continue;
    if (m_tiles * (w.npad / nt) >= c->num_sms || nt == 64) return nt;
  }
  return w.npad <= 256 ? w.npad : 128;
}

static int model_run(pcb_ctx* c, Model* m, Model::Run* r) {
  for (size_t i = 0; i < m->ops.size(); ++i) {
    const pcb_op& op = m->ops[i];
    int rc = PCB_OK;
    switch (op.kind) {
      case PCB_OP_CONV:
      case PCB_OP_FC: {
        ConvWeights w = m->conv[i];
        ConvArgs a{};
        a.in = &r->t[op.in0];
        a.w = &w;
        a.act = op.act;
        a.stride = (op.kind == PCB_OP_CONV && op.in0 != 0) ? op.stride : 1;
        if (op.kind == PCB_OP_FC) {
          a.out_f32 = r->fc_out;
          a.out_f32_stride = op.cout;
        } else {
          a.out = &r->t[op.out];
          a.residual = op.in1 >= 0 ? &r->t[op.in1] : nullptr;
          if (op.out2 >= 0) {
            a.out2 = &r->t[op.out2];
            a.scale2 = m->aff_scale[i];
            a.bias2 = m->aff_bias[i];
          }
        }
        w.n_tile = pick_n_tile(c, w, a.in->rows());
#ifdef PCB_VALIDATION_KERNEL
        rc = c->conv_impl == 0 ? pcb_conv_tc2(c, a) : c->conv_impl == 2 ? pcb_conv_tc(c, a) : pcb_conv_simple(c, a);
#else
        rc = c->conv_impl == 2 ? pcb_conv_tc(c, a) : pcb_conv_tc2(c, a);
#endif
        break;
      }
      case PCB_OP_AFFINE: rc = pcb_op_affine(c, r->t[op.in0], r->t[op.out], m->aff_scale[i], m->aff_bias[i]); break;
      case PCB_OP_AFFINE_FLATTEN: rc = pcb_op_affine_flatten(c, r->t[op.in0], r->t[op.out], m->aff_scale[i], m->aff_bias[i]); break;
      case PCB_OP_MAXPOOL3S2: rc = pcb_op_maxpool3s2(c, r->t[op.in0], r->t[op.out]); break;
      case PCB_OP_AVGPOOL2: rc = pcb_op_avgpool2(c, r->t[op.in0], r->t[op.out]); break;
      case PCB_OP_UPSAMPLE_ADD: rc = pcb_op_upsample_add(c, r->t[op.in0], r->t[op.in1], r->t[op.out]); break;
      case PCB_OP_ADD: rc = pcb_op_add(c, r->t[op.in0], r->t[op.in1], r->t[op.out]); break;
      default: rc = pcb_fail(c, PCB_ERR_ARG, "unknown op kind");
    }
    if (rc) return rc;
  }
  m->last = r;
  return PCB_OK;
}

__global__ void gather_nchw_kernel(const __half* __restrict__ in, float* __restrict__ out, int n, int c, int h, int w, int cp, int dense,
                                   int is_f32, int pad_lo, int pad) {
  const long long total = (long long)n * c * h * w;
  for (long long idx = blockIdx.x * (long long)blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int x = (int)(idx % w);
    const int y = (int)((idx / w) % h);
    const int ch = (int)((idx / ((long long)w * h)) % c);
    const int img = (int)(idx / ((long long)w * h * c));
    const long long row = dense ? img : pcb_prow_l(img, y, x, h, w, pad_lo, pad);
    out[idx] = is_f32 ? ((const float*)in)[row * cp + ch] : __half2float(in[row * cp + ch]);
  }
}

extern "C" int pcb_model_get_tensor(pcb_ctx* c, int slot, int tid, float* out_host, int* n, int* ch, int* h, int* w) {
  PCB_ENTER(c);
  if (slot < 0 || slot >= 4 || !c->models[slot] || !c->models[slot]->last) return pcb_fail(c, PCB_ERR_STATE, "get_tensor: model has not run");
  Model::Run* r = c->models[slot]->last;
  if (tid < 0 || tid >= (int)r->t.size() || !r->t[tid].data) return pcb_fail(c, PCB_ERR_ARG, "get_tensor: bad tensor id");
  const PTensor& t = r->t[tid];
  const int C = t.dense ? t.cp : t.c;
  if (n) *n = t.n;
  if (ch) *ch = C;
  if (h) *h = t.dense ? 1 : t.h;
  if (w) *w = t.dense ? 1 : t.w;
  if (!out_host) return PCB_OK;
  const size_t total = (size_t)t.n * C * (t.dense ? 1 : t.h * t.w);
  float* d = nullptr;
  PCB_CUDA(c, cudaMalloc(&d, total * sizeof(float)));
  gather_nchw_kernel<<<1024, 256, 0, c->stream>>>(t.data, d, t.n, C, t.dense ? 1 : t.h, t.dense ? 1 : t.w, t.cp, t.dense ? 1 : 0, t.f32 ? 1 : 0, t.pad_lo, t.pad);
  cudaError_t e = cudaMemcpyAsync(out_host, d, total * sizeof(float), cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  if (e != cudaSuccess) return pcb_fail(c, PCB_ERR_CUDA, "get_tensor copy", e);
  return PCB_OK;
}

// ---------------------------------------------------------------------------------------
// detect / letterbox / decode / embed
// ---------------------------------------------------------------------------------------
extern "C" int pcb_letterbox(pcb_ctx* c, const uint8_t* frames_dev, int n, int h, int w, int S, int rot_deg, int pad_replicate,
                             void* out_dev, uint8_t* det_img_dev) {
  PCB_ENTER(c);
  return pcb_letterbox_impl(c, frames_dev, n, h, w, S, rot_deg, pad_replicate, (__half*)out_dev, det_img_dev, nullptr);
}

extern "C" int pcb_decode_nms(pcb_ctx* c, const void* h8, const void* h16, const void* h32, const float* reg_scale3_host,
                              const pcb_detect_args* a, float det_scale) {
  PCB_ENTER(c);
  return pcb_decode_nms_impl(c, (const float*)h8, (const float*)h16, (const float*)h32, reg_scale3_host, a, det_scale);
}

extern "C" int pcb_detect(pcb_ctx* c, const pcb_detect_args* a) {
  PCB_ENTER(c);
  if (!a || !a->frames_dev || a->n <= 0) return pcb_fail(c, PCB_ERR_ARG, "detect: bad arguments");
  Model* m = c->models[PCB_MODEL_SCRFD];
  if (!m) return pcb_fail(c, PCB_ERR_STATE, "detect: no SCRFD graph loaded");
  if (m->outputs.size() != 3) return pcb_fail(c, PCB_ERR_STATE, "detect: SCRFD graph must have 3 outputs");
  if (a->S % 32 || a->S < 64) return pcb_fail(c, PCB_ERR_ARG, "detect: S must be a multiple of 32");
  Model::Run* r = nullptr;
  int rc = model_prepare(c, m, a->n, a->S / 2, a->S / 2, &r);
  if (rc) return rc;
  double det_scale = 1.0;
  rc = pcb_letterbox_impl(c, a->frames_dev, a->n, a->h, a->w, a->S, a->rot_deg, a->pad_replicate, r->t[0].data, nullptr, &det_scale);
  if (rc) return rc;
  rc = model_run(c, m, r);
  if (rc) return rc;
  for (int i = 0; i < 3; ++i)
    if (!r->t[m->outputs[i]].f32) return pcb_fail(c, PCB_ERR_STATE, "detect: SCRFD head outputs must be fp32 (PCB_OPF_OUT_F32)");
  return pcb_decode_nms_impl(c, (const float*)r->t[m->outputs[0]].data, (const float*)r->t[m->outputs[1]].data,
                             (const float*)r->t[m->outputs[2]].data,
                             m->has_reg_scale ? m->reg_scale : nullptr, a, (float)det_scale);
}

extern "C" int pcb_embed(pcb_ctx* c, const uint8_t* chips_dev, int f, float* emb_dev, float* emb_flip_dev) {
  PCB_ENTER(c);
  if (f < 0 || (f > 0 && (!chips_dev || (!emb_dev && !emb_flip_dev)))) return pcb_fail(c, PCB_ERR_ARG, "embed: bad arguments");
  Model* m = c->models[PCB_MODEL_ARCFACE];
  if (!m) return pcb_fail(c, PCB_ERR_STATE, "embed: no ArcFace graph loaded");
  // 0: e(x) only, 1: e(x) and e(flip x), 2: e(flip x) only
  const int mode = emb_dev ? (emb_flip_dev ? 1 : 0) : 2;
  // images per graph run: whole waves of 256-row tiles over 74 CTA pairs in the 14x14 stage, which is half of the FLOPs.
  // Trailing-pad layout (225 rows per 14x14 image): 504 images = 443 pair tiles = 5.99 waves (28x28: 11.97, 56x56: 44.8);
  // ring layout (256 rows): 444 images = 6.0 waves.  Activation memory ~12.5 GB for iResNet-100.
  static const int chunk_env = getenv("PCB_EMBED_CHUNK") ? atoi(getenv("PCB_EMBED_CHUNK")) : 0;
  const int run = m->small_pad_max >= 14 ? 504 : 444;
  const int chunk = chunk_env > 0 ? chunk_env : (mode == 1 ? run / 2 : run);
  // 1..8 faces (the lock-face ROI path embeds one face + its mirror per frame): the ~110 launches of the graph are
  // launch-latency bound at this size, so they are captured once per (mode, faces) into a CUDA graph that reads a fixed
  // staging buffer and is replayed with one launch.  The first call of a key runs eagerly (one-time attribute / allocation work
  // must not happen inside a capture), the second captures, later ones replay.  PCB_EMBED_GRAPH=0 disables.
  static const int graph_on = getenv("PCB_EMBED_GRAPH") ? atoi(getenv("PCB_EMBED_GRAPH")) : 1;
  if (graph_on && f >= 1 && f <= 8 && !c->profile && c->conv_impl == 0) {
    const int imgs = mode == 1 ? 2 * f : f;
    Model::Run* r = nullptr;
    int rc = model_prepare(c, m, imgs, PCB_CHIP, PCB_CHIP, &r);
    if (rc) return rc;
    const size_t chip_bytes = (size_t)PCB_CHIP * PCB_CHIP * 3;
    if (!c->embed_stage) {
      c->embed_stage = (uint8_t*)pcb_dev_alloc(c, 8 * chip_bytes, true);
      if (!c->embed_stage) return pcb_fail(c, PCB_ERR_CUDA, "embed: staging alloc failed");
    }
    pcb_ctx::EmbedGraph& g = c->embed_graphs[mode * 64 + f];
    if (g.exec && g.gen != m->generation) {      // the activation set moved: the captured pointers are stale
      cudaGraphExecDestroy(g.exec);
      g.exec = nullptr;
    }
    PCB_CUDA(c, cudaMemcpyAsync(c->embed_stage, chips_dev, f * chip_bytes, cudaMemcpyDeviceToDevice, c->stream));
    if (g.exec) {
      PCB_CUDA(c, cudaGraphLaunch(g.exec, c->stream));
      c->launches += 1;
    } else if (!g.warmed) {
      rc = pcb_chip_patch_impl(c, c->embed_stage, f, mode, r->t[0].data);
      if (!rc) rc = model_run(c, m, r);
      if (rc) return rc;
      g.warmed = true;
    } else {
      cudaGraph_t graph = nullptr;
      PCB_CUDA(c, cudaStreamBeginCapture(c->stream, cudaStreamCaptureModeThreadLocal));
      rc = pcb_chip_patch_impl(c, c->embed_stage, f, mode, r->t[0].data);
      if (!rc) rc = model_run(c, m, r);
      cudaError_t ce = cudaStreamEndCapture(c->stream, &graph);
      if (rc || ce != cudaSuccess || !graph) {
        if (graph) cudaGraphDestroy(graph);
        return rc ? rc : pcb_fail(c, PCB_ERR_CUDA, "embed: graph capture failed", ce);
      }
      ce = cudaGraphInstantiate(&g.exec, graph, 0);
      cudaGraphDestroy(graph);
      if (ce != cudaSuccess) { g.exec = nullptr; return pcb_fail(c, PCB_ERR_CUDA, "embed: graph instantiate failed", ce); }
      g.gen = m->generation;
      PCB_CUDA(c, cudaGraphLaunch(g.exec, c->stream));
    }
    m->last = r;
    float* first = mode == 2 ? emb_flip_dev : emb_dev;
    PCB_CUDA(c, cudaMemcpyAsync(first, r->fc_out, (size_t)f * PCB_FEAT_DIM * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
    if (mode == 1)
      PCB_CUDA(c, cudaMemcpyAsync(emb_flip_dev, r->fc_out + (size_t)f * PCB_FEAT_DIM, (size_t)f * PCB_FEAT_DIM * sizeof(float),
                                  cudaMemcpyDeviceToDevice, c->stream));
    return PCB_OK;
  }
  for (int f0 = 0; f0 < f; f0 += chunk) {
    const int fn = f - f0 < chunk ? f - f0 : chunk;
    const int imgs = mode == 1 ? 2 * fn : fn;
    Model::Run* r = nullptr;
    int rc = model_prepare(c, m, imgs, PCB_CHIP, PCB_CHIP, &r);
    if (rc) return rc;
    rc = pcb_chip_patch_impl(c, chips_dev + (size_t)f0 * PCB_CHIP * PCB_CHIP * 3, fn, mode, r->t[0].data);
    if (rc) return rc;
    rc = model_run(c, m, r);
    if (rc) return rc;
    float* first = mode == 2 ? emb_flip_dev : emb_dev;
    PCB_CUDA(c, cudaMemcpyAsync(first + (size_t)f0 * PCB_FEAT_DIM, r->fc_out, (size_t)fn * PCB_FEAT_DIM * sizeof(float),
                                cudaMemcpyDeviceToDevice, c->stream));
    if (mode == 1)
      PCB_CUDA(c, cudaMemcpyAsync(emb_flip_dev + (size_t)f0 * PCB_FEAT_DIM, r->fc_out + (size_t)fn * PCB_FEAT_DIM,
                                  (size_t)fn * PCB_FEAT_DIM * sizeof(float), cudaMemcpyDeviceToDevice, c->stream));
  }
  return PCB_OK;
}
