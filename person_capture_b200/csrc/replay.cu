// Host-side replay of the reference's sequential pre-scan loop over precomputed superset records
// (person_capture/gui_app.py:1468-1655: fd9 skip gate :1479-1492, per-face best distance / bank offers :1512-1549,
// hysteresis + span closing :1587-1622, tail :1648-1655; rotation choice of the fast pre-scan,
// person_capture/face_embedder.py:2354-2388) and of the live reference bank (Processor._stream_ref_bank_update,
// gui_app.py:922-986).  Every rank of a multi-GPU pre-scan replays ALL samples, so this loop bounds the scaling: it is
// native code over flat arrays (~50 ns per sample), the bank update is native (a few 512-d dot products per offer) and
// the distances of the not-yet-visited face rows are refreshed on the GPU through pcb_live_refresh (match.cu) -- one
// short launch + one small device->host copy per bank change.  The only callback left is the rare "flip-TTA features
// missing" event, which needs an ArcFace pass.  No kernels in this file.
#include <math.h>
#include <string.h>

#include <algorithm>
#include <vector>

#include <stdlib.h>

#include "../../include/pcb200.h"

// match.cu: one 512-float row of a device feature table -> host (synchronous on the context's stream)
int pcb_fetch_row(pcb_ctx* c, const float* row_dev, float* out_host);

// ---------------------------------------------------------------------------------------------------------------
// live bank
// ---------------------------------------------------------------------------------------------------------------
struct pcb_bank {
  pcb_bank_cfg cfg;
  int rows = 0;
  long long version = 0;
  std::vector<float> data;   // [cap][512]
  std::vector<float> gram;   // [cap][cap] pairwise cosines of the rows (kept current row by row)
};

namespace {

constexpr int kD = PCB_FEAT_DIM;

// float32 dot product with 16 independent partial sums (vectorises; the order is fixed, so every rank of a multi-GPU
// replay -- and every run -- gets the same bits)
inline float dot512(const float* a, const float* b) {
  float acc[16];
  for (int l = 0; l < 16; ++l) acc[l] = 0.f;
  for (int i = 0; i < kD; i += 16)
    for (int l = 0; l < 16; ++l) acc[l] += a[i + l] * b[i + l];
  for (int w = 8; w > 0; w >>= 1)
    for (int l = 0; l < w; ++l) acc[l] += acc[l + w];
  return acc[0];
}

void gram_row(pcb_bank* b, int k) {
  const int cap = b->cfg.cap;
  for (int j = 0; j < b->rows; ++j) {
    const float g = dot512(&b->data[(size_t)k * kD], &b->data[(size_t)j * kD]);
    b->gram[(size_t)k * cap + j] = g;
    b->gram[(size_t)j * cap + k] = g;
  }
}

}  // namespace

extern "C" pcb_bank* pcb_bank_create(const pcb_bank_cfg* cfg, const float* rows_host, int n) {
  if (!cfg || n < 0 || (n > 0 && !rows_host)) return nullptr;
  pcb_bank* b = new pcb_bank();
  b->cfg = *cfg;
  if (b->cfg.cap < 1) b->cfg.cap = 1;
  if (b->cfg.cap < n) b->cfg.cap = n;          // an initial bank larger than the cap is kept as it is (the reference never trims it)
  b->data.assign((size_t)b->cfg.cap * kD, 0.f);
  b->gram.assign((size_t)b->cfg.cap * b->cfg.cap, 0.f);
  for (int i = 0; i < n; ++i) {
    // RefBank.__init__: rows / max(|row|, 1e-6) in float32
    const float* src = rows_host + (size_t)i * kD;
    float nv = sqrtf(dot512(src, src));
    if (nv < 1e-6f) nv = 1e-6f;
    for (int d = 0; d < kD; ++d) b->data[(size_t)i * kD + d] = src[d] / nv;
    b->rows = i + 1;
    gram_row(b, i);
  }
  return b;
}

extern "C" void pcb_bank_destroy(pcb_bank* b) { delete b; }
extern "C" int pcb_bank_rows(const pcb_bank* b) { return b ? b->rows : 0; }
extern "C" long long pcb_bank_version(const pcb_bank* b) { return b ? b->version : 0; }
extern "C" const float* pcb_bank_data(const pcb_bank* b) { return b ? b->data.data() : nullptr; }

extern "C" int pcb_bank_offer(pcb_bank* b, const float* vec, double quality, int32_t* slot_out) {
  if (!b || !vec) return PCB_BANK_SKIP;
  const pcb_bank_cfg& c = b->cfg;
  float v[kD];
  const float nv = sqrtf(dot512(vec, vec));
  if (!(nv > 1e-6f)) return PCB_BANK_SKIP;
  for (int d = 0; d < kD; ++d) v[d] = vec[d] / nv;
  auto put = [&](int k) {
    memcpy(&b->data[(size_t)k * kD], v, sizeof v);
    gram_row(b, k);
    b->version++;
    if (slot_out) *slot_out = k;
  };
  if (b->rows == 0) {
    b->rows = 1;
    put(0);
    return PCB_BANK_ADDED;
  }
  float top = -3.0e38f, sim0 = 0.f;
  for (int j = 0; j < b->rows; ++j) {
    const float s = dot512(&b->data[(size_t)j * kD], v);
    if (j == 0) sim0 = s;
    if (s > top) top = s;
  }
  if ((double)top >= c.dedup) return PCB_BANK_DUP;
  if (b->rows < c.cap) {
    b->rows += 1;
    put(b->rows - 1);
    return PCB_BANK_ADDED;
  }
  // full: score the candidate against the worst row (anchor = row 0; float64 scalars for the candidate, float32 vector
  // arithmetic for the bank rows, as the numpy statement of the reference computes them)
  double cos_a = (double)sim0;
  cos_a = cos_a < -1.0 ? -1.0 : (cos_a > 1.0 ? 1.0 : cos_a);
  double t = 2.0 - 2.0 * cos_a;
  if (t < 0.0) t = 0.0;
  double q = quality > 0.0 ? quality : 0.0;
  if (q > 1000.0) q = 1000.0;
  const double s_new = c.wa * (1.0 - sqrt(t)) + c.wd * (1.0 - (double)top) + c.wq * (q / 300.0);
  const int cap = c.cap;
  int worst = 0;
  float s_worst = 0.f;
  for (int i = 0; i < b->rows; ++i) {
    float gmax = -1.0f;                           // diagonal counts as -1
    for (int j = 0; j < b->rows; ++j)
      if (j != i && b->gram[(size_t)i * cap + j] > gmax) gmax = b->gram[(size_t)i * cap + j];
    float ca = b->gram[(size_t)i * cap + 0];
    ca = ca < -1.0f ? -1.0f : (ca > 1.0f ? 1.0f : ca);
    float u = 2.0f - 2.0f * ca;
    if (u < 0.0f) u = 0.0f;
    const float s = (float)c.wa * (1.0f - sqrtf(u)) + (float)c.wd * (1.0f - gmax);
    if (i == 0 || s < s_worst) { s_worst = s; worst = i; }
  }
  if (s_new > (double)s_worst + c.margin) {
    put(worst);
    return PCB_BANK_REPLACED;
  }
  return PCB_BANK_SKIP;
}

// ---------------------------------------------------------------------------------------------------------------
// span state machine
// ---------------------------------------------------------------------------------------------------------------
namespace {

struct Tracker {
  const pcb_replay_cfg* c;
  std::vector<int64_t> spans;   // flattened (s, e)
  bool active = false;
  int64_t start = 0;
  int neg_run = 0;
  int fd9_streak = 0;

  bool gate_skips() const {
    if (active || !c->fd9_skip) return false;
    return fd9_streak >= c->fd9_grace && (fd9_streak % c->fd9_period) != 0;
  }
  void close(int64_t s, int64_t e) {
    if (e - s + 1 >= c->min_len) {
      if (!spans.empty() && s <= spans[spans.size() - 1] + 1) {
        if (e > spans[spans.size() - 1]) spans[spans.size() - 1] = e;
      } else {
        spans.push_back(s);
        spans.push_back(e);
      }
    }
  }
  void observe(int64_t idx, double best) {
    fd9_streak = best >= 8.99 ? fd9_streak + 1 : 0;
    if (best <= c->enter) {
      if (!active) {
        active = true;
        fd9_streak = 0;
        start = idx;
      }
      neg_run = 0;
    } else if (active) {
      neg_run += 1;
      if ((int64_t)neg_run * c->stride >= c->exit_cool || best >= c->exit_thr) {
        close(std::max<int64_t>(0, start - c->pad), std::min<int64_t>(c->total_frames - 1, idx + c->pad));
        active = false;
        neg_run = 0;
        fd9_streak = 0;
      }
    }
  }
  void finish() {
    if (active) close(std::max<int64_t>(0, start - c->pad), c->total_frames - 1);
  }
};

}  // namespace

extern "C" int pcb_replay(pcb_ctx* ctx, const pcb_replay_cfg* cfg, pcb_bank* bank, const pcb_replay_io* io, pcb_replay_state* st) {
  if (!cfg || !bank || !io || !st || !io->meta || !io->frame_idx || io->n_samples < 0 || !io->fd_plain || !io->fd_flip) return 2;
  if (!ctx && !io->refresh) return 2;
  const int n_samples = io->n_samples, n_rows = io->n_rows;
  long long refreshes = 0;
  // first face-table row any sample >= s can read: rows below it never need another distance
  std::vector<int> row_lo((size_t)n_samples + 1, n_rows);
  for (int s = n_samples - 1; s >= 0; --s) {
    const int32_t* m = io->meta + (size_t)s * PCB_REPLAY_META;
    int lo = row_lo[s + 1];
    if (m[0] >= 0 && m[0] < lo) lo = m[0];
    if (m[6] >= 0 && m[6] < lo) lo = m[6];
    if (m[8] >= 0 && m[8] < lo) lo = m[8];
    row_lo[s] = lo;
  }
  auto refresh = [&](int changed_slot, int s) -> int {
    ++refreshes;
    const int lo = row_lo[s];
    if (ctx) {
      if (n_rows == 0) return 0;
      const float* sim = nullptr;
      const int rc = pcb_live_refresh(ctx, pcb_bank_data(bank), pcb_bank_rows(bank), changed_slot, lo, 2, &sim);
      if (rc) return -rc;
      for (int r = lo; r < n_rows; ++r) {
        io->fd_plain[r] = 1.0 - (double)sim[r];
        io->fd_flip[r] = 1.0 - (double)sim[n_rows + r];
      }
      return 0;
    }
    return io->refresh(io->user, pcb_bank_data(bank), pcb_bank_rows(bank), changed_slot, lo);
  };
  if (refresh(-1, 0) < 0) return PCB_REPLAY_ABORTED;

  Tracker trk;
  trk.c = cfg;
  const char* df_env = getenv("PCB_REPLAY_DUP_FILTER");
  const bool dup_filter = !(df_env && df_env[0] == '0');
  const double dedup_certain = bank->cfg.dedup + 1e-4;
  float fetched[PCB_FEAT_DIM];
  long long last_add = -1000000000LL;
  std::vector<int> order;
  for (int s = 0; s < n_samples; ++s) {
    const int32_t* m = io->meta + (size_t)s * PCB_REPLAY_META;
    const bool active = trk.active;
    double best = 9.0;
    const bool skipped = trk.gate_skips();
    int nfaces = 0;
    if (!skipped) {
      st->frame_idx += 1;
      int c_start = m[0], c_cnt = m[1];
      if (c_start < 0) {
        st->no_face_streak += 1;
        st->rot_cycle += 1;
        int degs[2];
        int nd;
        if (active) {
          degs[0] = 0; degs[1] = 1; nd = 2;             // 90 then 270
        } else {
          degs[0] = (int)(st->prescan_rr % 2); nd = 1;  // round robin
          st->prescan_rr += 1;
        }
        for (int k = 0; k < nd; ++k) {
          const int d = degs[k];                        // 0: 90 deg, 1: 270 deg
          if (m[2 + d] == 0 || m[4 + d] == 0) continue; // no probe hit, or the heavy pass found nothing
          if (m[6 + 2 * d] >= 0) {
            c_start = m[6 + 2 * d];
            c_cnt = m[7 + 2 * d];
            break;
          }
        }
      } else {
        st->no_face_streak = 0;
        st->last_face_idx = st->frame_idx;
        st->rot_cycle = 0;
      }
      if (c_start >= 0) {
        nfaces = c_cnt;
        if (io->flip_ready && active) {
          bool all = true;
          for (int i = 0; i < c_cnt; ++i) all = all && io->flip_ready[c_start + i];
          if (!all) {
            // the callee computes the missing features (and re-arms the live table); every distance is stale after that
            if (!io->need_flip || io->need_flip(io->user, s) < 0) return PCB_REPLAY_ABORTED;
            if (refresh(-1, s) < 0) return PCB_REPLAY_ABORTED;
          }
        }
        const double* fd = active ? io->fd_flip : io->fd_plain;
        bool may_offer = false;
        if ((long long)s - last_add >= cfg->cooldown)
          for (int i = 0; i < c_cnt; ++i)
            may_offer = may_offer || (fd[c_start + i] <= cfg->fd_add && io->quality[c_start + i] >= cfg->quality_min);
        if (may_offer) {
          // a bank update is possible: the reference's face order (quality, area descending) matters, later faces
          // see the updated bank (the refresh rewrites the fd arrays in place)
          order.resize(c_cnt);
          for (int i = 0; i < c_cnt; ++i) order[i] = i;
          std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            const double qa = io->quality[c_start + a], qb = io->quality[c_start + b];
            if (qa != qb) return qa > qb;
            return io->area[c_start + a] > io->area[c_start + b];
          });
          for (int k = 0; k < c_cnt; ++k) {
            const int row = c_start + order[k];
            const double f = fd[row];
            if (f < best) best = f;
            if (f <= cfg->fd_add && (long long)s - last_add >= cfg->cooldown && io->quality[row] >= cfg->quality_min) {
              // `f` is 1 - max cosine of this row against the CURRENT bank (the refresh keeps it so), computed in fp32 on the
              // GPU; pcb_bank_offer's own fp32 maximum differs from it by < 3.1e-5 (512 x 2^-24 for unit vectors), so a
              // similarity at least 1e-4 above the de-duplication threshold is "dup" whatever the rounding: the offer (35+
              // dot products on the host, for every target face of the clip once the bank has settled) and the feature row
              // it would read are skipped.  Anything closer to the threshold takes the exact path.
              if (dup_filter && pcb_bank_rows(bank) > 0 && 1.0 - f >= dedup_certain) continue;
              const float* vec = nullptr;
              const float* host = active ? io->feat_flip : io->feat_plain;
              if (host) {
                vec = host + (size_t)row * PCB_FEAT_DIM;
              } else {
                const float* dev = active ? io->feat_flip_dev : io->feat_plain_dev;
                if (!ctx || !dev) return 2;
                if (pcb_fetch_row(ctx, dev + (size_t)row * PCB_FEAT_DIM, fetched)) return PCB_REPLAY_ABORTED;
                vec = fetched;
              }
              int32_t slot = -1;
              const int act = pcb_bank_offer(bank, vec, io->quality[row], &slot);
              if (act == PCB_BANK_ADDED || act == PCB_BANK_REPLACED) {
                last_add = s;
                if (refresh(slot, s) < 0) return PCB_REPLAY_ABORTED;
              }
            }
          }
        } else {
          for (int i = 0; i < c_cnt; ++i)
            if (fd[c_start + i] < best) best = fd[c_start + i];
        }
      }
    }
    if (io->best_out) io->best_out[s] = best;
    if (io->skip_out) io->skip_out[s] = skipped ? 1 : 0;
    if (io->active_out) io->active_out[s] = active ? 1 : 0;
    if (io->nfaces_out) io->nfaces_out[s] = nfaces;
    trk.observe(io->frame_idx[s], best);
  }
  trk.finish();
  const int n_spans = (int)(trk.spans.size() / 2);
  if (n_spans > io->max_spans) return 2;
  for (size_t i = 0; i < trk.spans.size(); ++i) io->spans_out[i] = trk.spans[i];
  *io->n_spans_out = n_spans;
  st->trk_active = trk.active ? 1 : 0;
  if (io->refreshes_out) *io->refreshes_out = refreshes;
  return 0;
}
