// Host-side replay of the reference's sequential pre-scan loop over precomputed superset records
// (person_capture/gui_app.py:1468-1655: fd9 skip gate :1479-1492, per-face best distance / bank offers :1512-1549,
// hysteresis + span closing :1587-1622, tail :1648-1655; rotation choice of the fast pre-scan,
// person_capture/face_embedder.py:2354-2388).  Every rank of a multi-GPU pre-scan replays ALL samples, so this loop
// bounds the scaling; it is native code over flat arrays (~50 ns per sample) and calls back into the host language
// only for the rare events that need the GPU or the bank: a possible bank update, or flip-TTA features that were not
// predicted.  No CUDA in this file.
#include <algorithm>
#include <vector>

#include "../../include/pcb200.h"

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

extern "C" int pcb_replay(const pcb_replay_cfg* cfg, const int32_t* meta, const int64_t* frame_idx, int n_samples,
                          const double* quality, const int64_t* area, const uint8_t* flip_ready, const double* fd_plain,
                          const double* fd_flip, pcb_replay_state* st, pcb_replay_offer_cb offer, pcb_replay_flip_cb need_flip,
                          void* user, double* best_out, uint8_t* skip_out, uint8_t* active_out, int32_t* nfaces_out,
                          int64_t* spans_out, int max_spans, int32_t* n_spans_out) {
  if (!cfg || !meta || !frame_idx || !st || n_samples < 0) return 2;
  Tracker trk;
  trk.c = cfg;
  long long last_add = -1000000000LL;
  std::vector<int> order;
  for (int s = 0; s < n_samples; ++s) {
    const int32_t* m = meta + (size_t)s * PCB_REPLAY_META;
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
        if (flip_ready && active) {
          bool all = true;
          for (int i = 0; i < c_cnt; ++i) all = all && flip_ready[c_start + i];
          if (!all && need_flip) need_flip(user, s);      // refreshes flip_ready / fd arrays in place
        }
        const double* fd = active ? fd_flip : fd_plain;
        bool may_offer = false;
        if ((long long)s - last_add >= cfg->cooldown)
          for (int i = 0; i < c_cnt; ++i)
            may_offer = may_offer || (fd[c_start + i] <= cfg->fd_add && quality[c_start + i] >= cfg->quality_min);
        if (may_offer) {
          // a bank update is possible: the reference's face order (quality, area descending) matters, later faces
          // see the updated bank (the callback rewrites the fd arrays in place)
          order.resize(c_cnt);
          for (int i = 0; i < c_cnt; ++i) order[i] = i;
          std::stable_sort(order.begin(), order.end(), [&](int a, int b) {
            const double qa = quality[c_start + a], qb = quality[c_start + b];
            if (qa != qb) return qa > qb;
            return area[c_start + a] > area[c_start + b];
          });
          for (int k = 0; k < c_cnt; ++k) {
            const int row = c_start + order[k];
            const double f = (active ? fd_flip : fd_plain)[row];
            if (f < best) best = f;
            if (f <= cfg->fd_add && (long long)s - last_add >= cfg->cooldown && quality[row] >= cfg->quality_min) {
              if (offer && offer(user, s, row, quality[row], active ? 1 : 0)) last_add = s;
            }
          }
        } else {
          for (int i = 0; i < c_cnt; ++i)
            if (fd[c_start + i] < best) best = fd[c_start + i];
        }
      }
    }
    if (best_out) best_out[s] = best;
    if (skip_out) skip_out[s] = skipped ? 1 : 0;
    if (active_out) active_out[s] = active ? 1 : 0;
    if (nfaces_out) nfaces_out[s] = nfaces;
    trk.observe(frame_idx[s], best);
  }
  trk.finish();
  const int n_spans = (int)(trk.spans.size() / 2);
  if (n_spans > max_spans) return 2;
  for (size_t i = 0; i < trk.spans.size(); ++i) spans_out[i] = trk.spans[i];
  *n_spans_out = n_spans;
  st->trk_active = trk.active ? 1 : 0;
  return 0;
}
