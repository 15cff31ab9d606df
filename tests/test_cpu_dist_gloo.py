"""CPU: the multi-rank exchange with world_size 2 on gloo -- each rank owns a contiguous
chunk of samples, records are all-gathered, and every rank's replay equals the single-process replay."""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from person_capture_b200 import prescan as PS
    from person_capture_b200.params import PrescanParams
    import test_cpu_host_logic as T
    rng = np.random.default_rng(11)
    n = 120
    target = T.unit(rng.normal(size=512))
    sc = T.make_scenario(rng, n, target)
    cfg = PrescanParams(prescan_stride=1, prescan_max_width=10 ** 6, prescan_fd_add=0.3, prescan_add_cooldown_samples=2,
                        face_quality_min=50.0, prescan_min_segment_sec=0.25, prescan_pad_sec=0.1)
    ref = T.unit(target + rng.normal(0, 0.03, 512))[None]
    idxs = PS.sample_indices(n, 1)
    per = (len(idxs) + world - 1) // world
    mine = idxs[rank * per:(rank + 1) * per]
    records, P, Fl = T.to_records({i: sc[i] for i in mine})
    table = PS.FaceTable()
    table.count = len(P)
    table.plain = torch.from_numpy(P) if len(P) else torch.zeros((1, 512))
    table.flip = torch.from_numpy(Fl) if len(Fl) else torch.zeros((1, 512))
    new_table, allp, allf, enc = PS._gather_shards(None, PS.encode_records(records, mine, len(P)), table, P, Fl, world, None)
    log = []
    trk, bank = PS.replay(None, None, (allp, allf), idxs, 24, n, T.FakeFace(sc), ref, cfg, log=log,
                          distances=T.NumpyDistances(allp, allf), encoded=enc)
    q.put((rank, trk.finish(), [(r["idx"], r["skip"], round(r["best"], 6)) for r in log], len(bank)))
    dist.destroy_process_group()


class _FakeLazyTable:
    """Stands in for a rank's lazy FaceTable: flips exist only after ensure_flip (filled from the scripted truth)."""
    lazy = True

    def __init__(self, P, Fl, ready_rows):
        self.count = len(P)
        self.truth = Fl
        self.plain = torch.from_numpy(P) if len(P) else torch.zeros((1, 512))
        self.flip = torch.zeros_like(self.plain)
        self.flip_ready = np.zeros(self.count, bool)
        self.flip_host = np.zeros((self.count, 512), np.float32)
        self.computed = 0
        self.ensure_flip(None, np.asarray(ready_rows, np.int64))

    def ensure_flip(self, eng, rows):
        rows = np.asarray(rows, np.int64)
        need = np.unique(rows[~self.flip_ready[rows]]) if len(rows) else rows
        if not len(need):
            return False
        self.flip_host[need] = self.truth[need]
        self.flip[torch.as_tensor(need)] = torch.from_numpy(self.truth[need])
        self.flip_ready[need] = True
        self.computed += len(need)
        return True


class _TableDistances:
    """fd of every row against the live bank from the (growing) host copies of a lazy table."""

    def __init__(self, plain_h, table):
        self.p, self.t = plain_h, table

    def invalidate(self):
        pass

    def get(self, bank):
        from oracle import prescan as OP
        B = bank.array()
        if B is None:
            return np.full(len(self.p), 9.0), np.full(len(self.p), 9.0)
        fdf = np.array([OP.fd_min(v, B) if ok else np.nan for v, ok in zip(self.t.flip_host, self.t.flip_ready)])
        return np.array([OP.fd_min(v, B) for v in self.p]), fdf


def _worker_lazy(rank, world, port, q):
    sys.path.insert(0, ROOT)
    sys.path.insert(0, HERE)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from person_capture_b200 import prescan as PS
    from person_capture_b200.params import PrescanParams
    import test_cpu_host_logic as T
    rng = np.random.default_rng(11)
    n = 120
    target = T.unit(rng.normal(size=512))
    sc = T.make_scenario(rng, n, target)
    cfg = PrescanParams(prescan_stride=1, prescan_max_width=10 ** 6, prescan_fd_add=0.3, prescan_add_cooldown_samples=2,
                        face_quality_min=50.0, prescan_min_segment_sec=0.25, prescan_pad_sec=0.1)
    ref = T.unit(target + rng.normal(0, 0.03, 512))[None]
    idxs = PS.sample_indices(n, 1)
    per = (len(idxs) + world - 1) // world
    mine = idxs[rank * per:(rank + 1) * per]
    records, P, Fl = T.to_records({i: sc[i] for i in mine})
    # flips predicted from the plain distances to the initial bank with a deliberately small margin, so that the replay
    # has to fetch some on demand across ranks
    from oracle import prescan as OP
    fd0 = np.array([OP.fd_min(v, ref) for v in P]) if len(P) else np.zeros((0,))
    pred = PS._predict_flip_rows(records, mine, fd0, cfg, 24, carry_in=False, margin=-0.3)
    table = _FakeLazyTable(P, Fl, pred)
    predicted = int(table.flip_ready.sum())
    new_table, allp, allf, enc = PS._gather_shards(None, PS.encode_records(records, mine, len(P)), table, P, None, world, None)
    log = []
    trk, bank = PS.replay(None, new_table, (allp, None), idxs, 24, n, T.FakeFace(sc), ref, cfg, log=log,
                          distances=_TableDistances(allp, new_table), encoded=enc)
    q.put((rank, trk.finish(), [(r["idx"], r["skip"], round(r["best"], 6)) for r in log], len(bank), predicted, table.computed,
           int(new_table.flip_ready.sum()), new_table.count))
    dist.destroy_process_group()


def test_two_rank_lazy_flip_on_demand_equals_single_process():
    """Multi-rank pre-scan with predicted + on-demand flip features: same log / spans / bank as the eager
    single-process replay, and strictly fewer flip passes than embedding both variants of every face."""
    sys.path.insert(0, HERE)
    from person_capture_b200 import prescan as PS
    from person_capture_b200.params import PrescanParams
    import test_cpu_host_logic as T
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker_lazy, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=180) for _ in procs])
    for p in procs:
        p.join(timeout=60)
    rng = np.random.default_rng(11)
    n = 120
    target = T.unit(rng.normal(size=512))
    sc = T.make_scenario(rng, n, target)
    cfg = PrescanParams(prescan_stride=1, prescan_max_width=10 ** 6, prescan_fd_add=0.3, prescan_add_cooldown_samples=2,
                        face_quality_min=50.0, prescan_min_segment_sec=0.25, prescan_pad_sec=0.1)
    ref = T.unit(target + rng.normal(0, 0.03, 512))[None]
    records, P, Fl = T.to_records(sc)
    log = []
    trk, bank = PS.replay(records, None, (P, Fl), PS.sample_indices(n, 1), 24, n, T.FakeFace(sc), ref, cfg, log=log,
                          distances=T.NumpyDistances(P, Fl))
    want = (trk.finish(), [(r["idx"], r["skip"], round(r["best"], 6)) for r in log], len(bank))
    for o in outs:
        assert (o[1], o[2], o[3]) == want
    computed = sum(o[5] for o in outs)
    predicted = sum(o[4] for o in outs)
    assert computed > predicted            # some flips were fetched on demand during the replay
    assert outs[0][6] == outs[1][6] < outs[0][7]      # both ranks end with the same ready set, smaller than the table
    assert len(want[0]) >= 1


def test_two_rank_gather_and_replay_equal_single_process():
    sys.path.insert(0, HERE)
    from person_capture_b200 import prescan as PS
    from person_capture_b200.params import PrescanParams
    import test_cpu_host_logic as T
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + (os.getpid() % 2000)
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    outs = sorted([q.get(timeout=180) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process result
    rng = np.random.default_rng(11)
    n = 120
    target = T.unit(rng.normal(size=512))
    sc = T.make_scenario(rng, n, target)
    cfg = PrescanParams(prescan_stride=1, prescan_max_width=10 ** 6, prescan_fd_add=0.3, prescan_add_cooldown_samples=2,
                        face_quality_min=50.0, prescan_min_segment_sec=0.25, prescan_pad_sec=0.1)
    ref = T.unit(target + rng.normal(0, 0.03, 512))[None]
    records, P, Fl = T.to_records(sc)
    log = []
    trk, bank = PS.replay(records, None, (P, Fl), PS.sample_indices(n, 1), 24, n, T.FakeFace(sc), ref, cfg, log=log,
                          distances=T.NumpyDistances(P, Fl))
    want = (trk.finish(), [(r["idx"], r["skip"], round(r["best"], 6)) for r in log], len(bank))
    assert outs[0][1:] == want and outs[1][1:] == want
    assert len(want[0]) >= 1
