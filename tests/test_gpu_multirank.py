"""Hardware check of SURVEY.md 8(e): N ranks (one per GPU, NCCL) reproduce the single-rank pre-scan exactly -- spans, bank
rows and every per-sample decision / best distance -- in both flip modes (lazy: predicted + on-demand flips exchanged
between ranks; eager: both variants embedded before the gather).  Skips on a box with fewer than 2 GPUs."""
import json
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


@pytest.mark.parametrize("mode", ["lazy", "eager"])
def test_n_ranks_equal_single_rank(tmp_path, mode):
    import torch
    n = torch.cuda.device_count()
    if n < 2:
        pytest.skip("needs at least 2 GPUs")
    world = 2 if n < 4 else 4
    out = tmp_path / f"multirank_{mode}.json"
    port = 29700 + (os.getpid() % 200)
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", f"--nproc-per-node={world}", "--master-addr", "127.0.0.1",
           "--master-port", str(port), os.path.join(HERE, "multirank_worker.py"), str(out), mode]
    r = subprocess.run(cmd, capture_output=True, text=True, timeout=900)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    res = json.loads(out.read_text())
    assert len(res) == world
    for rec in res:
        assert rec["spans_n"] == rec["spans_1"] == res[0]["spans_1"], rec
        assert rec["bank_equal"] and rec["same_log"], rec
    assert len(res[0]["spans_1"]) >= 2 and res[0]["bank_rows"] > res[0]["bank_rows_initial"]
    for rec in res:                                   # span-sharded main pass == sequential main pass on every rank
        assert rec["main_equal"] and rec["main_hits"] >= 20 and "lock_roi" in rec["main_sites"], rec
