"""World-size-2 gloo test (CPU) of the multi-GPU host logic: shard bounds, the
all-gather of fixed-width results and reassembly in input order."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from igm_b200 import dist as idist
from igm_b200._lib import PAIR_RESULT_DTYPE


def test_shard_bounds():
    for n in (0, 1, 7, 8, 9, 1001):
        for w in (1, 2, 3, 8):
            per, b = idist.shard_bounds(n, w)
            assert len(b) == w and b[0][0] == 0 and b[-1][1] == n
            assert all(b[k][1] == b[k + 1][0] for k in range(w - 1))
            assert all(hi - lo <= per for lo, hi in b)
            assert sum(hi - lo for lo, hi in b) == n


def _fake_results(lo, hi):
    r = np.zeros(hi - lo, dtype=PAIR_RESULT_DTYPE)
    k = np.arange(lo, hi)
    r["d2_sel_bits"] = k * 7 + 1
    r["contact_count"] = k % 13
    r["o"] = k
    r["nrec"] = (k % 3) + 1
    r["p"] = k / 1000.0
    r["dist"] = k * 0.5
    r["prob"] = k * 0.25
    return r


def _worker(rank, world, port, n, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)

    def compute(lo, hi, out):
        out.copy_(torch.from_numpy(_fake_results(lo, hi).view(np.uint8).reshape(hi - lo, 32)))
    full, per, bounds = idist.run_sharded(compute, n, rank, world, torch.device("cpu"))
    res = idist.gathered_to_results(full, per, bounds)
    q.put((rank, res.tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("n", [0, 5, 1000, 1001])
def test_gloo_world2_allgather(n):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, n, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = _fake_results(0, n).tobytes()
    assert got[0] == exp and got[1] == exp


class _FakeEngine:
    """Stands in for ActdistEngine in replicate_population: rows of 6 floats per bead."""

    def __init__(self, nbead):
        self.nbead = nbead
        self.buf = torch.zeros((nbead + 1 + 64, 6), dtype=torch.float32)

    def upload_coordinates(self, xyz, bead0=0):
        self.buf[bead0:bead0 + len(xyz)] = torch.from_numpy(np.ascontiguousarray(xyz)).reshape(len(xyz), 6)

    def coords_tensor(self):
        return self.buf


def _replicate_worker(rank, world, port, nbead, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    xyz = np.arange(nbead * 6, dtype=np.float32).reshape(nbead, 2, 3) + 1.0
    eng = _FakeEngine(nbead)
    idist.replicate_population(eng, xyz, rank, world)
    q.put((rank, eng.buf.numpy().tobytes()))
    dist.barrier()
    dist.destroy_process_group()


@pytest.mark.parametrize("nbead", [1, 2, 7, 10])
def test_gloo_world2_replicate_population(nbead):
    """Every rank uploads its share of the beads; one all-gather of whole rows completes both
    copies; the all-zero row behind the population stays zero."""
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_replicate_worker, args=(r, 2, port, nbead, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = dict(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    exp = np.zeros((nbead + 1 + 64, 6), np.float32)
    exp[:nbead] = (np.arange(nbead * 6, dtype=np.float32) + 1.0).reshape(nbead, 6)
    assert got[0] == exp.tobytes() and got[1] == exp.tobytes()


def test_bead_shares():
    for n in (1, 5, 29838):
        for w in (1, 2, 8):
            per, sh = idist.bead_shares(n, w)
            assert sh[0][0] == 0 and sh[-1][1] == n and w * per <= n + 1 + 64
