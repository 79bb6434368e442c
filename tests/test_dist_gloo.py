"""world_size-2 gloo test of the multi-GPU sharding path (pairs sharded across ranks, one final gather of
(model, stats, mask) records).  The per-rank engine is injected: the host emulation stands in for the GPU engine."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _free_port():
    s = socket.socket(); s.bind(("127.0.0.1", 0)); p = s.getsockname()[1]; s.close(); return p


def _engine(p1, p2, seeds=None, **kw):
    from tests.hostemu import emu
    F = np.zeros((len(p1), 3, 3)); M = np.zeros((len(p1), p1.shape[1]), bool); S = np.zeros((len(p1), 4), np.int32)
    for i in range(len(p1)):
        F[i], M[i], S[i] = emu.find_fundamental(p1[i], p2[i], 1.0, 0.99, 300, seed=int(seeds[i]))
    return F, M, S


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"; os.environ["MASTER_PORT"] = str(port)
    import torch.distributed as dist
    from pydegensac_b200.parallel import find_fundamental_sharded
    from pydegensac_b200.scenes import batch_F
    dist.init_process_group("gloo", rank=rank, world_size=world)
    b1, b2 = batch_F(5, 200, 0.5, seed0=3)          # 5 pairs over 2 ranks: ragged split 3 + 2
    out = find_fundamental_sharded(b1, b2, dist, _engine, seeds=np.arange(5, dtype=np.uint64) + 3)
    if rank == 0:
        F, M, S = out
        q.put((F, M, S))
    else:
        assert out is None
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_shard_and_gather():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    F, M, S = q.get(timeout=180)
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    from pydegensac_b200.scenes import batch_F
    b1, b2 = batch_F(5, 200, 0.5, seed0=3)
    Fs, Ms, Ss = _engine(b1, b2, seeds=np.arange(5) + 3)
    assert np.array_equal(F, Fs) and np.array_equal(M, Ms) and np.array_equal(S, Ss)


def test_shard_bounds_and_records():
    from pydegensac_b200.parallel import shard_bounds, pack_records, unpack_records
    assert [shard_bounds(8192, 8, r) for r in range(8)] == [(1024 * r, 1024 * (r + 1)) for r in range(8)]
    assert [shard_bounds(5, 2, r) for r in range(2)] == [(0, 3), (3, 5)]
    assert [shard_bounds(2, 4, r) for r in range(4)] == [(0, 1), (1, 2), (2, 2), (2, 2)]
    rng = np.random.default_rng(0)
    F = rng.normal(size=(3, 3, 3)); M = rng.integers(0, 2, (3, 17)).astype(bool); S = rng.integers(0, 1000, (3, 4)).astype(np.int32)
    rec = pack_records(F, M, S)
    assert rec.shape == (3, 72 + 16 + 17)
    F2, M2, S2 = unpack_records(rec)
    assert np.array_equal(F, F2) and np.array_equal(M, M2) and np.array_equal(S, S2)
