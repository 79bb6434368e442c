"""Multi-GPU sharding of a batch of image pairs.

Pairs are independent units: rank r owns the contiguous block [r*ceil(P/G), ...) and runs the single-GPU
engine on it; there is NO data-path collective.  The only exchange is the final gather of fixed-stride
records (model[9] f64 | stats[4] i32 | mask[N] u8) to rank 0 -- one `torch.distributed` gather
(NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import numpy as np


def shard_bounds(n_pairs, world_size, rank):
    """Contiguous block partition: ceil(P/G) pairs per rank (the last ranks may get fewer or none)."""
    per = (n_pairs + world_size - 1) // world_size
    lo = min(n_pairs, rank * per)
    hi = min(n_pairs, lo + per)
    return lo, hi


def pack_records(model, mask, stats):
    """[p,3,3] f64, [p,N] bool/u8, [p,4] i32 -> [p, 72+16+N] uint8 records."""
    p = model.shape[0]
    n = mask.shape[1]
    rec = np.zeros((p, 72 + 16 + n), dtype=np.uint8)
    rec[:, :72] = np.ascontiguousarray(model.reshape(p, 9)).view(np.uint8).reshape(p, 72)
    rec[:, 72:88] = np.ascontiguousarray(stats.astype(np.int32)).view(np.uint8).reshape(p, 16)
    rec[:, 88:] = mask.astype(np.uint8)
    return rec


def unpack_records(rec):
    p = rec.shape[0]
    n = rec.shape[1] - 88
    model = np.ascontiguousarray(rec[:, :72]).view(np.float64).reshape(p, 3, 3)
    stats = np.ascontiguousarray(rec[:, 72:88]).view(np.int32).reshape(p, 4)
    mask = rec[:, 88:].astype(bool)
    return model, mask, stats


def gather_records(rec_local, n_pairs, dist, device=None):
    """Gather every rank's records on rank 0 (returns [P, stride] uint8 there, None elsewhere).

    `dist` is torch.distributed (already initialised).  Ranks pad their block to ceil(P/G) rows so one
    fixed-size gather suffices."""
    import torch
    world, rank = dist.get_world_size(), dist.get_rank()
    per = (n_pairs + world - 1) // world
    stride = rec_local.shape[1]
    buf = torch.zeros((per, stride), dtype=torch.uint8, device=device)
    if rec_local.shape[0]:
        buf[:rec_local.shape[0]] = torch.from_numpy(rec_local).to(buf.device)
    if rank == 0:
        parts = [torch.empty_like(buf) for _ in range(world)]
        dist.gather(buf, gather_list=parts, dst=0)
        full = torch.cat(parts, 0)[:n_pairs].cpu().numpy()
        return full
    dist.gather(buf, gather_list=None, dst=0)
    return None


def find_fundamental_sharded(pts1, pts2, dist, engine, device=None, **kw):
    """Run `engine(p1_block, p2_block, seeds=..., **kw) -> (F, mask, stats)` on this rank's block of the batch and
    gather to rank 0.  `engine` is pydegensac_b200._cabi-based on GPUs; tests inject a CPU stand-in."""
    P = pts1.shape[0]
    lo, hi = shard_bounds(P, dist.get_world_size(), dist.get_rank())
    seeds = kw.pop("seeds", None)
    if seeds is None:
        seeds = np.arange(P, dtype=np.uint64)
    if hi > lo:
        F, mask, stats = engine(pts1[lo:hi], pts2[lo:hi], seeds=np.asarray(seeds)[lo:hi], **kw)
    else:
        n = pts1.shape[1]
        F, mask, stats = np.zeros((0, 3, 3)), np.zeros((0, n), bool), np.zeros((0, 4), np.int32)
    rec = pack_records(F, mask, stats)
    full = gather_records(rec, P, dist, device)
    if full is None:
        return None
    return unpack_records(full)
