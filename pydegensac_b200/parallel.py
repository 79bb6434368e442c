"""Multi-GPU sharding of a batch of image pairs.

Pairs are independent units: rank r owns the contiguous block [r*ceil(P/G), ...) and runs the single-GPU
engine on it; there is NO data-path collective.  The only exchange is the final gather of fixed-stride
records (model[9] f64 | stats[4] i32 | mask[N] u8) to rank 0 -- one `torch.distributed` gather
(NCCL over NVLink on GPUs, gloo in the CPU tests).
"""
import os

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


# ------------------------------------------------------------------------------------------------------------------
# Device-resident sharded run (what bench.py times end to end at N GPUs): every rank copies ITS block of the batch from
# pinned host memory, runs the kernel through the device-pointer C ABI, packs fixed-stride records on the device and
# contributes them to ONE gather; rank 0 reads the gathered records back (the step's result).  No host-side packing, no
# re-upload, no intermediate host copy on the other ranks.
# ------------------------------------------------------------------------------------------------------------------
def gpu_numa_cpus(device_index):
    """CPUs of the NUMA node the GPU hangs off (None when the topology cannot be read)."""
    try:
        import torch
        prop = torch.cuda.get_device_properties(device_index)
        bus = "%04x:%02x:%02x.0" % (getattr(prop, "pci_domain_id", 0), prop.pci_bus_id, prop.pci_device_id)
        node = int(open("/sys/bus/pci/devices/%s/numa_node" % bus).read().strip())
        if node < 0:
            return None
        cpus = set()
        for part in open("/sys/devices/system/node/node%d/cpulist" % node).read().strip().split(","):
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        return cpus or None
    except Exception:
        return None


def pin_to_gpu_numa(device_index):
    """Restrict this process to the CPUs next to its GPU (pinned-memory copies then stay on the local socket).
    Returns the number of CPUs kept, or 0 when nothing was changed."""
    cpus = gpu_numa_cpus(device_index)
    if not cpus:
        return 0
    try:
        allowed = os.sched_getaffinity(0) & cpus
        if allowed:
            os.sched_setaffinity(0, allowed)
            return len(allowed)
    except Exception:
        pass
    return 0


class ShardedBatch:
    """Persistent buffers + one call per step for a rank's block of a sharded batch (kind 'F' or 'H').

    step(h1, h2, seeds) -> on rank 0 a pinned uint8 array [world*P, 88+N] of (model | stats | mask) records, else None.
    `dist` = torch.distributed (initialised) or None for a single GPU."""

    def __init__(self, kind, P, N, dim, params, device, dist=None):
        import torch
        self.kind, self.P, self.N, self.dim, self.params, self.dev, self.dist = kind, P, N, dim, params, device, dist
        self.world = dist.get_world_size() if dist is not None else 1
        self.rank = dist.get_rank() if dist is not None else 0
        self.d1 = torch.empty((P, N, dim), dtype=torch.float64, device=device)
        self.d2 = torch.empty((P, N, dim), dtype=torch.float64, device=device)
        self.dseed = torch.empty(P, dtype=torch.int64, device=device)
        self.model = torch.zeros((P, 9), dtype=torch.float64, device=device)
        self.mask = torch.zeros((P, N), dtype=torch.uint8, device=device)
        self.stats = torch.zeros((P, 4), dtype=torch.int32, device=device)
        self.stride = 72 + 16 + N
        self.rec = torch.empty((P, self.stride), dtype=torch.uint8, device=device)
        self.parts = None
        self.h_out = None
        if self.rank == 0:
            self.gathered = torch.empty((self.world, P, self.stride), dtype=torch.uint8, device=device)
            self.parts = [self.gathered[r] for r in range(self.world)]
            self.h_out = torch.empty((self.world * P, self.stride), dtype=torch.uint8).pin_memory()
        self.h2d_bytes = 2 * P * N * dim * 8 + P * 8
        self.d2h_bytes = self.world * P * self.stride if self.rank == 0 else 0

    def launch(self, stream):
        from . import _cabi
        q = self.params
        if self.kind == "F":
            _cabi.fundamental_batch_dev(self.d1.data_ptr(), self.d2.data_ptr(), self.P, self.N, self.dim, q["px_th"],
                                        q["conf"], q["max_iters"], q.get("error_type", 0), True, 0.0,
                                        q.get("degen", True), self.dseed.data_ptr(), self.model.data_ptr(),
                                        self.mask.data_ptr(), self.stats.data_ptr(), stream)
        else:
            _cabi.homography_batch_dev(self.d1.data_ptr(), self.d2.data_ptr(), self.P, self.N, self.dim, q["px_th"],
                                       q["conf"], q["max_iters"], q.get("error_type", 0), True, 0.0,
                                       self.dseed.data_ptr(), self.model.data_ptr(), self.mask.data_ptr(),
                                       self.stats.data_ptr(), stream)

    def gather(self):
        """records of every rank on rank 0 (device): the path's only collective."""
        import torch
        self.rec[:, :72] = self.model.view(torch.uint8).view(self.P, 72)
        self.rec[:, 72:88] = self.stats.view(torch.uint8).view(self.P, 16)
        self.rec[:, 88:] = self.mask
        if self.dist is not None:
            self.dist.gather(self.rec, gather_list=self.parts, dst=0)
        elif self.rank == 0:
            self.gathered[0].copy_(self.rec)

    def step(self, h1, h2, hseeds):
        """One end-to-end step: pinned host block -> device -> kernel -> gather -> rank 0 reads the records back."""
        import torch
        stream = torch.cuda.current_stream(self.dev)
        self.d1.copy_(h1, non_blocking=True)
        self.d2.copy_(h2, non_blocking=True)
        self.dseed.copy_(hseeds, non_blocking=True)
        self.launch(stream.cuda_stream)
        self.gather()
        if self.rank == 0:
            self.h_out.copy_(self.gathered.view(self.world * self.P, self.stride), non_blocking=True)
            stream.synchronize()
            return self.h_out.numpy()
        stream.synchronize()
        return None
