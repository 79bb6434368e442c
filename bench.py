#!/usr/bin/env python
"""bench.py -- image-pairs/sec of findFundamentalMatrix (2000 correspondences, 30 % inliers, 10k iterations)
on N B200s vs the reference on the host CPU.  One JSON line on stdout (rank 0).

    python bench.py --gpus 1 --steps 5 --warmup 3
    python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 --master-port P \
        bench.py --gpus N --steps K --warmup W
    python bench.py --impl reference --gpus 1 --steps 3 --warmup 1      # the reference's own CPU path

A step = one pass of the hot path over one batch of `--pairs-per-gpu` synthetic image pairs per GPU
(scene F(2000, 0.30, seed=s), distinct data per pair; BASELINE.json configs[1]/[4]).
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_CORR = 2000
INLIER_RATIO = 0.30
PX_TH = 1.0
CONF = 0.9999
MAX_ITERS = 10000
METRIC = "image-pairs/sec (2000 corr, 10k iters) F-matrix"
ALGO_BYTES_PER_PAIR = N_CORR * 4 * 8 + 72 + N_CORR + 16      # 66 088 B
REF_EQUIV_FLOP_PER_PAIR = 1.44e8                               # 1 940 passes x 2000 x 37


def gen_batch(n_pairs, seed0):
    from pydegensac_b200.scenes import batch_F
    return batch_F(n_pairs, N_CORR, INLIER_RATIO, seed0)


# ----------------------------------------------------------------------------- CPU reference arm
def _cpu_worker(args):
    seed0, count = args
    os.environ["OPENBLAS_NUM_THREADS"] = "1"
    from oracle import ref
    from pydegensac_b200.scenes import scene_F
    scenes = [scene_F(N_CORR, INLIER_RATIO, seed0 + i) for i in range(count)]
    ref.find_fundamental(scenes[0][0], scenes[0][1], PX_TH, CONF, 100, seed=1, rng=ref.RNG_GLIBC)  # load/warm
    t = time.perf_counter()
    inl = 0
    for i, (p1, p2, _) in enumerate(scenes):
        F, m, st = ref.find_fundamental(p1, p2, PX_TH, CONF, MAX_ITERS, degen_check=True, seed=seed0 + i,
                                        rng=ref.RNG_GLIBC)
        inl += int(m.sum())
    return time.perf_counter() - t, inl


def cpu_reference_rate(pairs_per_proc, procs, seed0=100000):
    """Unmodified reference (oracle/_ref, glibc RNG) on `procs` processes, one per core (the library is not
    thread-safe: global hash table + libc RNG). Returns (pairs/s over the pool, single-process pairs/s)."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    jobs = [(seed0 + i * pairs_per_proc, pairs_per_proc) for i in range(procs)]
    with ctx.Pool(procs) as pool:
        pool.map(_cpu_worker, [(seed0, 1)] * procs)          # spawn + import + warm-up, untimed
        t = time.perf_counter()
        res = pool.map(_cpu_worker, jobs, chunksize=1)
        wall = time.perf_counter() - t
    busy = [r[0] for r in res]
    total = pairs_per_proc * procs
    return total / max(max(busy), 1e-9), pairs_per_proc / (sum(busy) / len(busy)), wall


def host_cores():
    """Cores this process may actually use: CPU affinity capped by the cgroup CPU quota (a container that sees
    128 logical CPUs but has cpu.max = 16 CPUs gets 16 worker processes, not 128 oversubscribed ones)."""
    try:
        n = len(os.sched_getaffinity(0))
    except Exception:
        n = os.cpu_count() or 1
    for path in ("/sys/fs/cgroup/cpu.max", "/sys/fs/cgroup/cpu/cpu.cfs_quota_us"):
        try:
            txt = open(path).read().split()
            if path.endswith("cpu.max"):
                if txt[0] != "max":
                    n = min(n, max(1, int(float(txt[0]) / float(txt[1]) + 0.5)))
            else:
                q = int(txt[0])
                if q > 0:
                    per = int(open("/sys/fs/cgroup/cpu/cpu.cfs_period_us").read())
                    n = min(n, max(1, int(q / per + 0.5)))
            break
        except Exception:
            continue
    return n


def run_reference_arm(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    from oracle import ref
    cores = host_cores()
    if not ref.available():
        print(json.dumps({"impl": "reference", "unavailable": "oracle/_ref/libdegensac_ref.so not built"}))
        return
    per = max(1, args.cpu_pairs_per_core)
    vals = []
    for s in range(args.warmup + args.steps):
        rate, single, wall = cpu_reference_rate(per, cores, seed0=100000 + s * per * cores)
        if s >= args.warmup:
            vals.append((rate, single, wall))
    rate = float(np.mean([v[0] for v in vals]))
    single = float(np.mean([v[1] for v in vals]))
    ms = float(np.mean([v[2] for v in vals])) * 1e3
    line = {
        "impl": "reference", "metric": METRIC, "value": rate, "unit": "pairs/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "findFundamentalMatrix scene F(2000,0.30,seed), px_th 1.0, conf 0.9999, max_iters 10000, "
                               "sampson, sym check on, degeneracy check on", "pairs_per_step": per * cores},
        "cpu_baseline": {"value": rate, "unit": "pairs/s", "cores": cores, "kind": "reference",
                         "sample": "%d pairs per process x %d processes (one per core), unmodified reference C core "
                                   "with its own glibc RNG; single-process rate %.1f pairs/s" % (per, cores, single)},
        "e2e": {"value": rate, "unit": "pairs/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    print(json.dumps(line))


# ----------------------------------------------------------------------------- GPU arm
class ClockSampler(threading.Thread):
    def __init__(self, index):
        super().__init__(daemon=True)
        self.index = index
        self.samples = []
        self.reasons = set()
        self.stop_flag = False
        self.max_mhz = None

    def run(self):
        q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
             "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        while not self.stop_flag:
            try:
                o = subprocess.run(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + q,
                                    "--format=csv,noheader,nounits"], capture_output=True, text=True, timeout=5).stdout
                f = [x.strip() for x in o.strip().split(",")]
                self.samples.append(float(f[0]))
                self.max_mhz = float(f[1])
                for nme, v in zip(names, f[2:6]):
                    if v.lower().startswith("active"):
                        self.reasons.add(nme)
            except Exception:
                pass
            time.sleep(0.2)


def run_gpu_arm(args):
    import torch
    from pydegensac_b200 import _cabi
    import pydegensac_b200 as pdg

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device - the engine has no CPU path (use --impl reference for the CPU arm)")
    torch.cuda.set_device(local_rank)
    _cabi.lib().dgb200_set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)

    P = args.pairs_per_gpu
    p1, p2 = gen_batch(P, seed0=rank * P)
    seeds = (np.arange(P, dtype=np.uint64) + np.uint64(rank * P))
    # device-resident inputs / outputs for `value`
    d1 = torch.from_numpy(p1).to(dev)
    d2 = torch.from_numpy(p2).to(dev)
    dseed = torch.from_numpy(seeds.astype(np.int64)).to(dev)
    dF = torch.zeros((P, 9), dtype=torch.float64, device=dev)
    dmask = torch.zeros((P, N_CORR), dtype=torch.uint8, device=dev)
    dstats = torch.zeros((P, 4), dtype=torch.int32, device=dev)
    rec_stride = 72 + 16 + N_CORR
    flush = torch.empty(256 * 1024 * 1024, dtype=torch.uint8, device=dev)   # > 126 MB L2
    stream = torch.cuda.current_stream()

    def step_device():
        _cabi.fundamental_batch_dev(d1.data_ptr(), d2.data_ptr(), P, N_CORR, 2, PX_TH, CONF, MAX_ITERS, 0, True, 0.0,
                                    True, dseed.data_ptr(), dF.data_ptr(), dmask.data_ptr(), dstats.data_ptr(),
                                    stream.cuda_stream)
        if world > 1:   # final gather of (F, stats, mask) records on rank 0: the path's only collective
            rec = torch.cat([dF.view(torch.uint8).view(P, 72), dstats.view(torch.uint8).view(P, 16), dmask], 1)
            parts = [torch.empty_like(rec) for _ in range(world)] if rank == 0 else None
            dist.gather(rec, gather_list=parts, dst=0)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    for _ in range(args.warmup):
        step_device()
    barrier()
    sampler = ClockSampler(local_rank)
    sampler.start()
    launches0 = _cabi.kernel_launches()
    times = []
    barrier()
    for _ in range(args.steps):
        flush.fill_(1)                       # evict L2 between timed iterations (not timed)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        step_device()
        e1.record(stream)
        torch.cuda.synchronize()
        times.append(e0.elapsed_time(e1))
    barrier()
    launches = _cabi.kernel_launches() - launches0
    sampler.stop_flag = True
    total_ms = float(sum(times))
    tt = torch.tensor([total_ms], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(tt, op=dist.ReduceOp.MAX)
    total_ms = float(tt.item())
    value = world * P * args.steps / (total_ms / 1e3)
    # kernel-only time of the dominant (only) kernel, rank 0: events around the launch alone
    kt = []
    for _ in range(max(3, args.steps)):
        flush.fill_(1)
        torch.cuda.synchronize()
        e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        _cabi.fundamental_batch_dev(d1.data_ptr(), d2.data_ptr(), P, N_CORR, 2, PX_TH, CONF, MAX_ITERS, 0, True, 0.0,
                                    True, dseed.data_ptr(), dF.data_ptr(), dmask.data_ptr(), dstats.data_ptr(),
                                    stream.cuda_stream)
        e1.record(stream)
        torch.cuda.synchronize()
        kt.append(e0.elapsed_time(e1))
    kernel_ms = float(np.mean(kt))

    # e2e: public batched API with HOST buffers (pinned), H2D + kernel + D2H inside the timed region
    hp1 = torch.from_numpy(p1).pin_memory().numpy()
    hp2 = torch.from_numpy(p2).pin_memory().numpy()
    pdg.findFundamentalMatrixBatch(hp1[:64], hp2[:64], PX_TH, CONF, MAX_ITERS, seeds=seeds[:64])
    barrier()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        F_h, mask_h = pdg.findFundamentalMatrixBatch(hp1, hp2, PX_TH, CONF, MAX_ITERS, seeds=seeds)
        if world > 1:
            from pydegensac_b200.parallel import pack_records, gather_records
            gather_records(pack_records(F_h, mask_h, np.zeros((P, 4), np.int32)), world * P, dist, dev)
    torch.cuda.synchronize()
    e2e_s = time.perf_counter() - t0
    te = torch.tensor([e2e_s], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(te, op=dist.ReduceOp.MAX)
    e2e_value = world * P * args.steps / float(te.item())
    mean_inl = float(mask_h.sum(1).mean())

    if rank == 0:
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        hbm_peak = float(peaks.get("hbm_gbs", 6650.0))
        peak_src = "measured (MEASURED_PEAKS.json)" if "hbm_gbs" in peaks else "fallback (B200_PROFILING.md)"
        achieved = ALGO_BYTES_PER_PAIR * P / (kernel_ms / 1e3) / 1e9
        traffic = None
        tf = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tf):
            try:
                traffic = json.load(open(tf)).get("dram_bytes_per_launch")
            except Exception:
                traffic = None
        cpu = None
        if not args.no_cpu_baseline:
            try:
                from oracle import ref
                if ref.available():
                    # timed in a FRESH interpreter (the reference arm with one step): forking 16 workers out of
                    # this process -- CUDA context, pinned staging buffers, sampler thread -- halves their speed
                    out = subprocess.run([sys.executable, os.path.abspath(__file__), "--impl", "reference", "--gpus", "1",
                                          "--steps", "1", "--warmup", "1", "--cpu-pairs-per-core",
                                          str(args.cpu_pairs_per_core)], capture_output=True, text=True, timeout=900,
                                         env={k: v for k, v in os.environ.items()
                                              if k not in ("RANK", "LOCAL_RANK", "WORLD_SIZE")})
                    ref_line = json.loads(out.stdout.strip().splitlines()[-1])
                    cpu = ref_line["cpu_baseline"]
                    cpu["sample"] += " (separate process, 1 timed step after 1 warm-up step)"
            except Exception as ex:  # pragma: no cover
                cpu = {"value": None, "unit": "pairs/s", "cores": 0, "kind": "reference", "sample": "failed: %r" % (ex,)}
        sm = sorted(sampler.samples)
        line = {
            "metric": METRIC, "value": value, "unit": "pairs/s", "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": total_ms / args.steps, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": "findFundamentalMatrix batch: scene F(2000,0.30,seed=s) per pair, px_th 1.0, "
                                   "conf 0.9999, max_iters 10000, sampson, sym check on, degeneracy check on",
                       "pairs_per_gpu": P, "global_pairs": world * P, "parallelism": "pairs sharded x%d, final gather" % world,
                       "l2": "256 MB flush write between timed iterations", "mean_inliers": mean_inl},
            "e2e": {"value": e2e_value, "unit": "pairs/s", "h2d_bytes_per_step": int(2 * P * N_CORR * 2 * 8 + P * 8),
                    "d2h_bytes_per_step": int(P * (72 + N_CORR + 16))},
            "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "achieved": achieved, "peak": hbm_peak, "unit": "GB/s",
                         "frac": achieved / hbm_peak, "traffic": traffic, "peak_source": peak_src,
                         "kernel": "ransac_pairs_kernel<F>", "kernel_ms": kernel_ms,
                         "note": "path is FP64-issue/latency bound, not HBM bound (SURVEY.md §8(d)); "
                                 "reference-equivalent FP64 work %.3g flop/pair -> %.2f TFLOP/s achieved"
                                 % (REF_EQUIV_FLOP_PER_PAIR, REF_EQUIV_FLOP_PER_PAIR * P / (kernel_ms / 1e3) / 1e12)},
            "cpu_baseline": cpu,
            "clocks": {"sm_mhz": sm[len(sm) // 2] if sm else None, "sm_max_mhz": sampler.max_mhz,
                       "reasons": sorted(sampler.reasons)},
        }
        print(json.dumps(line))
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=5)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--pairs-per-gpu", type=int, default=16384)
    ap.add_argument("--cpu-pairs-per-core", type=int, default=16)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference_arm(args)
    else:
        run_gpu_arm(args)


if __name__ == "__main__":
    main()
