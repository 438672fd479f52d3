#!/usr/bin/env python
"""Benchmark of the infer_pa hot path (BASELINE.json metric: infer_pa UTRs/s and
read*comp*EM-iter/s, next to the host-CPU reference).

  python bench.py [--gpus N] [--steps K] [--warmup W]            this repo's CUDA path
  python bench.py --impl reference [--steps K] [--warmup W]      CPU arm: the oracle port on all host cores

A step = one pass of the whole path (binning -> theta table -> marginal tensor -> 50+ EM chains per
UTR -> BIC selection / pruning / re-run -> labels) over one batch of synthetic UTRs.
Workload at N=1: BASELINE.json configs[1], 10,000 UTRs x 500 reads, Kmax=5, 100 chunk files (= 100
RNG streams, seed 1 each, like one `scape infer_pa` per file).  N>1 (torchrun): every rank fits its
own 10,000-UTR set of the same shape (weak scaling; UTRs are independent, no data-path collective).

  value  = UTRs / device time, device time = time the GPU spent executing the library's kernels
           (union of the CUDA-event kernel intervals on the library's streams) -- no host work counted
  e2e    = UTRs / wall time of the public call scape_b200.apa_core.fit_chunks-equivalent
           (Engine.fit on HOST read columns: host binning + RNG replay + H2D + kernels + D2H of
           results and per-read labels)
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

N_UTR = 10000
READS = 500
PER_FILE = 100


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        with open(p) as fh:
            d = json.load(fh)
        return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def scan_traffic():
    """DRAM bytes per em_scan_kernel launch from the committed ncu capture (profiles/), or None."""
    import glob
    found = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_scan_traffic.json")))
    if not found:
        return None, "no ncu capture committed"
    p = found[-1]                      # the latest capture (names sort by round)
    with open(p) as fh:
        d = json.load(fh)
    return float(d["dram_bytes_per_launch"]), d.get("source", "profiles/r01_scan_traffic.json")


class ClockSampler:
    """nvidia-smi clocks / throttle reasons during the timed region (B200_PROFILING.md recipe)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            threading.Thread(target=self._pump, daemon=True).start()
        except Exception:
            self.proc = None

    def _pump(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc:
            self.proc.terminate()
        sm = [float(r[0]) for r in self.rows if r and r[0].replace(".", "").isdigit()]
        mx = [float(r[1]) for r in self.rows if len(r) > 1 and r[1].replace(".", "").isdigit()]
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        reasons = [n for i, n in enumerate(names) if any(len(r) > 3 + i and r[3 + i] == "Active" for r in self.rows)]
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm)}


def pack(utrs):
    off = np.zeros(len(utrs) + 1, np.int64)
    np.cumsum([u.n_reads for u in utrs], out=off[1:])
    cat = lambda k: np.concatenate([np.asarray(getattr(u, k), dtype=np.float64) for u in utrs])
    return off, cat("x"), cat("l"), cat("r"), cat("pa")


def cfg3_plan(total, per_file, world):
    """cfg-3: reads per UTR, the long-UTR cut, the chunk files (lists of UTR indices) and which files
    every rank fits (LPT packing of the a-priori stream costs, scape_b200/shard.py)."""
    from scape_b200 import shard, synth
    counts = synth.heavy_tail_read_counts(total)
    cut = np.quantile(synth.heavy_tail_read_counts(max(total, 1000)), 0.99)
    files = [list(range(f, min(f + per_file, total))) for f in range(0, total, per_file)]
    costs = shard.stream_costs([[counts[i] for i in f] for f in files],
                               [[20000 if counts[i] >= cut else 2000 for i in f] for f in files])
    return counts, cut, files, shard.lpt_partition(costs, world)


def _cpu_fit_one(args):
    idx, reads = args[0], args[1]
    from oracle import scape_oracle as so
    from scape_b200 import synth
    u = synth.make_utr(idx, reads, long_utr=bool(args[2]) if len(args) > 2 else False)
    t = time.perf_counter()
    res = so.fit_utr(u.x, u.l, u.r, u.pa, np.random.RandomState(1))
    return time.perf_counter() - t, res.n_frag, res.K


def cpu_sample(first, count, reads, cores, jobs=None):
    """Oracle port on `cores` processes over UTRs [first, first+count) of the workload (or the explicit
    (index, reads, long_utr) job list); wall clock."""
    import multiprocessing as mp
    ctx = mp.get_context("fork")
    t0 = time.perf_counter()
    with ctx.Pool(cores) as pool:
        rows = pool.map(_cpu_fit_one, jobs or [(first + i, reads) for i in range(count)], chunksize=1)
    wall = time.perf_counter() - t0
    return count / wall, wall, rows


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    per_step = cores * 2
    for w in range(args.warmup):
        cpu_sample(N_UTR - per_step, min(per_step, cores), READS, cores)
    t0 = time.perf_counter()
    for s in range(args.steps):
        cpu_sample(s * per_step, per_step, READS, cores)
    wall = time.perf_counter() - t0
    v = args.steps * per_step / wall
    sample = f"{per_step} UTRs/step (UTR indices s*{per_step}..) of the 10000 x 500-read workload, fresh RandomState(1) per UTR"
    print(json.dumps({
        "impl": "reference", "metric": "infer_pa_utrs_per_s", "value": v, "unit": "UTR/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": "cfg-2: synthetic 10k UTRs x 500 reads, Kmax=5 (bounded sample per step)"},
        "cpu_baseline": {"value": v, "unit": "UTR/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "UTR/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle/scape_oracle.py (vectorised numpy restatement, bit-identical to the reference's results and "
                "~6x faster than the reference's own Python loops) because the Python/Taichi reference cannot travel",
    }))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--utrs", type=int, default=N_UTR, help="UTRs per GPU per step (default: the named config)")
    ap.add_argument("--reads", type=int, default=READS)
    ap.add_argument("--per-file", type=int, default=PER_FILE)
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configs[1..4]; cfg2 is the benchmark line, the others are reported beside it")
    ap.add_argument("--kmax", type=int, default=5, help="n_max_apa (cfg5 sweeps it)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    from scape_b200 import _lib, shard, synth
    scaling = "weak"
    pre_para = None
    if args.workload == "cfg2":
        utrs = synth.make_batch(args.utrs, args.reads, first=rank * args.utrs)
        label = (f"cfg-2: synthetic {args.utrs} UTRs x {args.reads} reads per GPU, Kmax={args.kmax}, "
                 f"{(args.utrs + args.per_file - 1) // args.per_file} chunk files = RNG streams (seed 1 each)")
    elif args.workload in ("cfg3", "cfg4"):
        # 20k heavy-tailed UTRs in TOTAL (strong scaling): chunk files are bin-packed over the ranks
        # by the a-priori cost model (scape_b200/shard.py), every rank generates and fits only its own
        total = args.utrs if args.utrs != N_UTR else 20000
        counts, cut, files, parts = cfg3_plan(total, args.per_file, world)
        mine = parts[rank]
        utrs = [synth.make_utr(i, int(counts[i]), long_utr=bool(counts[i] >= cut)) for f in mine for i in files[f]]
        scaling = "strong"
        label = (f"cfg-3: synthetic {total} UTRs in total, heavy-tailed reads per UTR (10..200k, median 600), "
                 f"{len(files)} chunk files cost-balanced (LPT) over {world} GPU(s), Kmax={args.kmax}")
        if args.workload == "cfg4":
            class _Pre:           # what --pre_para_pkl_file supplies: first Parameters object of the file
                alpha_arr = np.array([500, 900, 1400]); beta_arr = np.array([20.0, 35.0, 30.0]); K = 3; L = 21000
            pre_para = _Pre
            label = label.replace("cfg-3", "cfg-4 (pre_para fixed mode, K=3, restricted theta grid)")
        args.utrs = len(utrs)
    else:
        # cfg-5 point: 256 independent UTRs (one stream each, so all of them are one wave)
        args.utrs = min(args.utrs, 256)
        args.per_file = 1
        utrs = synth.make_batch(args.utrs, args.reads, first=rank * args.utrs)
        label = f"cfg-5 point: {args.utrs} UTRs x {args.reads} reads, n_max_apa={args.kmax}, one stream per UTR"
    off, x, l, r, pa = pack(utrs)
    n_files = (args.utrs + args.per_file - 1) // args.per_file
    sid = (np.arange(args.utrs) // args.per_file).astype(np.int32)
    seeds = np.ones(n_files, np.uint32)
    eng = _lib.Engine(_lib.make_params(pre_para=pre_para, n_max_apa=args.kmax), device=local)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(args.warmup):
        eng.fit(off, x, l, r, pa, sid, seeds)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    dev_ms, em_ms, tim = 0.0, 0.0, None
    acc = {}
    for _ in range(args.steps):
        out = eng.fit(off, x, l, r, pa, sid, seeds)
        tim = out.timing
        for k, v in tim.items():
            acc[k] = acc.get(k, 0.0) + v
    barrier()
    wall = time.perf_counter() - t0
    clocks = sampler.stop() if rank == 0 else None
    # One extra, untimed pass with the wave pipelining off: every kernel then has the GPU to itself,
    # which is what the per-kernel roofline needs (in the timed passes the next wave's table/tensor
    # kernels share the SMs with the EM kernels, so their event times overlap).
    alone = None
    if rank == 0:
        eng.set_overlap(False)
        alone = eng.fit(off, x, l, r, pa, sid, seeds).timing
        eng.set_overlap(True)
    dev_ms = acc["device_busy_ms"]        # union of the kernel intervals of all lanes (CUDA events)
    work = float(out.em_work[:, 0].sum())          # sum N*(K+1) over chains and iterations, one step
    iters = float(out.em_work[:, 1].sum())
    if dist is not None:
        import torch
        t = torch.tensor([wall, dev_ms, acc["em_ms"]], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, dev_ms, em_max = t.tolist()
        s = torch.tensor([work, float(args.utrs)], dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        work_all, utr_all = s.tolist()
    else:
        work_all, utr_all = work, float(args.utrs)
    if rank != 0:
        if dist is not None:
            dist.destroy_process_group()
        return

    K = args.steps
    hbm_peak, hbm_src = measured_peaks()
    fp64 = eng.fp64_peaks()                     # measured on this GPU, this run (no FP64 entry in MEASURED_PEAKS.json)
    cluster = alone.get("cluster_ms", 0.0) > alone["scan_ms"]     # which EM kernel dominates this workload
    if cluster:
        scan_s = alone["cluster_ms"] / 1e3
        ach_tflops = alone["cluster_grid_flops"] / scan_s / 1e12
        dom_launches = max(int(alone["cluster_launches"]), 1)
        dom_flops = alone["cluster_grid_flops"]
    else:
        scan_s = alone["scan_ms"] / 1e3
        ach_tflops = alone["em_grid_flops"] / scan_s / 1e12
        dom_launches = max(int(alone["scan_launches"]), 1)
        dom_flops = alone["em_grid_flops"]
    traffic, traffic_src = scan_traffic()
    res = {
        "metric": "infer_pa_utrs_per_s", "value": utr_all * K / (dev_ms / 1e3), "unit": "UTR/s",
        "n_gpus": world, "steps": K, "warmup": args.warmup, "ms_per_step": dev_ms / K,
        "higher_is_better": True, "scaling": scaling, "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": label,
                   "seed_policy": "file", "tensor_storage": "f32 (FP64 arithmetic)",
                   "l2": "per-wave tensor working set (~0.26 GB) exceeds the 126 MB L2; nothing is flushed between steps, "
                         "every step streams 100 waves x 0.26 GB",
                   "value_time": "GPU busy time = union of the CUDA-event kernel intervals (EM stream and the "
                                 "likelihood stream that works one wave ahead)",
                   "e2e_time": "wall clock of Engine.fit (C ABI scape_b200_fit_batch) on host buffers"},
        "read_comp_em_iter_per_s": work_all * K / (dev_ms / 1e3),
        "read_comp_em_iter_per_s_e2e": work_all * K / wall,
        "em_iterations_per_step": iters,
        "e2e": {"value": utr_all * K / wall, "unit": "UTR/s", "h2d_bytes_per_step": acc["h2d_bytes"] / K,
                "d2h_bytes_per_step": acc["d2h_bytes"] / K, "ms_per_step": 1e3 * wall / K},
        "gpu_launches": int(acc["launches"]),
        "roofline": {"kernel": ("em_cluster_kernel (cluster-resident EM: E passes + max_alpha_beta grid arg-max as FP64 MMA tiles, "
                                "all iterations of a UTR in one launch; only the grid search's flops are counted)") if cluster else
                               "em_scan_kernel (max_alpha_beta grid arg-max as a blocked FP64 MMA product)",
                     "bound": "tensor", "achieved": ach_tflops, "peak": fp64["dmma_tflops"], "unit": "TFLOP/s",
                     "frac": ach_tflops / fp64["dmma_tflops"], "traffic": traffic, "traffic_source": traffic_src,
                     "timing": "CUDA events around every scan launch of one extra pass with wave pipelining off "
                               "(kernel alone on the GPU)",
                     "peak_source": "FP64 mma.m8n8k4 stream measured in this run (scape_b200_fp64_peaks); "
                                    f"CUDA-core DFMA stream {fp64['dfma_tflops']:.1f} TFLOP/s",
                     "algorithmic_flops_per_launch": dom_flops / dom_launches,
                     "avg_launch_ms": 1e3 * scan_s / dom_launches,
                     "launches": dom_launches,
                     "share_of_gpu_time": 1e3 * scan_s / alone["device_busy_ms"],
                     "hbm_view": {"algorithmic_GBps": alone["em_grid_bytes"] / scan_s / 1e9, "peak_GBps": hbm_peak,
                                  "peak_source": hbm_src, "loaded_GBps": alone["em_scan_bytes"] / scan_s / 1e9,
                                  "note": "SURVEY 8d algorithmic bytes = 8*W_k*B*N per chain iteration; the blocked scan "
                                          "loads each tensor block once per step for all chains of the UTR"}},
        "phases_ms_per_step": {k: acc[k] / K for k in ("table_ms", "tensor_ms", "em_ms", "cluster_ms", "estep_ms", "scan_ms", "label_ms",
                                                        "host_prep_ms", "host_rng_ms", "device_busy_ms", "total_ms")},
        "phases_alone_ms": {k: alone[k] for k in ("table_ms", "tensor_ms", "em_ms", "cluster_ms", "estep_ms", "scan_ms", "label_ms",
                                                  "device_busy_ms", "total_ms")},
        "tensor_exp_per_s": alone["tensor_exp"] / (alone["tensor_ms"] / 1e3),
        "waves_per_step": acc["waves"] / K,
        "clocks": clocks,
    }
    if not args.no_cpu:
        cores = os.cpu_count() or 1
        n = cores * 2
        if args.workload == "cfg3":
            # bounded sample: the first 2*cores UTRs of the set with at most 5000 reads (a 100k-read UTR
            # alone takes the oracle minutes); it therefore flatters the CPU arm on this workload
            jobs = [(i, int(c), bool(c >= cut)) for i, c in enumerate(counts) if c <= 5000][:n]
            v, wall_cpu, rows = cpu_sample(0, n, args.reads, cores, jobs=jobs)
            sample = f"first {len(jobs)} UTRs with <= 5000 reads of the heavy-tailed set, oracle port, {wall_cpu:.1f} s wall"
        elif args.workload == "cfg2":
            v, wall_cpu, rows = cpu_sample(0, n, args.reads, cores)
            sample = f"first {n} UTRs of the workload, oracle port, {wall_cpu:.1f} s wall"
        else:
            v = None
        if v is not None:
            res["cpu_baseline"] = {"value": v, "unit": "UTR/s", "cores": cores, "kind": "port", "sample": sample}
    print(json.dumps(res))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
