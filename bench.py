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


def scan_traffic(cluster=False):
    """DRAM bytes per launch of the dominant EM kernel from the committed ncu capture (profiles/), or None."""
    import glob
    found = sorted(glob.glob(os.path.join(ROOT, "profiles", "r*_cluster_traffic.json" if cluster else "r*_scan_traffic.json")))
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


def workload_label(args, world):
    """`config.workload` of the benchmark line: the same string for both arms (--impl b200 / reference)."""
    if args.workload == "cfg2":
        return (f"cfg-2: synthetic {args.utrs} UTRs x {args.reads} reads per GPU, Kmax={args.kmax}, "
                f"{(args.utrs + args.per_file - 1) // args.per_file} chunk files = RNG streams (seed 1 each)")
    if args.workload in ("cfg3", "cfg4"):
        total = args.utrs if args.utrs != N_UTR else 20000
        label = (f"cfg-3: synthetic {total} UTRs in total, heavy-tailed reads per UTR (10..200k, median 600), "
                 f"{(total + args.per_file - 1) // args.per_file} chunk files cost-balanced (LPT) over {world} GPU(s), Kmax={args.kmax}")
        if args.workload == "cfg4":
            label = label.replace("cfg-3", "cfg-4 (pre_para fixed mode, K=3, restricted theta grid)")
        return label
    return f"cfg-5 point: {min(args.utrs, 256)} UTRs x {args.reads} reads, n_max_apa={args.kmax}, one stream per UTR"


# ---- CPU arm: the oracle port on whole chunk-file prefixes under the reference's seed policy ----------
def _cpu_fit_file(job):
    """First `m` UTRs of chunk file `f` of the cfg-2 set, np.random.seed(1) once per file, UTRs serial on
    that stream (apa_core.py:125, 1104-1137).  Returns per-UTR (index, K, alpha, ws, lb_last, n_frag)."""
    f, m, per_file, reads, kmax = job
    from oracle import scape_oracle as so
    from scape_b200 import synth
    rng = np.random.RandomState(1)
    rows = []
    for j in range(m):
        u = synth.make_utr(f * per_file + j, reads)
        res = so.fit_utr(u.x, u.l, u.r, u.pa, rng, n_max_apa=kmax)
        rows.append((f * per_file + j, int(res.K), [int(a) for a in res.alpha_arr], [float(b) for b in res.beta_arr],
                     [float(w) for w in res.ws], float(res.lb_arr[-1]), int(res.n_frag)))
    return rows


class CpuArm:
    """One persistent process pool (fork, before any CUDA context exists in this process) fitting
    `m` leading UTRs of `cores` chunk files per step."""

    def __init__(self, cores):
        import multiprocessing as mp
        import warnings
        warnings.simplefilter("ignore")
        self.cores = cores
        self.pool = mp.get_context("fork").Pool(cores)

    def step(self, first_file, m, args):
        n_files = (args.utrs + args.per_file - 1) // args.per_file
        jobs = [((first_file + i) % n_files, m, args.per_file, args.reads, args.kmax) for i in range(self.cores)]
        t0 = time.perf_counter()
        rows = self.pool.map(_cpu_fit_file, jobs, chunksize=1)
        return time.perf_counter() - t0, [r for file_rows in rows for r in file_rows]

    def close(self):
        self.pool.close()
        self.pool.join()


def run_reference(args):
    """--impl reference: the reference's CPU implementation of the path (the oracle port: the Python /
    Taichi reference cannot travel to the GPU box) on all host cores, same workload string."""
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if rank != 0:
        return
    cores = os.cpu_count() or 1
    m = args.ref_utrs_per_file
    arm = CpuArm(cores)
    for w in range(args.warmup):
        arm.step(w * cores, 1, args)
    wall = 0.0
    for s in range(args.steps):
        dt, _ = arm.step((args.warmup + s) * cores, m, args)
        wall += dt
    arm.close()
    v = args.steps * cores * m / wall
    sample = (f"per step: the first {m} UTRs of {cores} chunk files of the workload (distinct files per step), each file on its "
              f"own process with np.random.seed(1) per file and its UTRs serial on that stream; persistent pool of {cores} processes")
    print(json.dumps({
        "impl": "reference", "metric": "infer_pa_utrs_per_s", "value": v, "unit": "UTR/s", "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / args.steps,
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": workload_label(args, world), "seed_policy": "file"},
        "cpu_baseline": {"value": v, "unit": "UTR/s", "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": "UTR/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "note": "oracle/scape_oracle.py (vectorised numpy restatement, bit-identical to the reference's results and "
                "~6x faster than the reference's own Python loops) because the Python/Taichi reference cannot travel; "
                "the CPU arm does not use the GPUs, so its value does not change with --gpus",
    }))


def files_leg(utrs, per_file, local):
    """The path a `scape infer_pa` user runs, file to file: `prepare_input`-format chunk pickles on disk ->
    scape_b200.apa_core.infer_files (unpickle the DataFrames, pack CSR columns, fit_batch, build and
    pickle the Parameters objects) -> result pickles on disk (apa_core.py:1104-1137 per file).  One
    warm-up call (worker processes, page cache, engine arenas), one timed call."""
    import shutil
    import tempfile
    from scape_b200 import apa_core, synth
    root = tempfile.mkdtemp(prefix="scape_bench_", dir="/dev/shm" if os.path.isdir("/dev/shm") else None)
    try:
        paths = synth.write_chunk_files(utrs, root, per_file=per_file)
        in_bytes = sum(os.path.getsize(p) for p in paths)
        apa_core.infer_files(paths, root, device=local)
        t0 = time.perf_counter()
        outs = apa_core.infer_files(paths, root, device=local)
        wall = time.perf_counter() - t0
        out_bytes = sum(os.path.getsize(o) for o in outs)
    finally:
        shutil.rmtree(root, ignore_errors=True)
    return {"value": len(utrs) / wall, "unit": "UTR/s", "wall_s": wall, "chunk_files": len(paths), "input_bytes": in_bytes,
            "output_bytes": out_bytes,
            "path": "chunk pickles (tmpfs) -> infer_files: worker processes unpickle / pack, one fit_batch call, worker "
                    "processes build and pickle scape.apa_core.Parameters -> result pickles"}


def make_workload(name, args, rank, world):
    from scape_b200 import synth
    a = argparse.Namespace(**vars(args))
    a.workload = name
    wl = dict(name=name, scaling="weak", pre_para=None, per_file=args.per_file, kmax=args.kmax)
    if name == "cfg2":
        wl["utrs"] = synth.make_batch(args.utrs, args.reads, first=rank * args.utrs)
    elif name in ("cfg3", "cfg4"):
        # 20k heavy-tailed UTRs in TOTAL (strong scaling): chunk files are bin-packed over the ranks
        # by the a-priori cost model (scape_b200/shard.py), every rank generates and fits only its own
        total = args.utrs if args.utrs != N_UTR else 20000
        counts, cut, files, parts = cfg3_plan(total, args.per_file, world)
        wl["utrs"] = [synth.make_utr(i, int(counts[i]), long_utr=bool(counts[i] >= cut)) for f in parts[rank] for i in files[f]]
        wl["scaling"] = "strong"
        if name == "cfg4":
            class _Pre:           # what --pre_para_pkl_file supplies: first Parameters object of the file
                alpha_arr = np.array([500, 900, 1400]); beta_arr = np.array([20.0, 35.0, 30.0]); K = 3; L = 21000
            wl["pre_para"] = _Pre
    else:
        # cfg-5 point: 256 independent UTRs (one stream each, so all of them are one wave)
        wl["per_file"] = 1
        wl["utrs"] = synth.make_batch(min(args.utrs, 256), args.reads, first=rank * min(args.utrs, 256))
    wl["label"] = workload_label(a, world)
    return wl


def run_workload(wl, args, steps, warmup, rank, local, world, dist, cpu_arm=None, files=False):
    """W warm-up passes, K timed passes (barrier + synchronize on both sides), one extra un-pipelined
    pass for the per-kernel roofline.  Returns the result dict on rank 0, None elsewhere."""
    from scape_b200 import _lib
    utrs = wl["utrs"]
    n_utr = len(utrs)
    off, x, l, r, pa = pack(utrs)
    n_files = (n_utr + wl["per_file"] - 1) // wl["per_file"]
    sid = (np.arange(n_utr) // wl["per_file"]).astype(np.int32)
    seeds = np.ones(n_files, np.uint32)
    eng = _lib.Engine(_lib.make_params(pre_para=wl["pre_para"], n_max_apa=wl["kmax"]), device=local)

    def barrier():
        if dist is not None:
            import torch
            dist.barrier()
            torch.cuda.synchronize()

    for _ in range(warmup):
        eng.fit(off, x, l, r, pa, sid, seeds)
    sampler = ClockSampler(local)
    if rank == 0:
        sampler.start()
    barrier()
    t0 = time.perf_counter()
    acc = {}
    for _ in range(steps):
        out = eng.fit(off, x, l, r, pa, sid, seeds)
        for k, v in out.timing.items():
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
        t = torch.tensor([wall, dev_ms], dtype=torch.float64, device="cuda")
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        wall, dev_ms = t.tolist()
        s = torch.tensor([work, float(n_utr), acc["launches"], acc["h2d_bytes"], acc["d2h_bytes"], acc["waves"]],
                         dtype=torch.float64, device="cuda")
        dist.all_reduce(s, op=dist.ReduceOp.SUM)
        work_all, utr_all, launches, h2d, d2h, waves = s.tolist()
    else:
        work_all, utr_all, launches, h2d, d2h, waves = work, float(n_utr), acc["launches"], acc["h2d_bytes"], acc["d2h_bytes"], acc["waves"]
    if rank != 0:
        eng.close()
        return None

    K = steps
    hbm_peak, hbm_src = measured_peaks()
    fp64 = eng.fp64_peaks()                     # measured on this GPU, this run (no FP64 entry in MEASURED_PEAKS.json)
    sfu = eng.sfu_peaks()                       # FP32 FMA / MUFU / FP64 exp, log streams
    cluster = alone.get("resident_ms", 0.0) > alone["scan_ms"]     # which EM kernel dominates this workload
    if cluster:
        dom_s = alone["resident_ms"] / 1e3
        dom_flops = alone["resident_grid_flops"]
        dom_launches = max(int(alone["resident_launches"]), 1)
        kernel = ("em_cluster_kernel (cluster-resident EM: every E pass and every max_alpha_beta grid arg-max -- FP64 MMA "
                  "tiles -- of a UTR's chains in one launch; only the grid search's algorithmic flops are counted, the E passes' "
                  "time is inside the denominator)")
    else:
        dom_s = alone["scan_ms"] / 1e3
        dom_flops = alone["em_grid_flops"]
        dom_launches = max(int(alone["scan_launches"]), 1)
        kernel = "em_scan_kernel (max_alpha_beta grid arg-max as a blocked FP64 MMA product)"
    ach_tflops = dom_flops / dom_s / 1e12
    traffic, traffic_src = scan_traffic(cluster)
    res = {
        "metric": "infer_pa_utrs_per_s", "value": utr_all * K / (dev_ms / 1e3), "unit": "UTR/s",
        "n_gpus": world, "steps": K, "warmup": warmup, "ms_per_step": dev_ms / K,
        "higher_is_better": True, "scaling": wl["scaling"], "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["label"],
                   "seed_policy": "file", "tensor_storage": "f32 (FP64 arithmetic)",
                   "l2": "inputs larger than L2: a wave's marginal tensors (~0.26 GB for 100 cfg-2 UTRs) exceed the 126 MB L2 and every "
                         "step streams all waves again; nothing is flushed between steps",
                   "value_time": "GPU busy time = union of the CUDA-event kernel intervals (EM stream and the "
                                 "likelihood stream that works one wave ahead), max over ranks",
                   "e2e_time": "wall clock of Engine.fit (C ABI scape_b200_fit_batch) on host buffers, max over ranks"},
        "read_comp_em_iter_per_s": work_all * K / (dev_ms / 1e3),
        "read_comp_em_iter_per_s_e2e": work_all * K / wall,
        "em_iterations_per_step": iters,
        "e2e": {"value": utr_all * K / wall, "unit": "UTR/s", "h2d_bytes_per_step": h2d / K,
                "d2h_bytes_per_step": d2h / K, "ms_per_step": 1e3 * wall / K},
        "gpu_launches": int(launches),
        "roofline": {"kernel": kernel,
                     "bound": "tensor", "achieved": ach_tflops, "peak": fp64["dmma_tflops"], "unit": "TFLOP/s",
                     "frac": ach_tflops / fp64["dmma_tflops"], "traffic": traffic, "traffic_source": traffic_src,
                     "timing": "CUDA events around every launch of that kernel in one extra pass with wave pipelining off "
                               "(kernel alone on the GPU), rank 0",
                     "peak_source": "FP64 mma.m8n8k4 stream measured in this run (scape_b200_fp64_peaks); "
                                    f"CUDA-core DFMA stream {fp64['dfma_tflops']:.1f} TFLOP/s",
                     "algorithmic_flops_per_launch": dom_flops / dom_launches,
                     "avg_launch_ms": 1e3 * dom_s / dom_launches,
                     "launches": dom_launches,
                     "share_of_gpu_time": 1e3 * dom_s / alone["device_busy_ms"],
                     "hbm_view": {"algorithmic_GBps": alone["em_grid_bytes"] / max(alone["em_ms"], 1e-9) / 1e6, "peak_GBps": hbm_peak,
                                  "peak_source": hbm_src, "loaded_GBps": alone["em_scan_bytes"] / max(alone["em_ms"], 1e-9) / 1e6,
                                  "note": "SURVEY 8d algorithmic bytes = 8*W_k*B*N per chain iteration over the EM time; the kernels "
                                          "load a tensor tile once for all chains that need it, from L2 where the UTR stays resident"}},
        "phases_ms_per_step": {k: acc[k] / K for k in ("table_ms", "tensor_ms", "em_ms", "resident_ms", "estep_ms", "scan_ms", "label_ms",
                                                        "host_prep_ms", "host_rng_ms", "device_busy_ms", "total_ms")},
        "phases_alone_ms": {k: alone[k] for k in ("table_ms", "tensor_ms", "em_ms", "resident_ms", "estep_ms", "scan_ms", "label_ms",
                                                  "device_busy_ms", "total_ms")},
        "tensor_exp_per_s": alone["tensor_exp"] / (alone["tensor_ms"] / 1e3),
        "likelihood_rooflines": {
            "peaks": sfu,
            "table_kernel": {"exp_per_s": alone["table_exp"] / (alone["table_ms"] / 1e3), "peak_exp_per_s": sfu["exp_f64_gops"] * 1e9,
                             "frac": alone["table_exp"] / (alone["table_ms"] / 1e3) / (sfu["exp_f64_gops"] * 1e9),
                             "note": "N*T*13 FP64 exp per UTR (SURVEY 8d) over the kernel's time vs the measured FP64 exp() stream; "
                                     "entries of reads that cannot reach theta skip the exps (the count is the reference's)"},
            "tensor_kernels": {"reference_exp_per_s": alone["tensor_exp"] / (alone["tensor_ms"] / 1e3),
                               "peak_exp_per_s": sfu["exp_f64_gops"] * 1e9,
                               "x_of_exp_peak": alone["tensor_exp"] / (alone["tensor_ms"] / 1e3) / (sfu["exp_f64_gops"] * 1e9),
                               "note": "the reference's 307 exp per (fragment, alpha) over the kernels' time; above 1 because the "
                                       "kernels take ~6 exp + 559 FMA per (fragment, alpha) instead and skip the all-sentinel entries"}},
        "waves_per_step": waves / K,
        "clocks": clocks,
    }
    if cpu_arm is not None and wl["name"] == "cfg2":
        # cpu_baseline leg (rank 0, N = 1): bounded sample = the first m UTRs of `cores` chunk files under
        # the file seed policy -- the same UTRs, same streams the GPU arm just fitted, so the results are
        # compared UTR by UTR (the oracle as the checker).
        m = args.ref_utrs_per_file
        a = argparse.Namespace(**vars(args)); a.utrs = n_utr; a.per_file = wl["per_file"]; a.kmax = wl["kmax"]
        wall_cpu, rows = cpu_arm.step(0, m, a)
        v = len(rows) / wall_cpu
        same_k = ok = 0
        for idx, Kc, alpha, beta, ws, lb, nf in rows:
            if int(out.K[idx]) != Kc:
                continue
            same_k += 1
            lbg = out.lb_arr[idx, out.n_lb[idx] - 1]
            ok += bool(np.max(np.abs(out.alpha[idx, :Kc] - alpha), initial=0) <= 1 and
                       np.max(np.abs(out.beta[idx, :Kc] - beta), initial=0) <= 1e-3 and
                       np.max(np.abs(out.ws[idx, :Kc + 1] - ws)) <= 1e-3 and abs(lbg - lb) <= 1e-6 * abs(lb))
        res["cpu_baseline"] = {"value": v, "unit": "UTR/s", "cores": cpu_arm.cores, "kind": "port",
                               "sample": f"first {m} UTRs of {cpu_arm.cores} chunk files of the workload (np.random.seed(1) per "
                                         f"file, UTRs serial on the stream), oracle port, {wall_cpu:.1f} s wall"}
        res["parity_check"] = {"utrs": len(rows), "K_identical": same_k, "within_north_star_tolerance": ok,
                               "against": "cpu_baseline leg (oracle port), same UTRs and RNG streams as the GPU arm"}
    eng.close()
    if files and wl["name"] == "cfg2":
        res["e2e_files"] = files_leg(utrs, wl["per_file"], local)
    return res


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
    ap.add_argument("--no-files", action="store_true", help="skip the file-to-file leg (e2e_files)")
    ap.add_argument("--no-cfg3", action="store_true", help="skip the cfg-3 block reported beside the cfg-2 line")
    ap.add_argument("--ref-utrs-per-file", type=int, default=3,
                    help="CPU arm / cpu_baseline leg: leading UTRs of each chunk file fitted per step (bounded sample)")
    ap.add_argument("--workload", default="cfg2", choices=["cfg2", "cfg3", "cfg4", "cfg5"],
                    help="BASELINE.json configs[1..4]; cfg2 is the benchmark line (with a cfg3 block beside it)")
    ap.add_argument("--kmax", type=int, default=5, help="n_max_apa (cfg5 sweeps it)")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference(args)

    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    # the CPU pool is forked before this process touches CUDA
    cpu_arm = CpuArm(os.cpu_count() or 1) if (rank == 0 and world == 1 and not args.no_cpu and args.workload == "cfg2") else None
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))

    res = run_workload(make_workload(args.workload, args, rank, world), args, args.steps, args.warmup, rank, local, world,
                       dist, cpu_arm, files=(world == 1 and not args.no_files))
    if cpu_arm is not None:
        cpu_arm.close()
    if args.workload == "cfg2" and not args.no_cfg3 and args.utrs == N_UTR:
        # BASELINE.json configs[2] beside the headline: the 20k-UTR heavy-tailed set, strong scaling
        # (chunk files LPT-packed over the ranks), same timing rules
        blk = run_workload(make_workload("cfg3", args, rank, world), args, max(args.steps, 3), max(args.warmup, 3), rank,
                           local, world, dist)
        if rank == 0:
            res["cfg3"] = {k: blk[k] for k in ("value", "unit", "ms_per_step", "scaling", "steps", "warmup", "e2e", "gpu_launches",
                                               "waves_per_step", "read_comp_em_iter_per_s", "phases_ms_per_step",
                                               "phases_alone_ms", "clocks")}
            res["cfg3"]["workload"] = blk["config"]["workload"]
            res["cfg3"]["roofline"] = {k: blk["roofline"][k] for k in ("kernel", "bound", "achieved", "peak", "unit", "frac",
                                                                       "share_of_gpu_time")}
    if rank == 0:
        print(json.dumps(res))
    if dist is not None:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
