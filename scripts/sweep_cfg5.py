#!/usr/bin/env python
"""BASELINE.json configs[4]: stress sweep n_max_apa 3..10 x 1k..1M reads per UTR.

One row per (n_max_apa, reads per UTR): UTR/s (device busy time and end to end), read*comp*EM-iter/s,
E-step / scan time, and the scan's algorithmic FP64 rate against the DMMA peak measured in the same
process.  Every UTR is its own RNG stream, so a batch is one wave.  Batch sizes shrink with the read
count (256 UTRs up to 10k reads, 32 at 100k, 4 at 1M) to bound the host-side generation time; rows
say how many UTRs they used.  Output: JSON lines on stdout.
"""
import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scape_b200 import _lib, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--kmax", type=int, nargs="*", default=[3, 5, 8, 10])
    ap.add_argument("--reads", type=int, nargs="*", default=[1000, 10000, 100000, 1000000])
    ap.add_argument("--budget-s", type=float, default=240.0, help="stop starting new points after this many seconds")
    args = ap.parse_args()
    t_start = time.time()
    peaks = None
    for reads in args.reads:
        n_utr = 256 if reads <= 10000 else 32 if reads <= 100000 else 4
        utrs = [synth.make_utr(900000 + i, reads, long_utr=(reads >= 100000)) for i in range(n_utr)]
        off = np.zeros(n_utr + 1, np.int64)
        np.cumsum([u.n_reads for u in utrs], out=off[1:])
        cat = lambda k: np.concatenate([np.asarray(getattr(u, k), dtype=np.float64) for u in utrs])
        cols = (cat("x"), cat("l"), cat("r"), cat("pa"))
        sid = np.arange(n_utr, dtype=np.int32)
        seeds = np.ones(n_utr, np.uint32)
        for kmax in args.kmax:
            if time.time() - t_start > args.budget_s:
                return
            with _lib.Engine(_lib.make_params(n_max_apa=kmax)) as eng:
                if peaks is None:
                    peaks = eng.fp64_peaks()
                eng.set_overlap(False)
                eng.fit(off, *cols, sid, seeds)                  # warm-up (allocations)
                t0 = time.perf_counter()
                out = eng.fit(off, *cols, sid, seeds)
                wall = time.perf_counter() - t0
                tm = out.timing
                work = float(out.em_work[:, 0].sum())
                scan_s = max(tm["scan_ms"], 1e-9) / 1e3
                print(json.dumps({
                    "n_max_apa": kmax, "reads_per_utr": reads, "n_utr": n_utr,
                    "n_frag_mean": float(out.n_frag.mean()), "n_theta_mean": float(out.n_theta.mean()),
                    "utr_per_s_device": n_utr / (tm["device_busy_ms"] / 1e3), "utr_per_s_e2e": n_utr / wall,
                    "read_comp_em_iter_per_s": work / (tm["device_busy_ms"] / 1e3),
                    "raw_read_comp_em_iter_per_s": work * (reads / max(float(out.n_frag.mean()), 1.0)) / (tm["device_busy_ms"] / 1e3),
                    "estep_ms": tm["estep_ms"], "scan_ms": tm["scan_ms"], "tensor_ms": tm["tensor_ms"], "table_ms": tm["table_ms"],
                    "device_busy_ms": tm["device_busy_ms"], "wall_ms": 1e3 * wall,
                    "scan_tflops": tm["em_grid_flops"] / scan_s / 1e12, "dmma_peak_tflops": peaks["dmma_tflops"],
                    "scan_frac_of_dmma_peak": tm["em_grid_flops"] / scan_s / 1e12 / peaks["dmma_tflops"],
                    "K_selected_hist": np.bincount(out.K, minlength=kmax + 3).tolist(),
                }), flush=True)


if __name__ == "__main__":
    main()
