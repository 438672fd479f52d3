#!/usr/bin/env python
"""End to end through the file interface ("incl. unpickle / bin / init / pickle"):
chunk pickles on disk -> `scape_b200.apa_core.infer_files` -> result pickles on disk.

bench.py's `e2e` starts from read columns in host memory; this script adds what the `scape infer_pa`
user also pays: unpickling the `prepare_input` DataFrames, packing them into CSR columns, building the
`Parameters` objects and pickling them.  One JSON line: UTR/s over the whole call, and the split.

  python scripts/bench_cli_e2e.py [--utrs 10000] [--reads 500] [--per-file 100] [--devices 0,1,...]
"""
import argparse
import json
import os
import sys
import tempfile
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from scape_b200 import apa_core, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--utrs", type=int, default=10000)
    ap.add_argument("--reads", type=int, default=500)
    ap.add_argument("--per-file", type=int, default=100)
    ap.add_argument("--devices", default="0")
    ap.add_argument("--repeat", type=int, default=2, help="timed repetitions after one warm-up call")
    args = ap.parse_args()
    devices = [int(d) for d in args.devices.split(",")]
    with tempfile.TemporaryDirectory() as tmp:
        utrs = synth.make_batch(args.utrs, args.reads)
        t0 = time.perf_counter()
        paths = synth.write_chunk_files(utrs, tmp, per_file=args.per_file)
        t_write = time.perf_counter() - t0
        in_bytes = sum(os.path.getsize(p) for p in paths)
        apa_core.infer_files(paths, tmp, devices=devices)                 # warm-up (allocations, page cache)
        walls = []
        for _ in range(args.repeat):
            t0 = time.perf_counter()
            outs = apa_core.infer_files(paths, tmp, devices=devices)
            walls.append(time.perf_counter() - t0)
        t0 = time.perf_counter()
        n_obj = sum(len(apa_core.read_chunk_file(p)) for p in paths)
        t_unpickle = time.perf_counter() - t0
        out_bytes = sum(os.path.getsize(o) for o in outs)
    wall = min(walls)
    print(json.dumps({
        "metric": "infer_pa_utrs_per_s_file_to_file", "value": args.utrs / wall, "unit": "UTR/s",
        "n_gpus": len(devices), "utrs": args.utrs, "reads_per_utr": args.reads, "chunk_files": len(paths),
        "wall_s": wall, "unpickle_inputs_s": t_unpickle, "input_bytes": in_bytes, "output_bytes": out_bytes,
        "objects": n_obj, "generate_inputs_s": t_write,
        "note": "chunk pickles -> infer_files (unpickle, CSR packing, fit_batch, Parameters, pickle) -> result pickles",
    }))


if __name__ == "__main__":
    main()
