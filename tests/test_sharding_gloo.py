"""Multi-GPU partitioning on CPU: deterministic LPT packing, exact cover, and
a world_size-2 gloo run in which every rank derives the same plan and the results gathered on rank
0 come back in input order."""
import os
import socket
import sys

import numpy as np
import pytest

from scape_b200 import shard


def test_lpt_partition_is_an_exact_balanced_cover():
    rng = np.random.default_rng(0)
    reads = [np.clip(np.rint(np.exp(rng.normal(np.log(600), 1.6, size=100))), 10, 200000).astype(int) for _ in range(200)]
    costs = shard.stream_costs(reads)
    for g in (1, 2, 4, 8):
        parts = shard.lpt_partition(costs, g)
        flat = sorted(i for p in parts for i in p)
        assert flat == list(range(200))
        assert all(p == sorted(p) for p in parts)
        assert shard.imbalance(costs, parts) < 1.05
    assert shard.lpt_partition(costs, 8) == shard.lpt_partition(costs, 8)


def test_cost_model_is_monotone():
    c = [shard.utr_cost(n) for n in (10, 100, 1000, 10000, 100000)]
    assert all(a < b for a, b in zip(c, c[1:]))
    assert shard.utr_cost(500, 20000) > shard.utr_cost(500, 2000)


def _worker(rank, world, port, q):
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    reads = [[100 + 37 * ((s * 7 + i) % 11) for i in range(5)] for s in range(9)]
    parts = shard.lpt_partition(shard.stream_costs(reads), world)
    mine = [(s, [f"stream{s}-utr{i}" for i in range(5)]) for s in parts[rank]]      # stand-in for the GPU fit
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    if rank == 0:
        merged = dict(kv for part in gathered for kv in part)
        q.put([merged[s] for s in sorted(merged)])
    dist.barrier()
    dist.destroy_process_group()


def test_world_size_2_gloo_gather_keeps_input_order():
    import torch.multiprocessing as mp
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        port = s.getsockname()[1]
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    out = q.get(timeout=120)
    for p in procs:
        p.join(timeout=120)
        assert p.exitcode == 0
    assert out == [[f"stream{s}-utr{i}" for i in range(5)] for s in range(9)]


def test_infer_files_multi_gpu_plan_and_output_order(tmp_path, monkeypatch):
    """`infer_files(devices=[...])`: files are dealt to the GPUs by cost, every file is fitted exactly
    once as its own stream (seed 1), and the result pickles land under the reference's names in input
    order.  The GPU fit is replaced by a stand-in here (CPU test); the real thing is the same
    `fit_chunks` call the single-GPU path makes."""
    import pickle
    import threading
    from scape_b200 import apa_core, synth

    sizes = [[60, 40], [4000, 30, 30], [200], [900, 800, 700], [25]]
    utrs, k = [], 0
    for f, ss in enumerate(sizes):
        for n in ss:
            utrs.append(synth.make_utr(600 + k, n))
            k += 1
    paths, pos = [], 0
    for f, ss in enumerate(sizes):
        paths += synth.write_chunk_files(utrs[pos:pos + len(ss)], str(tmp_path), per_file=len(ss), stem=f"f{f}")
        pos += len(ss)

    seen, lock = [], threading.Lock()

    def fake_fit_chunks(chunks, seeds=None, device=0, **kw):
        with lock:
            seen.append((device, [c[0][0] for c in chunks], list(seeds)))
        return [[("fit", gi, len(df), device) for gi, df in c] for c in chunks]

    monkeypatch.setattr(apa_core, "fit_chunks", fake_fit_chunks)
    outs = apa_core.infer_files(paths, str(tmp_path), devices=[0, 1, 2])
    assert [os.path.basename(o) for o in outs] == [os.path.basename(p)[:-10] + ".res.pkl" for p in paths]
    firsts = sorted(g for _, gs, _ in seen for g in gs)
    assert firsts == sorted(apa_core.read_chunk_file(p)[0][0] for p in paths)          # every file once
    assert all(set(sd) == {1} for _, _, sd in seen) and len({d for d, _, _ in seen}) == 3
    pos = 0
    for o, ss in zip(outs, sizes):
        with open(o, "rb") as fh:
            recs = []
            while True:
                try:
                    recs.append(pickle.load(fh))
                except EOFError:
                    break
        assert [r[1] for r in recs] == [u.gene_info_str for u in utrs[pos:pos + len(ss)]]
        assert [r[2] for r in recs] == ss
        pos += len(ss)
    # the heaviest file sits alone on its GPU
    plan = apa_core.plan_multi_gpu([apa_core.read_chunk_file(p) for p in paths], 3)
    assert sorted(i for p in plan for i in p) == list(range(len(paths)))
    assert [1] in plan


def test_bench_cfg3_plan_covers_every_utr_once():
    """bench.py --workload cfg3 (strong scaling): whatever the number of ranks, every chunk file of the
    20k-UTR heavy-tailed set is fitted by exactly one rank and the estimated loads are balanced."""
    sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
    import bench
    for world in (1, 2, 4, 8):
        counts, cut, files, parts = bench.cfg3_plan(20000, 100, world)
        assert len(files) == 200 and len(parts) == world
        assert sorted(f for p in parts for f in p) == list(range(200))
        utrs = sorted(i for p in parts for f in p for i in files[f])
        assert utrs == list(range(20000))
        loads = [sum(shard.utr_cost(int(counts[i]), 20000 if counts[i] >= cut else 2000) for f in p for i in files[f]) for p in parts]
        assert max(loads) / (sum(loads) / world) < 1.10


class _FakeEngine:
    """Stand-in for the GPU fit: a UTR's results are a deterministic function of its own reads only (like
    the real fit with one seeded stream per file, they must not depend on what else is in the call)."""

    def __init__(self, params, device=0, tensor_dtype=None):
        self.params = params

    def close(self):
        pass

    def set_host_threads(self, n):
        self.host_threads = n

    def fit(self, off, x, l, r, pa, sid, seeds, stream_state=None):
        from scape_b200 import _lib
        n_utr = len(off) - 1
        out = _lib.FitOutput(n_utr, int(off[-1]))
        for u in range(n_utr):
            n = int(off[u + 1] - off[u])
            k = 1 + n % 3
            out.K[u], out.L[u], out.bic[u], out.n_lb[u] = k, 2000 + n % 11, -float(n) - 0.5 * float(x[off[u]] % 3), 3
            out.alpha[u, :k] = 100.0 * (np.arange(k) + 1) + x[off[u]] % 7
            out.beta[u, :k] = 10.0 + 5 * np.arange(k)
            out.ws[u, :k + 1] = 1.0 / (k + 1)
            out.lb_arr[u, :3] = [-3.0 * n, -2.5 * n, -2.4 * n]
            out.label[off[u]:off[u + 1]] = (np.arange(n) + int(l[off[u]])) % (k + 1)
        return out


def test_infer_files_worker_processes_write_the_same_pickles(tmp_path, monkeypatch):
    """Many-file calls unpickle / pack the inputs and build / pickle the outputs in worker processes;
    the result files must hold exactly what the in-process path writes (GPU fit replaced by a stand-in)."""
    import pickle
    from scape_b200 import _lib, apa_core, synth
    monkeypatch.setattr(_lib, "Engine", _FakeEngine)
    monkeypatch.setattr(apa_core, "_engines", {})         # (infer_files keeps one engine per device and parameter set)
    utrs = [synth.make_utr(700 + i, 30 + 17 * (i % 9)) for i in range(36)]
    a_dir, b_dir = tmp_path / "inproc", tmp_path / "pooled"
    a_dir.mkdir(); b_dir.mkdir()
    a_paths = synth.write_chunk_files(utrs, str(a_dir), per_file=4)
    b_paths = synth.write_chunk_files(utrs, str(b_dir), per_file=4)
    a_out = apa_core.infer_files(a_paths, str(a_dir), io_workers=0)
    b_out = apa_core.infer_files(b_paths, str(b_dir), io_workers=3)
    assert [os.path.basename(p) for p in a_out] == [os.path.basename(p) for p in b_out] and len(a_out) == 9
    # ... and with the files dealt over three (stand-in) GPUs on top of the worker processes
    c_dir = tmp_path / "pooled3"
    c_dir.mkdir()
    c_paths = synth.write_chunk_files(utrs, str(c_dir), per_file=4)
    c_out = apa_core.infer_files(c_paths, str(c_dir), io_workers=3, devices=[0, 1, 2])

    def load(path):
        recs = []
        with open(path, "rb") as fh:
            while True:
                try:
                    recs.append(pickle.load(fh))
                except EOFError:
                    return recs

    n = 0
    for pa_, pb_ in list(zip(a_out, b_out)) + list(zip(a_out, c_out)):
        ra, rb = load(pa_), load(pb_)
        assert len(ra) == len(rb) == 4
        for x, y in zip(ra, rb):
            assert type(x) is type(y) and type(y).__module__ == "scape.apa_core"
            assert sorted(vars(x)) == sorted(vars(y))
            for key, vx in vars(x).items():
                vy = getattr(y, key)
                if isinstance(vx, np.ndarray):
                    assert vx.dtype == vy.dtype and np.array_equal(vx, vy), key
                elif isinstance(vx, list):
                    assert vx == vy and all(type(p) is type(q) for p, q in zip(vx, vy)), key
                else:
                    assert vx == vy and type(vx) is type(vy), key
            n += 1
    assert n == 72
