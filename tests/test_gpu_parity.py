"""Parity of the CUDA path (through the C ABI) with the CPU oracle and the reference goldens.
Run on the B200 box: python -m pytest tests -m gpu."""
import numpy as np
import pytest

from oracle import scape_oracle as so
from scape_b200 import _lib, synth
from scape_b200.apa_core import fit_chunks
from _helpers import check_against_golden, golden_chunk

pytestmark = pytest.mark.gpu


# Tensor storage: "f64" is the strict mode (results equal the oracle to rounding), "f32" is the
# default (tensor values rounded once to float, all arithmetic FP64).
TOL = {"f64": dict(tensor=1e-12, lb=1e-10, ws=1e-12), "f32": dict(tensor=1.2e-7, lb=1e-7, ws=1e-6)}


@pytest.fixture(scope="module", params=["f64", "f32"])
def engine(request):
    assert _lib.load().scape_b200_device_count() > 0, "no CUDA device: the gpu tests cannot fall back to anything"
    with _lib.Engine(_lib.make_params(), tensor_dtype=request.param) as e:
        e.dtype = request.param
        yield e


def _model(u, with_tensor=True):
    prm = dict(so.DEFAULTS)
    prm["utr_length"] = so.resolve_utr_length(u.x, u.l, prm)
    m = so.build_model(u.x, u.l, u.r, u.pa, prm)
    m.unif_loglik = so.uniform_loglik(m)
    m.table = so.theta_table(m, m.theta)
    if with_tensor:
        m.tensor = so.get_loglik_marginal_tensor(m.theta, m.betas, m.table)
    return m


def _rel_err_on_finite(got, want):
    fin = want > -1e30
    assert np.array_equal(fin, got > -1e30), "sentinel pattern differs"
    assert np.all(got[~fin] == so.SENTINEL)
    if not fin.any():
        return 0.0
    return float(np.max(np.abs(got[fin] - want[fin]) / np.maximum(np.abs(want[fin]), 1e-300)))


# ---- gate (i): element-wise table / tensor ---------------------------------------------------------
@pytest.mark.parametrize("ui,reads", [(0, 300), (7, 60), (22, 1500)])
def test_theta_table_and_marginal_tensor(engine, ui, reads):
    m = _model(synth.make_utr(ui, reads))
    tab = engine.loglik_table(m.x, m.l, m.r, m.pa, m.theta)
    assert _rel_err_on_finite(tab, m.table) < 1e-12
    ten = engine.marginal_tensor(m.theta, m.betas, m.table)
    assert _rel_err_on_finite(ten, m.tensor) < TOL[engine.dtype]["tensor"]


def test_theta_table_with_known_polya_lengths(engine):
    """r_known reads (taichi_core.py:111-132) never occur in 10x data; cover the kernel branch anyway,
    including r above every s (mass = 0 -> +inf like the reference's log(0))."""
    rng = np.random.default_rng(11)
    n = 120
    x = rng.integers(0, 1500, n).astype(float)
    l = rng.integers(20, 133, n).astype(float)
    r = np.where(rng.random(n) < 0.6, rng.integers(5, 139, n), np.nan).astype(float)
    pa = np.where(np.isnan(r) & (rng.random(n) < 0.3), x + l - 1, np.nan)
    theta = np.arange(20.0, 2000.0, 9.0)
    s_dis = np.arange(20, 150, 10)
    pmf = np.repeat(1 / 13, 13); pmf = pmf / sum(pmf)
    want = np.zeros((n, len(theta)))
    kn, un, pj = ~np.isnan(r) & np.isnan(pa), np.isnan(r) & np.isnan(pa), ~np.isnan(pa)
    for t, th in enumerate(theta):
        want[kn, t] = so.loglik_xlr_t_r_known(x[kn], l[kn], r[kn], s_dis, pmf, th, 300, 50)
        want[un, t] = so.loglik_xlr_t_r_unknown(x[un], l[un], r[un], s_dis, pmf, th, 300, 50)
        want[pj, t] = so.loglik_xlr_t_pa(x[pj], l[pj], pa[pj], th, 50)
    got = engine.loglik_table(x, l, r, pa, theta)
    assert _rel_err_on_finite(got, want) < 1e-12


def test_marginal_fast_and_generic_kernels_agree(engine):
    """Interior alpha rows of the default grid go through tensor_interior_kernel (constant weights,
    one exp per table entry and tile); SCAPE_B200_TENSOR_FAST=0 engines use the generic kernel for
    every row.  Both must match the oracle, including a long UTR with many tiles."""
    import os
    m = _model(synth.make_utr(22, 2500, long_utr=True))
    fast = engine.marginal_tensor(m.theta, m.betas, m.table)
    assert _rel_err_on_finite(fast, m.tensor) < TOL[engine.dtype]["tensor"]
    os.environ["SCAPE_B200_TENSOR_FAST"] = "0"
    try:
        with _lib.Engine(_lib.make_params(), tensor_dtype=engine.dtype) as slow_engine:
            slow = slow_engine.marginal_tensor(m.theta, m.betas, m.table)
    finally:
        del os.environ["SCAPE_B200_TENSOR_FAST"]
    assert _rel_err_on_finite(slow, m.tensor) < TOL[engine.dtype]["tensor"]
    assert np.array_equal(fast > -1e30, slow > -1e30)
    # the rows whose windows are clipped by the grid ends: constant-weight kernel (default) vs generic kernel
    os.environ["SCAPE_B200_TENSOR_EDGES"] = "0"
    try:
        with _lib.Engine(_lib.make_params(), tensor_dtype=engine.dtype) as edge_engine:
            mixed = edge_engine.marginal_tensor(m.theta, m.betas, m.table)
    finally:
        del os.environ["SCAPE_B200_TENSOR_EDGES"]
    assert _rel_err_on_finite(mixed, m.tensor) < TOL[engine.dtype]["tensor"]
    assert np.array_equal(fast > -1e30, mixed > -1e30)


@pytest.mark.parametrize("n_theta", [1, 5, 22, 42, 43, 50])
def test_marginal_tensor_on_grids_shorter_than_the_widest_window(engine, n_theta):
    """Every alpha row of a short UTR has its windows clipped on one side or both (a grid of fewer
    than 43 points never holds a whole beta = 70 window)."""
    m = _model(synth.make_utr(3, 300), with_tensor=False)
    theta = m.theta[:n_theta]
    table = np.ascontiguousarray(m.table[:, :n_theta])
    want = so.get_loglik_marginal_tensor(theta, m.betas, table)
    got = engine.marginal_tensor(theta, m.betas, table)
    assert _rel_err_on_finite(got, want) < TOL[engine.dtype]["tensor"]


def test_marginal_tensor_on_irregular_fixed_mode_grid(engine):
    m = _model(synth.make_utr(3, 300), with_tensor=False)
    keep = np.r_[10:60, 100:131, 180:200]
    theta = m.theta[keep]
    betas = np.array([45.0, 50.0, 55.0, 60.0])
    table = np.ascontiguousarray(m.table[:, keep])
    want = so.get_loglik_marginal_tensor(theta, betas, table)
    got = engine.marginal_tensor(theta, betas, table)
    assert _rel_err_on_finite(got, want) < TOL[engine.dtype]["tensor"]


# ---- gate (ii): per-chain traces from identical init blobs -------------------------------------------
@pytest.mark.parametrize("ui,reads", [(0, 300), (31, 300), (22, 3000)])
def test_em_chain_traces(engine, ui, reads):
    u = synth.make_utr(ui, reads)
    m = _model(u)
    m.prof_x, m.prof_y = so.coverage_profile(m)
    m.peak_idx, m.peak_w = so.find_profile_peaks(m)
    rng = np.random.RandomState(3)
    inits, want = [], []
    for K in (7, 5, 4, 3, 2, 1):
        for _ in range(2):
            ch = so.draw_chain(m, K, rng)
            init = dict(K=K, a_idx=ch.a_idx.copy(), b_idx=ch.b_idx.copy(), ws=ch.ws.copy())
            m.trace = []
            done = so.run_chain(m, ch, rng)
            init["k_order"] = done.k_order
            inits.append(init)
            want.append((done, m.trace))
    # one weights-only chain (fixed_inference, apa_core.py:708-711)
    base = want[2][0]
    slim = so.Chain(a_idx=base.a_idx[:3].copy(), b_idx=base.b_idx[:3].copy(), ws=so.draw_weights(m, 3, rng))
    init = dict(K=3, a_idx=slim.a_idx.copy(), b_idx=slim.b_idx.copy(), ws=slim.ws.copy(), weights_only=1)
    m.trace = []
    done = so.run_chain(m, slim, rng, weights_only=True)
    init["k_order"] = done.k_order
    inits.append(init)
    want.append((done, m.trace))

    got, (ta, tb, tw) = engine.em_chains(m.tensor, m.cnt, m.unif_loglik, inits, trace=True)
    for i, (io, (ref, tr)) in enumerate(zip(got, want)):
        K = io.K
        assert io.n_iter == len(ref.lb_arr), f"chain {i}: iteration count"
        for j, step in enumerate(tr):
            assert list(ta[i, j, :K]) == list(step["a_idx"]) and list(tb[i, j, :K]) == list(step["b_idx"])
            assert np.allclose(tw[i, j, :K + 1], step["ws"], rtol=0, atol=TOL[engine.dtype]["ws"])
            assert abs(io.lb_arr[j] - step["lb"]) <= TOL[engine.dtype]["lb"] * abs(step["lb"])
        assert abs(io.bic - ref.bic) <= TOL[engine.dtype]["lb"] * abs(ref.bic)
        assert list(io.a_idx[:K]) == list(ref.a_idx) and list(io.b_idx[:K]) == list(ref.b_idx)


# ---- gate (iii): whole-UTR results --------------------------------------------------------------------
@pytest.mark.parametrize("dtype", ["f64", "f32"])
@pytest.mark.parametrize("case", ["toy", "chr17", "chr19", "synth8", "synth_rerun"])
def test_fit_matches_reference_goldens(golden, case, dtype):
    spec = golden["cases"][case]
    res = fit_chunks([golden_chunk(golden, case)], seeds=[1], tensor_dtype=dtype, **spec["params"])[0]
    assert len(res) == len(spec["utrs"])
    for i, r in enumerate(res):
        check_against_golden(r, spec["utrs"][i], golden["labels"][f"{case}/{i}"], tight=True,
                             lb_rtol=1e-9 if dtype == "f64" else 1e-7, ws_atol=1e-9 if dtype == "f64" else 1e-6)


def test_fixed_mode_matches_reference_goldens(golden, tmp_path):
    import pickle
    from scape.apa_core import Parameters
    spec = golden["cases"]["synth_fixed"]
    pp = spec["pre_para"]
    pre = Parameters(alpha_arr=np.array(pp["alpha_arr"]), beta_arr=np.array(pp["beta_arr"]), ws=None, L=pp["L"])
    f = tmp_path / "pre.pkl"
    with open(f, "wb") as fh:
        pickle.dump(pre, fh)
    res = fit_chunks([golden_chunk(golden, "synth_fixed")], seeds=[1], fixed_run_mode=True,
                     pre_para_pkl_file=str(f))[0]
    for i, r in enumerate(res):
        # the third UTR collapses both sites onto one grid point: a degenerate fit whose weight split
        # is decided by rounding noise, so only the BASELINE.json tolerances are required there
        check_against_golden(r, spec["utrs"][i], golden["labels"][f"synth_fixed/{i}"], tight=(i < 2),
                             lb_rtol=1e-7, ws_atol=1e-6)


def test_many_streams_in_one_call_equal_one_call_per_file():
    """Wave scheduling must not change results: 3 chunks fitted together == fitted one by one, and
    equal to the oracle's per-file streams."""
    chunks = [[synth.make_utr(100 + 4 * f + i, 200) for i in range(4)] for f in range(3)]
    frames = [[(u.gene_info_str, synth.to_dataframe(u)) for u in c] for c in chunks]
    together = fit_chunks(frames, seeds=[1, 1, 1])
    for f, c in enumerate(chunks):
        alone = fit_chunks([frames[f]], seeds=[1])[0]
        rng = np.random.RandomState(1)
        for u, a, b in zip(c, together[f], alone):
            assert a.K == b.K and np.array_equal(a.alpha_arr, b.alpha_arr) and np.array_equal(a.ws, b.ws)
            assert a.lb_arr == b.lb_arr and np.array_equal(a.label_arr, b.label_arr)
            w = so.fit_utr(u.x, u.l, u.r, u.pa, rng)
            assert a.K == w.K and np.array_equal(a.alpha_arr, w.alpha_arr)
            assert np.allclose(a.beta_arr, w.beta_arr) and np.allclose(a.ws, w.ws, atol=1e-6)
            assert abs(a.lb_arr[-1] - w.lb_arr[-1]) <= 1e-7 * abs(w.lb_arr[-1])
            assert np.array_equal(a.label_arr, w.label_arr)


def test_ragged_and_tiny_utrs():
    """10-read UTRs up to a 30k-read one in the same wave (cfg-3's heavy tail), vs the oracle."""
    sizes = [10, 13, 4000, 25, 30000, 101]
    utrs = [synth.make_utr(500 + i, n, long_utr=(n >= 30000)) for i, n in enumerate(sizes)]
    frames = [[(u.gene_info_str, synth.to_dataframe(u))] for u in utrs]      # one stream each
    got = fit_chunks(frames, seeds=[1] * len(utrs))
    for u, (g,) in zip(utrs, got):
        w = so.fit_utr(u.x, u.l, u.r, u.pa, np.random.RandomState(1))
        assert g.K == w.K and np.array_equal(g.alpha_arr, w.alpha_arr), (u.n_reads, g.alpha_arr, w.alpha_arr)
        assert np.allclose(g.beta_arr, w.beta_arr) and np.allclose(g.ws, w.ws, atol=1e-6)
        assert abs(g.lb_arr[-1] - w.lb_arr[-1]) <= 1e-6 * abs(w.lb_arr[-1])
        assert np.mean(g.label_arr == w.label_arr) >= 0.999
        assert len(g.label_arr) == u.n_reads and g.label_arr.dtype == np.int64


def test_heavy_tailed_waves_do_not_depend_on_their_composition():
    """cfg-3 shape: the scan cuts big UTRs into narrower chain sub-batches, dealt to separate CTAs,
    depending on what else is in the wave.  That may not change a single bit: every chain's sum
    runs over the same fragments in the same order.  Two heavy-tailed chunks fitted together == fitted one by one == one UTR at a time."""
    counts = np.minimum(synth.heavy_tail_read_counts(4000)[3000:3020], 40000)
    counts[3] = 25000
    counts[11] = 9000
    chunks = [[synth.make_utr(7000 + 10 * f + i, int(counts[10 * f + i])) for i in range(10)] for f in range(2)]
    frames = [[(u.gene_info_str, synth.to_dataframe(u)) for u in c] for c in chunks]
    together = fit_chunks(frames, seeds=[1, 1])
    for f in range(2):
        alone = fit_chunks([frames[f]], seeds=[1])[0]
        for a, b in zip(together[f], alone):
            assert a.K == b.K and np.array_equal(a.alpha_arr, b.alpha_arr) and np.array_equal(a.beta_arr, b.beta_arr)
            assert np.array_equal(a.ws, b.ws) and a.lb_arr == b.lb_arr and a.bic == b.bic
            assert np.array_equal(a.label_arr, b.label_arr)
    # first UTR of each chunk starts from the fresh seed: a single-UTR call must reproduce it exactly
    for f in range(2):
        one = fit_chunks([[frames[f][0]]], seeds=[1])[0][0]
        a = together[f][0]
        assert a.K == one.K and np.array_equal(a.ws, one.ws) and a.lb_arr == one.lb_arr


_KNOB_SCRIPT = r"""
import hashlib, sys
import numpy as np
sys.path.insert(0, %r)
from scape_b200 import synth
from scape_b200.apa_core import fit_chunks
sizes = [300, 40, 2500, 700, 12000, 150]
chunks = [[synth.make_utr(8000 + 10 * f + i, n) for i, n in enumerate(sizes)] for f in range(3)]
frames = [[(u.gene_info_str, synth.to_dataframe(u)) for u in c] for c in chunks]
h = hashlib.sha256()
for res in fit_chunks(frames, seeds=[1, 1, 1]):
    for r in res:
        for a in (np.asarray(r.alpha_arr), np.asarray(r.beta_arr), np.asarray(r.ws), np.asarray(r.lb_arr, dtype=float),
                  np.asarray(r.label_arr), np.asarray([r.K, r.L]), np.asarray([float(r.bic)])):
            h.update(np.ascontiguousarray(a).tobytes())
print("DIGEST", h.hexdigest())
"""


def test_scheduling_knobs_do_not_change_a_bit():
    """The wave pipelining, the pre-drawn chains, the cost-split scan work items, the pipelined E-step
    loads and the staged chain record only change WHEN and WHERE work runs: the results of a mixed
    batch (3 RNG streams, 40 ... 12,000 reads per UTR, prunes and refits included) must be bit-identical
    with each of them switched off.  The knobs are read once per process, hence one process each."""
    import os
    import subprocess
    import sys
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    # SCAPE_B200_POISON=1 fills the tensor arena with NaN bits before every wave: whatever the kernels
    # do not write themselves (pitch padding, slack rows the scan's prefetch ring touches) would then
    # poison the grid search -- the digest must not change either.
    # Three families of EM execution (api.cu run_chains): all steps bulk-synchronous (the default), a few
    # bulk-synchronous steps followed by the chain-resident kernel (SCAPE_B200_EM=tail), and the cluster-resident
    # kernel (SCAPE_B200_EM=cluster, any cluster size).  Their E passes sum a chain's fragments in
    # different orders (one warp / a CTA / G warps), so bits may differ BETWEEN the families (each
    # matches the oracle: test_scale_parity); within a family nothing may.
    families = {
        "bsp": ("", "SCAPE_B200_EM=bsp", "SCAPE_B200_PREDRAW=0", "SCAPE_B200_OVERLAP=0", "SCAPE_B200_POISON=1",
                "SCAPE_B200_SCAN_SPLIT=0", "SCAPE_B200_SCAN_TILES=1", "SCAPE_B200_SPLIT=1", "SCAPE_B200_STEP_EVENTS=1", "SCAPE_B200_ESTEP_FORK=0", "SCAPE_B200_DEFER=0",
                "SCAPE_B200_LANES=1", "SCAPE_B200_LANES=1 SCAPE_B200_SPLIT=1", "SCAPE_B200_LANES=3", "SCAPE_B200_LANES=2 SCAPE_B200_SPLIT=2",
                "SCAPE_B200_SCAN_CVT=2"),
        # warps per chain of the warp E-step kernel (default 2): another summation tree each
        "bsp_wpc1": ("SCAPE_B200_WARP_WPC=1", "SCAPE_B200_WARP_WPC=1 SCAPE_B200_SPLIT=1", "SCAPE_B200_WARP_WPC=1 SCAPE_B200_WARP_PF=0"),
        "bsp_wpc4": ("SCAPE_B200_WARP_WPC=4", "SCAPE_B200_WARP_WPC=4 SCAPE_B200_POISON=1"),
        "tail": ("SCAPE_B200_EM=tail", "SCAPE_B200_TAIL_CHAINS=100000", "SCAPE_B200_EM=tail SCAPE_B200_WARP_PF=0",
                 "SCAPE_B200_EM=tail SCAPE_B200_STAGE_CHAIN=0", "SCAPE_B200_EM=tail SCAPE_B200_POISON=1"),
        "cluster": ("SCAPE_B200_EM=cluster", "SCAPE_B200_EM=cluster SCAPE_B200_CLUSTER=1", "SCAPE_B200_EM=cluster SCAPE_B200_CLUSTER=4",
                    "SCAPE_B200_EM=cluster SCAPE_B200_POISON=1"),
    }
    for fam, knobs in families.items():
        digests = {}
        for knob in knobs:
            env = dict(os.environ)
            for kv in knob.split():
                k, v = kv.split("=")
                env[k] = v
            out = subprocess.run([sys.executable, "-c", _KNOB_SCRIPT % root], env=env, capture_output=True, text=True, timeout=300)
            assert out.returncode == 0, out.stderr[-2000:]
            digests[knob or "default"] = [l for l in out.stdout.splitlines() if l.startswith("DIGEST")][-1][7:19]
        assert len(set(digests.values())) == 1, (fam, digests)


def test_global_numpy_rng_is_consumed_like_the_reference():
    """subsample_run / infer use np.random's global legacy stream (apa_core.py:125); afterwards the
    stream must be where the reference would have left it."""
    from scape.apa_core import subsample_run
    u = synth.make_utr(3, 300)
    np.random.seed(1)
    res = subsample_run(data=synth.to_dataframe(u), gene_info_str=u.gene_info_str, n_max_apa=5)
    after = np.random.get_state()
    rng = np.random.RandomState(1)
    want = so.fit_utr(u.x, u.l, u.r, u.pa, rng)
    assert res.K == want.K and np.array_equal(res.alpha_arr, want.alpha_arr)
    assert np.array_equal(after[1], rng.get_state()[1]) and after[2] == rng.get_state()[2]


def test_cli_writes_reference_format_pickles(tmp_path, golden):
    """`scape infer_pa` end to end: file naming, TOML, pickle stream of scape.apa_core.Parameters."""
    import pickle
    from click.testing import CliRunner
    from scape.cli import cli
    us = [synth.make_utr(i, 300) for i in range(3)]
    paths = synth.write_chunk_files(us, str(tmp_path), per_file=100, stem="demo")
    synth.write_default_toml(str(tmp_path))
    r = CliRunner().invoke(cli, ["infer_pa", "--pkl_input_file", paths[0], "--output_dir", str(tmp_path)])
    assert r.exception is None, r.output
    out = tmp_path / "pkl_output" / "demo.100.1.1.res.pkl"
    assert out.exists()
    got = []
    with open(out, "rb") as fh:
        while True:
            try:
                got.append(pickle.load(fh))
            except EOFError:
                break
    assert len(got) == 3
    spec = golden["cases"]["synth8"]
    for i, g in enumerate(got):
        assert type(g).__module__ == "scape.apa_core" and type(g).__name__ == "Parameters"
        assert g.gene_info_str == us[i].gene_info_str
        assert np.array_equal(g.cb_id_arr, us[i].cb_id) and np.array_equal(g.readID_arr, us[i].read_id)
        assert g.alpha_arr.dtype == np.int64 and g.label_arr.dtype == np.int64
        check_against_golden(g, spec["utrs"][i], golden["labels"][f"synth8/{i}"], tight=True, lb_rtol=1e-7, ws_atol=1e-6)


def test_infer_files_on_the_gpu_equals_one_infer_pa_per_file(tmp_path):
    """The many-file entry point on real hardware, through the worker-process I/O path and with the files
    dealt over two handles (both on device 0 here): the result files must hold exactly what one
    `scape infer_pa`-style call per file gives (same streams, seed 1 per file)."""
    import pickle
    from scape_b200 import apa_core
    us = [synth.make_utr(9100 + i, 120 + 40 * (i % 5)) for i in range(24)]
    paths = synth.write_chunk_files(us, str(tmp_path), per_file=4, stem="many")
    outs = apa_core.infer_files(paths, str(tmp_path), devices=[0, 0], io_workers=2)
    assert len(outs) == 6

    def load(path):
        recs = []
        with open(path, "rb") as fh:
            while True:
                try:
                    recs.append(pickle.load(fh))
                except EOFError:
                    return recs

    for f, (pin, pout) in enumerate(zip(paths, outs)):
        want = fit_chunks([apa_core.read_chunk_file(pin)], seeds=[1])[0]
        got = load(pout)
        assert len(got) == len(want) == 4
        for g, w in zip(got, want):
            assert g.gene_info_str == w.gene_info_str and g.K == w.K and g.L == w.L and g.title == w.title
            assert np.array_equal(g.alpha_arr, w.alpha_arr) and np.array_equal(g.beta_arr, w.beta_arr)
            assert np.array_equal(g.ws, w.ws) and g.lb_arr == w.lb_arr and g.bic == w.bic
            assert np.array_equal(g.label_arr, w.label_arr) and g.label_arr.dtype == np.int64
            assert np.array_equal(g.cb_id_arr, w.cb_id_arr) and np.array_equal(g.readID_arr, w.readID_arr)
