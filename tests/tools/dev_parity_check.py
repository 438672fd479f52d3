"""Development check: kernel-seam parity + whole-UTR parity against the oracle on a few UTRs."""
import os, sys, time, json
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np
from oracle import scape_oracle as so
from scape_b200 import _lib, synth
from scape_b200.apa_core import fit_chunks

P = _lib.make_params()
eng = _lib.Engine(P)
u = synth.make_utr(0, 300)
prm = dict(so.DEFAULTS); prm['utr_length'] = so.resolve_utr_length(u.x, u.l, prm)
m = so.build_model(u.x, u.l, u.r, u.pa, prm)
m.unif_loglik = so.uniform_loglik(m)
t0 = time.time(); tab = so.theta_table(m, m.theta); ten = so.get_loglik_marginal_tensor(m.theta, m.betas, tab); t1 = time.time()
gtab = eng.loglik_table(m.x, m.l, m.r, m.pa, m.theta)
fin = tab > -1e30
print("table: sentinel pattern equal", np.array_equal(fin, gtab > -1e30), "max rel err", np.max(np.abs(gtab[fin] - tab[fin]) / np.abs(tab[fin])))
gten = eng.marginal_tensor(m.theta, m.betas, tab)
fin = ten > -1e30
print("tensor: sentinel pattern equal", np.array_equal(fin, gten > -1e30), "max rel err", np.max(np.abs(gten[fin] - ten[fin]) / np.abs(ten[fin])), "oracle lik time %.2fs" % (t1 - t0))
# chains
m.tensor = ten
m.prof_x, m.prof_y = so.coverage_profile(m); m.peak_idx, m.peak_w = so.find_profile_peaks(m)
rng = np.random.RandomState(1)
chains, traces = [], []
for K in (5, 4, 3, 2, 1):
    for _ in range(3):
        ch = so.draw_chain(m, K, rng)
        init = dict(K=K, a_idx=ch.a_idx.copy(), b_idx=ch.b_idx.copy(), ws=ch.ws.copy())
        tr = []; m.trace = tr
        res = so.run_chain(m, ch, rng)
        init['k_order'] = res.k_order
        chains.append(init); traces.append((res, tr))
arr, (ta, tb, tw) = eng.em_chains(ten, m.cnt, m.unif_loglik, chains, trace=True)
for i, (io, (res, tr)) in enumerate(zip(arr, traces)):
    K = io.K
    ok_iter = io.n_iter == len(res.lb_arr)
    n = min(io.n_iter, len(res.lb_arr))
    lb_err = max(abs(io.lb_arr[j] - res.lb_arr[j]) / abs(res.lb_arr[j]) for j in range(n))
    same_path = all(list(ta[i, j, :K]) == list(tr[j]['a_idx']) and list(tb[i, j, :K]) == list(tr[j]['b_idx']) for j in range(n))
    print("chain", i, "K", K, "iters", io.n_iter, len(res.lb_arr), "lb relerr %.2e" % lb_err, "path same", same_path,
          "bic relerr %.2e" % (abs(io.bic - res.bic) / abs(res.bic)), "ws err %.2e" % max(abs(io.ws[j] - res.ws[j]) for j in range(K + 1)))
# whole UTRs
us = [synth.make_utr(i, 300) for i in range(6)]
chunk = [(u.gene_info_str, synth.to_dataframe(u)) for u in us]
t0 = time.time(); (got,), raw = fit_chunks([chunk], seeds=[1], engine=eng, return_raw=True); t1 = time.time()
print("gpu fit 6 UTRs: %.3fs" % (t1 - t0), json.dumps(raw.timing))
rng = np.random.RandomState(1)
for u, g in zip(us, got):
    w = so.fit_utr(u.x, u.l, u.r, u.pa, rng)
    print(g.K, w.K, g.alpha_arr, w.alpha_arr, g.beta_arr, w.beta_arr, "ws err %.2e" % np.max(np.abs(g.ws - w.ws)) if g.K == w.K else "K!",
          "lb %.3e" % (abs(g.lb_arr[-1] - w.lb_arr[-1]) / abs(w.lb_arr[-1])), len(g.lb_arr), len(w.lb_arr), "labels", np.mean(g.label_arr == w.label_arr), w.path)
