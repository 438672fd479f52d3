"""Known-answer cases of the reference's kernel parity script (src/scape/taichi_code_test.py:514-593):
each scalar helper on the literal inputs it uses; expected values are the closed forms of the
reference's pure-Python twins (taichi_code_test.py:26-230), written out here with `math`."""
import math

import numpy as np
import pytest

from oracle import scape_oracle as so

SENT = float(np.finfo("f").min)
PI = 3.141592653589793


def test_sentinel_value():
    assert so.SENTINEL == -3.4028234663852886e38 == SENT


def test_logpdf_and_pdf_normal():            # taichi_code_test.py:527-535
    want = -0.5 * ((0.75 - 1.0) / 0.5) ** 2 - math.log(0.5) - 0.5 * math.log(2 * PI)
    assert so._log_normal(0.75, 1.0, 0.5) == pytest.approx(want, rel=1e-15)
    want = math.exp(-0.5 * ((0.75 - 1.0) / 0.5) ** 2) / math.sqrt(2 * PI) / 0.5
    assert float(so._pdf_normal(0.75, 1.0, 0.5)) == pytest.approx(want, rel=1e-15)


def test_logsumexp():                        # taichi_code_test.py:537-543
    x = [3.65, 7.89, 5., 6.12, 1.23]
    m = max(x)
    want = math.log(sum(math.exp(v - m) for v in x)) + m
    assert float(so._lse_rows(np.array([x]))[0]) == pytest.approx(want, rel=1e-15)


def test_loglik_l_xt():                      # taichi_code_test.py:545-561
    assert float(so._loglik_l_given_xt(np.array([30.]), np.array([50.]), 70)[0]) == SENT     # 50 > 40
    assert float(so._loglik_l_given_xt(np.array([5.]), np.array([50.]), 70)[0]) == pytest.approx(-math.log(65))
    assert float(so._lik_l_given_xt(np.array([30.]), np.array([50.]), 70)[0]) == 0.0
    assert float(so._lik_l_given_xt(np.array([5.]), np.array([50.]), 70)[0]) == pytest.approx(1 / 65)


def test_loglik_x_st():                      # taichi_code_test.py:563-576
    want = -0.5 * ((187 - 460 - 0) / 50) ** 2 - math.log(50) - 0.5 * math.log(2 * PI)
    got = so.loglik_xlr_t_pa([0.], [10.], [187.], 460, 50)[0] + math.log(460 - 0)
    assert got == pytest.approx(want, rel=1e-14)
    want = -0.5 * ((44 - (460 + 144 - 50)) / 50) ** 2 - math.log(50) - 0.5 * math.log(2 * PI)
    assert so._log_normal(44, 460 + 144 - 50, 50) == pytest.approx(want, rel=1e-15)


def test_r_known_matches_scalar_twin():      # taichi_code_test.py:285-309 on a hand-made case
    s_dis = np.arange(20, 150, 10)
    pmf = np.repeat(1 / 13, 13)
    x, l, r, theta, mu, sig = 100.0, 98.0, 47.0, 420.0, 300.0, 50.0
    vals, mass = [], 0.0
    for s, p in zip(s_dis, pmf):
        if s < r:
            vals.append(SENT)
            continue
        mass += p
        vals.append(-math.log(s) + (-0.5 * ((x - (theta + s - mu)) / sig) ** 2 - math.log(sig) - 0.5 * math.log(2 * PI))
                    + (-math.log(theta - x)) + math.log(p))
    m = max(vals)
    want = math.log(sum(math.exp(v - m) for v in vals)) + m - math.log(mass)
    got = so.loglik_xlr_t_r_known([x], [l], [r], s_dis, pmf, theta, mu, sig)[0]
    assert got == pytest.approx(want, rel=1e-14)


def test_r_unknown_matches_scalar_twin():    # taichi_code_test.py:353-365
    s_dis = np.arange(20, 150, 10)
    pmf = np.repeat(1 / 13, 13)
    for x, l, theta in ((100.0, 98.0, 420.0), (100.0, 98.0, 150.0), (0.0, 40.0, 2900.0)):
        acc = 0.0
        for s, p in zip(s_dis, pmf):
            lik_x = math.exp(-0.5 * ((x - (theta + s - 300.0)) / 50.0) ** 2) / math.sqrt(2 * PI) / 50.0
            lik_l = 1 / (theta - x) if l <= theta - x else 0.0
            acc += 1 / s * lik_x * lik_l * p
        if acc < 1e-300:
            acc = 0.0
        want = SENT if acc <= 0 else math.log(acc)
        got = so.loglik_xlr_t_r_unknown([x], [l], [np.nan], s_dis, pmf, theta, 300.0, 50.0)[0]
        assert got == pytest.approx(want, rel=1e-14)


def test_marginal_matches_scalar_twin():     # taichi_code_test.py:398-430, alpha=37 beta=5 on a 23-point grid (:737-760)
    rng = np.random.default_rng(0)
    theta = np.arange(1.0, 70.0, 3.0)
    table = rng.random((6, len(theta)))
    table[2, 11:14] = SENT
    alpha, beta = 37.0, 5.0
    sel = [i for i, t in enumerate(theta) if alpha - 3 * beta <= t <= alpha + 3 * beta]
    logp = [-0.5 * ((theta[i] - alpha) / beta) ** 2 - math.log(beta) - 0.5 * math.log(2 * PI) for i in sel]
    lps = math.log(sum(math.exp(v) for v in logp))
    got = so.loglik_marginal_lxr(alpha, beta, theta, table)
    for n in range(6):
        vals = [table[n, i] + lp - lps for i, lp in zip(sel, logp)]
        m = max(vals)
        assert got[n] == pytest.approx(math.log(sum(math.exp(v - m) for v in vals)) + m, rel=1e-14)
    assert so.marginal_window(theta, alpha, beta) == (sel[0], sel[-1])
