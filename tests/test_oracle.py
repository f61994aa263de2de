"""CPU: the oracle against the golden vectors generated from the real reference
(tests/golden/make_golden.py), and against the published table rows."""
import json
import os

import numpy as np
import pytest

import helpers
from helpers import orc

FAST_CASES = [c for c, m in helpers.cases().items() if m["n"] * m["max_iter"] <= 1_000_000 and helpers.tier(c) == "full"]
PREFIX_FAST = [c for c, m in helpers.cases().items() if helpers.tier(c) == "prefix" and m["n"] * m["max_iter"] <= 400_000]


@pytest.mark.parametrize("case", FAST_CASES)
def test_oracle_matches_reference_goldens(case):
    """Same machine image => the oracle reproduces the stored reference histories; across CPU
    models OpenBLAS may pick another ddot kernel, so the assertion is the parity rule (and the
    bit-for-bit count is reported)."""
    A, b, x0, x_true, dinv, max_iter = helpers.case_problem(case)
    exact_bits = 0
    for tag in orc.VARIANTS:
        out = orc.solve(tag, A, b, x0, max_iter, dinv=dinv, x_true=x_true)
        ref = {h: helpers.golden_history(case, tag, h) for h in orc.HISTORIES}
        assert out["name"] == orc.VARIANTS[tag] and out["max_iter"] == max_iter
        helpers.check_parity(out, ref, helpers.cases()[case]["kstar"][tag], f"{case}/{tag}")
        exact_bits += all(np.array_equal(out[h], ref[h], equal_nan=True) for h in orc.HISTORIES)
    print(f"{case}: {exact_bits}/9 variants bit-identical to the stored reference run")


@pytest.mark.parametrize("case", PREFIX_FAST)
def test_oracle_matches_reference_prefixes(case):
    """The figure_gen.py cases stored as prefixes: the oracle reproduces the stored start of both
    residual histories (P1) and the reference's summary metrics of the same run."""
    A, b, x0, x_true, dinv, max_iter = helpers.case_problem(case)
    meta = helpers.cases()[case]
    for tag in ("hs", "pr", "pipe_pr", "gv"):
        out = orc.solve(tag, A, b, x0, max_iter, dinv=dinv, x_true=x_true)
        ref = {h: helpers.golden_history(case, tag, h) for h in helpers.RESIDUAL_HISTS}
        helpers.check_window(out, ref, meta["kstar"][tag], f"{case}/{tag}")
        helpers.check_metrics(out, meta["kstar"][tag], f"{case}/{tag}")


def test_every_figure_gen_case_has_a_fixture():
    """figure_gen.py:247-315 restricted to the matrices present: 43 (matrix, preconditioner) pairs."""
    c = helpers.cases()
    listed = [k for k in c if not k.startswith("poisson")]
    assert len(listed) == 43
    assert sum(helpers.tier(k) == "metrics" for k in c) == 10
    for k, m in c.items():
        if m["kstar"] is not None:
            for tag, band in m["kstar"].items():
                if "ensemble11" in band:
                    assert band["window"] == (band["ensemble11"] if band["kstar11"] is None else min(band["kstar11"], band["ensemble11"]))
                    assert band["window"] <= band["ensemble"] and (band["kstar10"] is None or band["window"] <= band["kstar10"])


def test_goldens_match_reference_first_values():
    """The KAT quoted in SURVEY.md section 8c (bcsstk03 + Jacobi, hs_pcg)."""
    r = helpers.golden_history("bcsstk03_jacobi", "hs", "updated_residual_2_norm")
    np.testing.assert_allclose(r[:4], [2.641158787675e10, 1.950679607079e9, 8.103599327493e8,
                                       7.571925366651e8], rtol=1e-12)
    e = helpers.golden_history("bcsstk03_jacobi", "hs", "error_A_norm")
    np.testing.assert_allclose(e[:3], [84328.24630597, 29789.74779395, 13699.54548127], rtol=1e-12)


def test_published_table_rows_coarse():
    """figures/convergence_table_data.tex (2019 numpy/MKL) vs today's reference run: the
    coarse regression band of SURVEY.md section 8c (iterations +-max(2,2%), accuracy x10^0.6)
    on the well-conditioned rows present in both."""
    table = json.load(open(os.path.join(helpers.GOLDEN, "table.json")))
    cols = ["hs", "cg", "m", "pr", "gv", "pipe_pr_m", "pipe_pr"]      # figure_gen.py:360
    checked = 0
    for case in ("nos4_jacobi", "model_48_8_3_None", "bcsstk03_jacobi", "494_bus_jacobi"):
        row = table[case]
        for j, tag in enumerate(cols):
            it, acc = orc.convergence_metrics(helpers.golden_history(case, tag, "error_A_norm"))
            assert abs(it - row["iters"][j]) <= max(2, 0.02 * row["iters"][j]), (case, tag, it, row["iters"][j])
            if tag != "gv":
                assert abs(acc - row["acc"][j]) <= 0.6, (case, tag, acc, row["acc"][j])
            checked += 1
    assert checked == 28


def test_poisson_generators_match_reference_matrix():
    """matrices/poisson_ca.mtx is the 16x16 5-point Laplacian."""
    P = helpers.load_matrix("poisson_ca")
    Q = orc.poisson2d(16)
    assert (P != Q).nnz == 0
    assert np.array_equal(P.indices, Q.indices) and np.array_equal(P.data, Q.data)


def test_mpi_style_kats():
    kat = json.load(open(os.path.join(helpers.GOLDEN, "mpi_kat.json")))
    for tag, ent in kat.items():
        n = ent["n"]
        lam = orc.model_problem_spectrum(n)
        x = orc.solve_mpi_style(tag, lam, lam / np.sqrt(n), ent["max_iter"])
        err = float(np.linalg.norm(np.ones(n) / np.sqrt(n) - x))
        assert 0.25 <= err / ent["error"] <= 4.0, (tag, err, ent["error"])


def test_departure_index_handles_early_break():
    ex = np.array([1.0, 0.5, 0.25, 0.0, 0.0])
    ref = np.array([1.0, 0.5, 0.25 * (1 + 1e-9), 0.1, 0.05])
    assert orc.departure_index(ref, ex) == 2
    assert orc.departure_index(np.array([1.0, 0.5, 0.25, 0.1]), ex) == 3
