"""GPU (-m gpu): parity of the CUDA path, called through the reference-shaped Python
functions and the C ABI, against the oracle run live on the same inputs and against the
golden fixtures.  Nothing here reads /root/reference."""
import json
import os

import numpy as np
import pytest

import helpers
from helpers import orc
from new_cg_variants_b200 import PoissonStencil, Session, callbacks as cbk, cg_variants

pytestmark = pytest.mark.gpu

STD_CALLBACKS = [cbk.error_A_norm, cbk.residual_2_norm, cbk.error_2_norm, cbk.updated_residual_2_norm]
ALL_TAGS = list(orc.VARIANTS)


# ------------------------------------------------------------------------------- primitives
@pytest.mark.parametrize("name", helpers.matrix_names())
def test_csr_spmv_bitwise_scipy(name):
    """y = A v for every matrix of predict_and_recompute/matrices: bit-identical to scipy."""
    A = helpers.load_matrix(name)
    v = np.random.default_rng(7).standard_normal(A.shape[0])
    with Session(A) as s:
        y = s.spmv(v)
    assert np.array_equal(y, A @ v)


@pytest.mark.parametrize("shape", [(16, 16, 1), (33, 7, 1), (6, 5, 4), (12, 12, 12), (64, 3, 5), (1, 1, 9),
                                   (256, 20, 1), (130, 9, 3), (258, 17, 9), (2, 3, 70), (128, 8, 33)])
def test_stencil_spmv_bitwise_csr(shape):
    """Both stencil SpMV implementations (TMA-staged for even nx, generic otherwise and when
    forced) reproduce scipy's product with the equivalent CSR matrix bit for bit."""
    nx, ny, nz = shape
    S = PoissonStencil(nx, ny, nz, dim=2 if nz == 1 else 3)
    v = np.random.default_rng(3).standard_normal(S.shape[0])
    ref = S.tocsr() @ v
    for tma in (1, 0):
        with Session(S) as s:
            s.set_option("tma", tma)
            assert np.array_equal(s.spmv(v), ref), (shape, tma)


def test_dot_deterministic_and_accurate():
    rng = np.random.default_rng(11)
    A = helpers.load_matrix("nos4")
    with Session(A) as s:
        for n in (1, 2, 255, 4097, 1_000_003):
            u, v = rng.standard_normal(n), rng.standard_normal(n)
            d1, d2 = s.dot(u, v), s.dot(u, v)
            assert d1 == d2                                    # bitwise repeatable
            exact = float(u.astype(np.longdouble) @ v.astype(np.longdouble))
            scale = float(np.abs(u) @ np.abs(v))
            assert abs(d1 - exact) <= 8 * np.finfo(float).eps * scale


# --------------------------------------------------------------------- variants vs oracle
def _device_solve(tag, A, b, x0, max_iter, dinv, x_true, **kw):
    f = getattr(cg_variants, orc.VARIANTS[tag])
    prec = (lambda v: v) if dinv is None else (lambda v: dinv * v)
    return f(A, b, x0, max_iter, preconditioner=prec, callbacks=STD_CALLBACKS, x_true=x_true, **kw)


@pytest.mark.parametrize("path", ["stream", "persistent"])
@pytest.mark.parametrize("case", helpers.cases_of("full", "prefix"))
def test_variants_match_oracle_and_goldens(case, path):
    """Every variant on every figure_gen.py case short enough for a live oracle run, through both
    execution paths: `stream` (1-3 fused kernels per iteration) and `persistent` (one cooperative
    launch for the whole solve).  P1-P3 against the oracle run live and against the stored
    reference histories (whole histories for tier "full", their first entries for tier "prefix")."""
    A, b, x0, x_true, dinv, max_iter = helpers.case_problem(case)
    meta = helpers.cases()[case]
    bands = meta["kstar"]
    full = helpers.tier(case) == "full"
    report = []
    for tag in ALL_TAGS:
        dev = _device_solve(tag, A, b, x0, max_iter, dinv, x_true, path=path, return_info=True)
        assert dev["_info"]["path"] == {"stream": 1, "persistent": 2}[path]
        assert dev["_info"]["kernel_launches"] > 0
        assert dev["name"] == orc.VARIANTS[tag] and dev["max_iter"] == max_iter
        live = orc.solve(tag, A, b, x0, max_iter, dinv=dinv, x_true=x_true)
        for h in orc.HISTORIES:
            assert dev[h].shape == (max_iter,)
            # k = 0: one reduction, no recurrence yet -> agreement at rounding level
            np.testing.assert_allclose(dev[h][:1], live[h][:1], rtol=1e-13, err_msg=f"{case}/{tag}/{h}")
        kd = min(helpers.first_deviation(dev[h], live[h]) for h in helpers.RESIDUAL_HISTS)
        it, acc = orc.convergence_metrics(dev["error_A_norm"])
        helpers.log_kd(case=case, variant=tag, path=path, kd=kd, window=bands[tag]["window"],
                       kstar10=bands[tag]["kstar10"], ensemble=bands[tag].get("ensemble"),
                       kstar11=bands[tag].get("kstar11"), ensemble11=bands[tag].get("ensemble11"), max_iter=max_iter,
                       iters=it, acc=acc, iters_band=bands[tag]["iters_band"], acc_band=bands[tag]["acc_band"])
        helpers.check_parity(dev, live, bands[tag], f"{case}/{tag} vs live oracle")
        gold = {h: helpers.golden_history(case, tag, h) for h in (orc.HISTORIES if full else helpers.RESIDUAL_HISTS)}
        helpers.check_window(dev, gold, bands[tag], f"{case}/{tag} vs golden")
        report.append(f"{tag}: agree<1e-10 to k={kd} (window {bands[tag]['window']}, k*10 {bands[tag]['kstar10']}) it={it} acc={acc:.2f}")
    print(f"\n[{case}/{path}] " + " | ".join(report))


@pytest.mark.parametrize("case", helpers.cases_of("metrics"))
def test_long_runs_match_reference_metrics(case):
    """The long figure_gen.py cases (5 000 ... 1 750 000 iterations; the persistent kernel runs each
    solve in one launch).  No oracle run here (minutes of interpreter time each): P2 / P3 against
    the band the reference itself spans under re-ordered inner products (make_golden.py ran it),
    P1 against the stored window prefix is not applicable (tier "metrics" stores no histories);
    bcsstk18_None / bcsstm25_None are pinned only by the reference's own stored results."""
    A, b, x0, x_true, dinv, max_iter = helpers.case_problem(case)
    meta = helpers.cases()[case]
    tags = ALL_TAGS if max_iter <= 200_000 else ["hs", "pr", "pipe_pr"]
    table = json.load(open(os.path.join(helpers.GOLDEN, "table.json"))).get(case)
    cols = ["hs", "cg", "m", "pr", "gv", "pipe_pr_m", "pipe_pr"]          # figure_gen.py:360
    for tag in tags:
        dev = _device_solve(tag, A, b, x0, max_iter, dinv, x_true, return_info=True)
        assert dev["_info"]["path"] == 2                                   # latency-bound: persistent kernel
        it, acc = orc.convergence_metrics(dev["error_A_norm"])
        band = meta["kstar"][tag] if meta["kstar"] else None
        stored = (meta.get("stored_metrics") or {}).get(tag)
        pub = (table["iters"][cols.index(tag)], table["acc"][cols.index(tag)]) if table and tag in cols else None
        helpers.log_kd(case=case, variant=tag, path="persistent", kd=None, window=None, kstar10=None, ensemble=None,
                       max_iter=max_iter, iters=it, acc=acc, iters_band=band and band["iters_band"],
                       acc_band=band and band["acc_band"], stored=stored, published=pub)
        if band:
            helpers.check_metrics(dev, band, f"{case}/{tag}")
        # coarse regression against the reference's 2019 runs: attainable accuracy
        # within one decade, iterations-to-1e-5 within 5 % when both reached it -- each widened by the
        # spread the reference itself shows under re-ordered dots (e.g. bcsstm24 GV: 1987 ... 34297
        # iterations across the ensemble; the 2019 run took 19411)
        aw = (band["acc_band"][1] - band["acc_band"][0]) if band else 0.0
        iw = (band["iters_band"][1] - band["iters_band"][0]) if band and min(band["iters_band"]) > 0 else 0
        for src, ref in (("stored .npy", stored), ("published table", pub)):
            if ref is None:
                continue
            assert abs(acc - ref[1]) <= 1.0 + aw, (case, tag, src, acc, ref)
            if ref[0] and it:
                assert abs(it - ref[0]) <= max(3, 0.05 * ref[0]) + iw, (case, tag, src, it, ref)


def test_unpreconditioned_twins_and_names():
    A = helpers.load_matrix("nos4")
    x_true, b, x0 = orc.setup_problem(A)
    for stem in ("hs", "cg", "gv", "pr", "m", "pipe_pr", "pipe_p"):
        out_cg = getattr(cg_variants, stem + "_cg")(A, b, x0, 60, callbacks=STD_CALLBACKS, x_true=x_true)
        out_pcg = getattr(cg_variants, stem + "_pcg")(A, b, x0, 60, callbacks=STD_CALLBACKS, x_true=x_true)
        assert out_cg["name"] == stem + "_cg" and out_pcg["name"] == stem + "_pcg"
        for h in orc.HISTORIES:
            assert np.array_equal(out_cg[h], out_pcg[h]), (stem, h)


def test_stencil_solve_equals_csr_solve():
    """Matrix-free operator vs the CSR matrix it stands for: the SpMV results are bit-identical
    (checked in test_stencil_spmv_bitwise_csr); the fused inner products are summed in a
    different (kernel-specific) fixed order, so the solves agree to rounding, not bitwise."""
    for S in (PoissonStencil(24, 20, 1, dim=2), PoissonStencil(10, 9, 8, dim=3)):
        A = S.tocsr()
        x_true, b, x0 = orc.setup_problem(A)
        assert np.array_equal(S @ x_true, b)
        dinv = 1 / A.diagonal()
        for tag in ALL_TAGS:
            d_s = _device_solve(tag, S, b, x0, 40, dinv, x_true)
            d_c = _device_solve(tag, A, b, x0, 40, dinv, x_true)
            for h in orc.HISTORIES:
                np.testing.assert_allclose(d_s[h][:12], d_c[h][:12], rtol=1e-10, err_msg=f"{tag}/{h}")
                # the tail is converged to rounding level: absolute floor relative to k = 0
                np.testing.assert_allclose(d_s[h], d_c[h], rtol=1e-6, atol=1e-11 * d_c[h][0],
                                           err_msg=f"{tag}/{h}")


def test_runs_are_bitwise_repeatable():
    A, b, x0, x_true, dinv, max_iter = helpers.case_problem("bcsstk15_jacobi")
    for tag in ("hs", "pr", "pipe_pr", "gv"):
        o1 = _device_solve(tag, A, b, x0, 200, dinv, x_true)
        o2 = _device_solve(tag, A, b, x0, 200, dinv, x_true)
        for h in orc.HISTORIES:
            assert np.array_equal(o1[h], o2[h], equal_nan=True)


def test_history_subset_and_missing_x_true():
    A = helpers.load_matrix("bcsstk03")
    x_true, b, x0 = orc.setup_problem(A)
    out = cg_variants.pr_pcg(A, b, x0, 30, callbacks=[cbk.updated_residual_2_norm])
    assert set(out) == {"name", "max_iter", "updated_residual_2_norm"}
    ref = orc.solve("pr", A, b, x0, 30, x_true=x_true)
    np.testing.assert_allclose(out["updated_residual_2_norm"][:8], ref["updated_residual_2_norm"][:8], rtol=1e-10)
    # no x_true given: solved for on the host as the reference callback does
    out = cg_variants.hs_pcg(A, b, x0, 10, callbacks=[cbk.error_A_norm])
    np.testing.assert_allclose(out["error_A_norm"][0], ref["error_A_norm"][0], rtol=1e-6)


def test_generic_callbacks_stepwise_protocol():
    """save_x / a user callable force the stepwise path: same histories, x_k per iteration."""
    A = helpers.load_matrix("nos4")
    x_true, b, x0 = orc.setup_problem(A)
    dinv = 1 / A.diagonal()
    seen = []

    def spy(**kw):
        seen.append((kw["k"], kw["a_k1"], kw["b_k"], float(np.linalg.norm(kw["r_k"]))))

    for tag in ("hs", "pr", "pipe_pr"):
        seen.clear()
        f = getattr(cg_variants, orc.VARIANTS[tag])
        fast = f(A, b, x0, 25, preconditioner=lambda v: dinv * v, callbacks=STD_CALLBACKS, x_true=x_true)
        slow = f(A, b, x0, 25, preconditioner=lambda v: dinv * v,
                 callbacks=STD_CALLBACKS + [cbk.save_x, cbk.save_r, spy], x_true=x_true)
        for h in orc.HISTORIES:
            assert np.array_equal(fast[h], slow[h])
        assert slow["x"].shape == (25, 100) and [s[0] for s in seen] == list(range(25))
        ref = orc.solve(tag, A, b, x0, 25, dinv=dinv, x_true=x_true, return_state=True)
        np.testing.assert_allclose(slow["x"][-1], ref["_state"]["x"], rtol=1e-9, atol=1e-14)
        np.testing.assert_allclose([s[3] for s in seen], slow["updated_residual_2_norm"], rtol=1e-14)
        assert seen[0][1] == 0.0 and seen[1][1] != 0.0


def test_state_vectors_after_one_iteration():
    """One loop trip of every variant against the oracle's vectors (elementwise and SpMV
    steps are bit-exact; the scalars differ only by the summation order of the dots)."""
    A = helpers.load_matrix("bcsstk03")
    x_true, b, x0 = orc.setup_problem(A)
    dinv = 1 / A.diagonal()
    with Session(A, dinv=dinv) as s:
        s.load_problem(b, x0, x_true)
        for tag in ALL_TAGS:
            ref = orc.solve(tag, A, b, x0, 2, dinv=dinv, x_true=x_true, return_state=True)["_state"]
            s.run(tag, 2)
            for name in ("x", "r"):
                np.testing.assert_allclose(s.vector(name), ref[name], rtol=1e-11,
                                           atol=1e-14 * np.abs(ref[name]).max(), err_msg=f"{tag}/{name}")
            sc = s.scalars()
            np.testing.assert_allclose(sc["a"], ref["a"], rtol=1e-12, err_msg=tag)
            np.testing.assert_allclose(sc["nu"], ref["nu"], rtol=1e-12, err_msg=tag)


def test_breakdown_is_reported_not_hidden():
    """bcsstm21 is diagonal with few distinct values: CG terminates exactly and 0/0 follows,
    in the reference as here (figure_gen.py:302 runs it for 10 iterations)."""
    A = helpers.load_matrix("bcsstm21")
    x_true, b, x0 = orc.setup_problem(A)
    out = cg_variants.hs_pcg(A, b, x0, 10, callbacks=STD_CALLBACKS, x_true=x_true, return_info=True)
    ref = orc.solve("hs", A, b, x0, 10, x_true=x_true)
    k_nan_ref = int(np.argmax(~np.isfinite(ref["updated_residual_2_norm"]))) if not np.all(np.isfinite(ref["updated_residual_2_norm"])) else None
    k_nan_dev = int(np.argmax(~np.isfinite(out["updated_residual_2_norm"]))) if not np.all(np.isfinite(out["updated_residual_2_norm"])) else None
    if k_nan_ref is not None:
        assert k_nan_dev is not None and out["_info"]["breakdown_iter"] >= 0
    np.testing.assert_allclose(out["error_A_norm"][:2], ref["error_A_norm"][:2], rtol=1e-10)


# ----------------------------------------------------------- BASELINE-size property checks
@pytest.mark.parametrize("tag", ["hs", "pr", "pipe_pr", "gv", "cg"])
def test_poisson3d_256_properties(tag):
    """configs[3] size (16.8 M unknowns): no oracle run at this size, so size-independent
    properties: (i) the recursively updated residual equals the true residual b - A x_k to
    rounding while far from convergence, (ii) the A-norm error decreases monotonically,
    (iii) agreement with the oracle on the first iterations of the SAME problem is implied by
    (iv) bitwise equality of stencil and CSR paths checked above at small size; here we
    check (v) the first history entries against their closed forms."""
    S = PoissonStencil(256, 256, 256, dim=3)
    n = S.shape[0]
    x_true = np.ones(n) / np.sqrt(n)
    b = S @ x_true
    x0 = np.zeros(n)
    dinv = np.full(n, 1.0 / 6.0)
    out = _device_solve(tag, S, b, x0, 40, dinv, x_true, return_info=True)
    r, ur, ea = out["residual_2_norm"], out["updated_residual_2_norm"], out["error_A_norm"]
    assert np.all(np.isfinite(r)) and np.all(np.isfinite(ea))
    np.testing.assert_allclose(r[0], np.linalg.norm(b), rtol=1e-13)
    np.testing.assert_allclose(ea[0], np.sqrt(x_true @ b), rtol=1e-13)
    np.testing.assert_allclose(out["error_2_norm"][0], 1.0, rtol=1e-13)
    np.testing.assert_allclose(ur, r, rtol=1e-9)
    assert np.all(np.diff(ea) < 0)
    assert out["_info"]["kernel_launches"] > 0 and out["_info"]["loop_ms"] > 0


def test_poisson2d_4096_first_iterations_match_small_oracle_structure():
    """configs[2] size: 4096^2 5-point; same property checks, PR-CG."""
    S = PoissonStencil(4096, 4096, 1, dim=2)
    n = S.shape[0]
    x_true = np.ones(n) / np.sqrt(n)
    b = S @ x_true
    out = _device_solve("pr", S, b, np.zeros(n), 30, np.full(n, 0.25), x_true)
    np.testing.assert_allclose(out["updated_residual_2_norm"], out["residual_2_norm"], rtol=1e-9)
    assert np.all(np.diff(out["error_A_norm"]) < 0)


@pytest.mark.parametrize("shape", [(256, 16, 1), (130, 9, 1), (64, 24, 20), (258, 10, 7), (12, 12, 12), (2, 3, 40)])
def test_tma_stencil_path_equals_generic_path(shape):
    """The TMA-staged stencil kernels (even nx) against the generic one-row-per-thread
    stencil kernel for every variant: partial tiles in x/y, several z-chunks, 2-D and 3-D.
    Row sums are bit-identical; the fused dots use another fixed summation order, hence
    agreement to rounding (1e-11 after 12 iterations), and bitwise repeatability per path."""
    nx, ny, nz = shape
    S = PoissonStencil(nx, ny, nz, dim=2 if nz == 1 else 3)
    n = S.shape[0]
    x_true = np.ones(n) / np.sqrt(n)
    b = S @ x_true
    x0 = np.zeros(n)
    for dinv in (None, 1 / S.diagonal(), 1 / (S.diagonal() + np.arange(n) % 3)):
        res = {}
        for tma in (1, 0):
            with Session(S, dinv=dinv) as s:
                s.set_option("tma", tma)
                for tag in ALL_TAGS:
                    x, hist, info = s.solve(tag, b, x0, 12, x_true=x_true, path="stream")
                    res[(tma, tag)] = (x, hist)
        for tag in ALL_TAGS:
            x1, h1 = res[(1, tag)]
            x0_, h0 = res[(0, tag)]
            np.testing.assert_allclose(x1, x0_, rtol=1e-10, atol=1e-13, err_msg=f"{shape}/{tag}")
            for h in orc.HISTORIES:
                np.testing.assert_allclose(h1[h], h0[h], rtol=1e-10, err_msg=f"{shape}/{tag}/{h}")


# ------------------------------------------------------------------ persistent vs stream
def test_persistent_path_equals_stream_path():
    """Same per-row arithmetic, another fixed summation order of the fused dots: agreement
    to rounding over the first iterations; each path bitwise repeatable; AUTO picks the
    persistent kernel for latency-bound sizes and the stream kernels for large ones."""
    S = PoissonStencil(24, 20, 12, dim=3)
    A = S.tocsr()
    x_true, b, x0 = orc.setup_problem(A)
    for op, dinv in ((S, 1 / A.diagonal()), (A, 1 / (A.diagonal() + np.arange(A.shape[0]) % 3)), (A, None)):
        for tag in ALL_TAGS:
            o_s = _device_solve(tag, op, b, x0, 30, dinv, x_true, path="stream")
            o_p = _device_solve(tag, op, b, x0, 30, dinv, x_true, path="persistent")
            o_p2 = _device_solve(tag, op, b, x0, 30, dinv, x_true, path="persistent")
            for h in orc.HISTORIES:
                np.testing.assert_allclose(o_p[h][:12], o_s[h][:12], rtol=1e-10, err_msg=f"{tag}/{h}")
                assert np.array_equal(o_p[h], o_p2[h], equal_nan=True)
    auto = _device_solve("pr", S, b, x0, 10, None, x_true, return_info=True)
    assert auto["_info"]["path"] == 2
    with Session(S) as s:
        s.set_option("persistent_threshold", 100)
        _, _, info = s.solve("pr", b, x0, 10, x_true=x_true)
        assert info["path"] == 1


def test_persistent_path_long_run_and_resume():
    """1250 iterations of un-preconditioned bcsstk03 in ONE launch, and the same run cut
    into pieces with cgx_advance (the scalars survive the kernel boundary bit for bit)."""
    A, b, x0, x_true, dinv, max_iter = helpers.case_problem("bcsstk03_None")
    with Session(A) as s:
        s.load_problem(b, x0, x_true)
        info = s.run("pipe_pr", max_iter, histories=orc.HISTORIES, path="persistent")
        assert info["path"] == 2 and info["kernel_launches"] < 40
        _, h_one = s.fetch()
        s.begin("pipe_pr", max_iter, histories=orc.HISTORIES, path="persistent")
        for chunk in (1, 7, 300, 5000):
            s.advance(chunk)
        _, h_cut = s.fetch()
        assert np.array_equal(h_one, h_cut, equal_nan=True)


@pytest.mark.parametrize("name", ["bcsstk16", "bcsstk18", "nos7", "bcsstm24", "model_48_8_3"])
def test_csr_stream_kernel_equals_row_kernel(name):
    """CSR-stream SpMV (CTA-wide coalesced sweep, products staged in shared memory, row sums in
    stored order) against the one-thread-per-row kernel: bit-identical products, solves equal to rounding."""
    A = helpers.load_matrix(name)
    x_true, b, x0 = orc.setup_problem(A)
    dinv = orc.jacobi_dinv(A)
    res = {}
    for flag in (1, 0):
        with Session(A, dinv=dinv) as s:
            s.set_option("csr_stream", flag)
            for tag in ("hs", "cg", "pr", "pipe_pr", "gv"):
                x, hist, _ = s.solve(tag, b, x0, 10, x_true=x_true, path="stream")
                res[(flag, tag)] = (x, hist)
            v = np.random.default_rng(3).standard_normal(A.shape[0])
            res[(flag, "spmv")] = s.spmv(v)
            assert np.array_equal(res[(flag, "spmv")], A @ v)            # scipy's csr_matvec, bit for bit
    for tag in ("hs", "cg", "pr", "pipe_pr", "gv"):
        for h in orc.HISTORIES:      # same row sums; the fused dots are summed in another fixed order
            np.testing.assert_allclose(res[(1, tag)][1][h][:6], res[(0, tag)][1][h][:6], rtol=1e-9,
                                       atol=1e-13 * res[(0, tag)][1][h][0], err_msg=f"{tag}/{h}")    # (diagonal matrices converge at once)


def _ragged_long_rows():
    """Rows of 0, 1, 1119, 1120 (= one work item of the bulk kernel exactly), 1121 and 3000 non-zeros
    between short ones; diagonally dominant where a diagonal exists."""
    import scipy.sparse as sps
    rng = np.random.default_rng(11)
    n = 3200
    lens = rng.integers(1, 9, n)
    for r, ln in ((5, 1120), (6, 1121), (7, 1119), (900, 3000), (901, 0), (902, 0), (2500, 2241), (n - 1, 1500)):
        lens[r] = ln
    rows, cols, vals = [], [], []
    for r in range(n):
        c = np.sort(rng.choice(n, size=lens[r], replace=False)) if lens[r] else np.zeros(0, int)
        rows.append(np.full(lens[r], r)); cols.append(c); vals.append(rng.standard_normal(lens[r]))
    return sps.csr_matrix((np.concatenate(vals), (np.concatenate(rows), np.concatenate(cols))), shape=(n, n))


@pytest.mark.parametrize("name", ["bcsstk16", "bcsstk18", "nos7", "bcsstm24", "model_48_8_3", "494_bus", "ragged_long"])
def test_csr_bulk_kernel_equals_stream_kernel(name):
    """csr_bulk_kernel (matrix stream staged by cp.async.bulk, ring of slots) against csr_stream_kernel
    and scipy: bit-identical products for several ring depths, solves equal to rounding (the fused
    dots are summed over other work items)."""
    if name == "ragged_long":
        A = _ragged_long_rows()
        with Session(A) as s:
            s.set_option("csr_bulk", 2)
            for ring in (0, 2, 3, 7):
                s.set_option("csr_bulk_ring", ring)
                for seed in (1, 2):
                    v = np.random.default_rng(seed).standard_normal(A.shape[0])
                    assert np.array_equal(s.spmv(v), A @ v), f"ring {ring}"
            s.set_option("csr_bulk", 0)
            assert np.array_equal(s.spmv(v), A @ v)
        return
    A = helpers.load_matrix(name)
    x_true, b, x0 = orc.setup_problem(A)
    dinv = orc.jacobi_dinv(A)
    res = {}
    for flag, ring in ((1, 0), (1, 3), (0, 0)):     # (rings that leave two CTAs per SM: same grid)
        with Session(A, dinv=dinv) as s:
            s.set_option("csr_bulk", 2 * flag)             # (2: also the two-right-hand-side pass)
            s.set_option("csr_bulk_ring", ring)
            s.set_option("csr_bulk_sum", 2)                # (the CTA width enters the order of the fused dots)
            for tag in ("hs", "cg", "pr", "pipe_pr", "gv", "m", "pipe_p"):
                x, hist, _ = s.solve(tag, b, x0, 10, x_true=x_true, path="stream")
                res[(flag, ring, tag)] = (x, hist)
            v = np.random.default_rng(3).standard_normal(A.shape[0])
            assert np.array_equal(s.spmv(v), A @ v)            # scipy's csr_matvec, bit for bit
    for tag in ("hs", "cg", "pr", "pipe_pr", "gv", "m", "pipe_p"):
        for h in orc.HISTORIES:
            assert np.array_equal(res[(1, 0, tag)][1][h], res[(1, 3, tag)][1][h]), f"{tag}/{h}: ring depth changed the bits"
            np.testing.assert_allclose(res[(1, 0, tag)][1][h][:6], res[(0, 0, tag)][1][h][:6], rtol=1e-9,
                                       atol=1e-13 * res[(0, 0, tag)][1][h][0], err_msg=f"{tag}/{h}")


# ------------------------------------------------ callers either side of the path
def test_figure_gen_style_driver_end_to_end(tmp_path):
    """experiments.test_matrix / parse_convergence_data (figure_gen.py:21-124 restated) on the
    GPU path: .npy dictionaries in the reference's layout, table row within the parity bands
    of the reference's published row."""
    import json, os
    from new_cg_variants_b200 import experiments as ex
    table = json.load(open(os.path.join(helpers.GOLDEN, "table.json")))
    for name, prec, max_iter in (("bcsstk03", "jacobi", 250), ("nos4", None, 150)):
        A = helpers.load_matrix(name)
        trials = ex.test_matrix(A, max_iter, name, preconditioner=prec, variants=ex.ALL_METHODS, data_dir=str(tmp_path))
        assert set(trials) == {m.__name__ for m in ex.ALL_METHODS}
        saved = np.load(tmp_path / f"{name}_{prec}" / "pr_pcg.npy", allow_pickle=True).item()
        assert saved["name"] == "pr_pcg" and saved["max_iter"] == max_iter
        assert np.array_equal(saved["error_A_norm"], trials["pr_pcg"]["error_A_norm"])
        row, iters, acc = ex.parse_convergence_data(name, prec, ex.TABLE_METHODS, A=A, data_dir=str(tmp_path))
        ref = table[f"{name}_{prec}"]
        assert row.startswith("\\texttt{" + name) and f"& {A.shape[0]} & {A.nnz}" in row
        # (1) the parity rule against the band of the reference's own runs (P2 / P3 of tests/helpers.py) ...
        bands = helpers.cases()[f"{name}_{prec}"]["kstar"]
        for fn, tag in zip(ex.TABLE_METHODS, ("hs", "cg", "m", "pr", "gv", "pipe_pr_m", "pipe_pr")):
            helpers.check_metrics(trials[fn], bands[tag], f"{name}_{prec}/{fn}")
        # (2) ... and the published 2019 row as a coarse regression: iterations within max(2, 3 %),
        # attainable accuracy within 0.6 decades (GV's is chaotic: one decade)
        for k, (i_new, i_ref, a_new, a_ref) in enumerate(zip(iters, ref["iters"], acc, ref["acc"])):
            if i_ref and i_new:
                assert abs(i_new - i_ref) <= max(2, 0.03 * i_ref), (name, ex.TABLE_METHODS[k], i_new, i_ref)
            assert abs(a_new - a_ref) <= (1.0 if ex.TABLE_METHODS[k] == "gv_pcg" else 0.6), (name, ex.TABLE_METHODS[k], a_new, a_ref)


def test_mpi4py_shaped_solvers_single_rank():
    """scaling_experiments_mpi4py signature f(comm, A, b, max_iter) -> (x, times) on the reference's
    model problem (n = 1536, 1500 iterations): final errors within x2 of the reference's own
    (tests/golden/mpi_kat.json, produced by running its files with a one-rank fake mpi4py)."""
    import json, os
    from new_cg_variants_b200 import cg_variants_mpi4py as m
    kat = json.load(open(os.path.join(helpers.GOLDEN, "mpi_kat.json")))
    n, its = kat["pr"]["n"], kat["pr"]["max_iter"]
    A, b = m.model_problem(n)
    comm = m.GpuComm()
    for tag, fn in (("hs", m.hs_cg), ("cg", m.cg_cg), ("gv", m.gv_cg), ("pr", m.pr_cg), ("pipe_pr", m.pipe_pr_cg)):
        x, times = fn(comm, A, b, its)
        err = float(np.linalg.norm(np.ones(n) / np.sqrt(n) - x))
        assert times["tot"] > 0 and x.shape == (n,)
        assert 0.5 <= err / kat[tag]["error"] <= 2.0, (tag, err, kat[tag]["error"])
        # `max_iter` counts updates of x: one more than the numerical-experiment functions
        ref = orc.solve_mpi_style(tag, A.diagonal(), b.copy(), its)
        np.testing.assert_allclose(x, ref, rtol=0, atol=4 * max(err, 1e-12))
    m.clear_sessions()


# ----------------------------------------------------------------- degenerate / ragged inputs
@pytest.mark.parametrize("path", ["stream", "persistent"])
def test_degenerate_and_ragged_inputs(path):
    """n = 1; max_iter = 1 (only the k = 0 entry); a diagonal matrix (the bcsstm* family);
    ragged rows incl. empty ones inside a non-singular matrix; odd n (scalar tail of the
    128-bit vector passes); all against the oracle."""
    import scipy.sparse as sps
    rng = np.random.default_rng(5)
    mats = {
        "n1": sps.csr_matrix(np.array([[3.0]])),
        "diag": sps.diags(rng.uniform(1, 9, 37)).tocsr(),
        "odd_tridiag": sps.diags([-np.ones(100), 2.5 * np.ones(101), -np.ones(100)], [-1, 0, 1]).tocsr(),
    }
    # ragged: arrow matrix (one dense row/column, 300 nnz) + diagonal; row lengths 2..300
    n = 300
    arrow = sps.lil_matrix((n, n))
    arrow.setdiag(np.linspace(400, 800, n))
    arrow[0, :] = 1.0
    arrow[:, 0] = 1.0
    arrow[0, 0] = 1000.0
    mats["arrow"] = sps.csr_matrix(arrow)
    for name, A in mats.items():
        x_true, b, x0 = orc.setup_problem(A)
        dinv = orc.jacobi_dinv(A)
        for max_iter in (1, 2, 9):
            for tag in ALL_TAGS:
                dev = _device_solve(tag, A, b, x0, max_iter, dinv, x_true, path=path)
                ref = orc.solve(tag, A, b, x0, max_iter, dinv=dinv, x_true=x_true)
                for h in orc.HISTORIES:
                    assert dev[h].shape == (max_iter,)
                    # (a diagonal system with Jacobi converges in one step: what follows is
                    # rounding noise, compared against an absolute floor relative to k = 0)
                    np.testing.assert_allclose(dev[h][:3], ref[h][:3], rtol=1e-10, atol=1e-13 * ref[h][0] + 1e-300,
                                               err_msg=f"{name}/{tag}/{h}")
    # a structurally empty row makes A singular: the breakdown is reported, never hidden
    S = sps.csr_matrix(np.diag([1.0, 0.0, 2.0]))
    out = cg_variants.hs_pcg(S, np.array([1.0, 1.0, 1.0]), np.zeros(3), 6, callbacks=[cbk.updated_residual_2_norm],
                             path=path, return_info=True)
    assert out["updated_residual_2_norm"].shape == (6,)


@pytest.mark.parametrize("name", ["bcsstk16", "bcsstk18", "nos7", "bcsstm24", "model_48_8_3", "1138_bus"])
def test_persistent_csr_slab_equals_global_reads(name):
    """Persistent kernel with each CTA's matrix rows resident in shared memory (default when they
    fit) against the same kernel reading the matrix from L2 (csr_slab = 0): same row sums, the
    CTA shape (hence the fixed summation order of the dots) may differ -> equal to rounding."""
    A = helpers.load_matrix(name)
    x_true, b, x0 = orc.setup_problem(A)
    dinv = orc.jacobi_dinv(A)
    res = {}
    for flag in (1, 0):
        with Session(A, dinv=dinv) as s:
            s.set_option("csr_slab", flag)
            for tag in ("hs", "cg", "pr", "pipe_pr", "gv"):
                x, hist, info = s.solve(tag, b, x0, 10, x_true=x_true, path="persistent")
                assert info["path"] == 2
                res[(flag, tag)] = hist
    for tag in ("hs", "cg", "pr", "pipe_pr", "gv"):
        for h in orc.HISTORIES:
            np.testing.assert_allclose(res[(1, tag)][h][:6], res[(0, tag)][h][:6], rtol=1e-9, err_msg=f"{name}/{tag}/{h}")


@pytest.mark.parametrize("shape", [(64, 24, 20), (130, 9, 7), (256, 16, 1)])
def test_cg_cg_with_elided_rt_is_bitwise_identical(shape):
    """CG-CG on the TMA stencil path does not store r~ (= M r, recomputed every iteration in the
    reference, cg_cg.py:132): the stencil pass multiplies it on the fly.  Same products, same
    summation orders -> the solve is BITWISE the one that streams r~ (cg_elide = 0)."""
    nx, ny, nz = shape
    S = PoissonStencil(nx, ny, nz, dim=2 if nz == 1 else 3)
    n = S.shape[0]
    x_true = np.ones(n) / np.sqrt(n)
    b, x0 = S @ x_true, np.zeros(n)
    for dinv in (1 / S.diagonal(), None):
        res = {}
        for flag in (1, 0):
            with Session(S, dinv=dinv) as s:
                s.set_option("cg_elide", flag)
                x, hist, info = s.solve("cg", b, x0, 25, x_true=x_true, path="stream")
                res[flag] = (x, hist, s.vector("rt"), s.vector("w"))
                x, hist, info = s.solve("gv", b, x0, 25, x_true=x_true, path="stream")       # same for GV's w~
                res[("gv", flag)] = (x, hist, s.vector("wt"), s.vector("t"))
        for a, b_ in ((res[1], res[0]), (res[("gv", 1)], res[("gv", 0)])):
            assert np.array_equal(a[0], b_[0])
            for h in orc.HISTORIES:
                assert np.array_equal(a[1][h], b_[1][h], equal_nan=True), h
            assert np.array_equal(a[2], b_[2]) and np.array_equal(a[3], b_[3])
    # a Jacobi VECTOR (non-constant diagonal) keeps the r~ stream
    with Session(S, dinv=1 / (S.diagonal() + np.arange(n) % 3)) as s:
        x, hist, info = s.solve("cg", b, x0, 12, x_true=x_true, path="stream")
        ref = orc.solve("cg", S.tocsr(), b, x0, 12, dinv=1 / (S.diagonal() + np.arange(n) % 3), x_true=x_true)
        np.testing.assert_allclose(hist["updated_residual_2_norm"][:8], ref["updated_residual_2_norm"][:8], rtol=1e-10)


@pytest.mark.parametrize("shape", [(256, 16, 1), (130, 9, 1), (64, 24, 20), (258, 10, 7), (12, 12, 12), (2, 3, 40),
                                   (128, 8, 33), (256, 40, 1)])
def test_pr_fused_single_launch_equals_two_kernel_path(shape):
    """PR-CG / M-CG on the TMA stencil path run ONE kernel per iteration (cgx_stencil_fused.cuh: the
    vector updates and s = A p fused by recomputing p on the tile halo).  Every elementwise and
    row-sum operation is the two-kernel path's, so after one loop trip (same alpha, beta) the state
    vectors are BITWISE those of ew_kernel<EW_PR> + stencil_tma_kernel<SP_PR>; afterwards only the
    summation order of the four fused dots differs -> equal to rounding.  Chunk sizes of 1, 3 and
    8 planes per CTA exercise column changes, outer planes and partial tiles."""
    nx, ny, nz = shape
    S = PoissonStencil(nx, ny, nz, dim=2 if nz == 1 else 3)
    n = S.shape[0]
    x_true = np.ones(n) / np.sqrt(n)
    b, x0 = S @ x_true, np.zeros(n)
    for dinv in (1 / S.diagonal(), None):
        for tag in ("pr", "m"):
            res = {}
            for key, fused, planes in (("two", 0, 8), ("f8", 1, 8), ("f3", 1, 3), ("f1", 1, 1)):
                with Session(S, dinv=dinv) as s:
                    s.set_option("pr_fused", fused)
                    s.set_option("fused_min_planes", planes)
                    s.load_problem(b, x0, x_true)
                    info = s.run(tag, 2, path="stream")
                    one = {v: s.vector(v) for v in ("x", "r", "rt", "p", "s")}
                    x, hist, info = s.solve(tag, b, x0, 14, x_true=x_true, path="stream")
                    res[key] = (one, x, hist, info["kernel_launches"])
            if nx % 2 == 0 and (ny >= 4 or nz == 1):                 # (else: no TMA path, the generic kernels run)
                assert res["f8"][3] < res["two"][3]                  # fewer launches: the fused kernel ran
            for key in ("f8", "f3", "f1"):
                for v in ("x", "r", "rt", "p", "s"):
                    assert np.array_equal(res[key][0][v], res["two"][0][v]), (shape, tag, key, v)
                np.testing.assert_allclose(res[key][1], res["two"][1], rtol=1e-10, atol=1e-13, err_msg=f"{shape}/{tag}/{key}")
                for h in orc.HISTORIES:
                    np.testing.assert_allclose(res[key][2][h], res["two"][2][h], rtol=1e-10, err_msg=f"{shape}/{tag}/{key}/{h}")


def test_banded_model_problem_final_errors_match_the_petsc_run():
    """the reference's PETSc driver problem (ex2b.c:86-97: n = 650 000,
    half-bandwidth 32, kappa = 1e6, rho = 0.95, off-diagonals 1e-4, x* = 1, no preconditioner, 4000
    fixed iterations) on the CSR kernels.  KAT: the final ||x - 1||_2 the reference printed for its
    cg / chcg / pipecg / pipeprcg / pipeprcg_0 solvers (slurm-864568.out:129-205 on 336 ranks, :224-300
    on 280 ranks -- they differ by up to 17 % between the two process counts, i.e. the value is an
    attainable-accuracy level, pinned here within a factor 3)."""
    from new_cg_variants_b200.experiments import BANDED_KAT, banded_model_problem
    A, b, x_true = banded_model_problem()
    assert A.shape[0] == 650000 and A.nnz == 650000 * 65 - 32 * 33
    x0 = np.zeros(A.shape[0])
    got = {}
    with Session(A) as s:
        s.load_problem(b, x0, None)
        for tag in BANDED_KAT:
            info = s.run(tag, 4001, histories=(), path="stream")
            assert info["iterations"] == 4000
            x, _ = s.fetch(want_hist=False)
            got[tag] = float(np.linalg.norm(x - x_true))
    print("banded model problem, ||x - 1||_2 after 4000 iterations (ours | reference 336 ranks, 280 ranks):",
          {t: (f"{got[t]:.3e}", BANDED_KAT[t]) for t in got})
    for tag, ref in BANDED_KAT.items():
        # CG-CG: the reference's PYTHON cg_cg (what this library restates, parity-checked above) is an order of
        # magnitude less accurate than PETSc's KSP chcg on such problems -- its own mpi4py model-problem errors
        # are 1.1e-7 (hs) vs 2.1e-6 (cg), SURVEY.md section 8c -- so that row is only bounded within x10
        f = 10 if tag == "cg" else 3
        lo, hi = min(ref) / f, max(ref) * f
        assert lo <= got[tag] <= hi, (tag, got[tag], ref)
    # the paper's point on this problem: pipe-PR recovers HS-level accuracy, GV / pipe-P lose 2-3 digits
    assert got["pipe_pr"] < 10 * got["hs"] and got["gv"] > 50 * got["hs"] and got["pipe_p"] > 50 * got["hs"]


@pytest.mark.parametrize("cfg", [("3d", 256, ("pr", "pipe_pr", "hs"), 21), ("2d", 4096, ("pr", "pipe_pr"), 21)])
def test_baseline_size_matches_oracle(cfg):
    """BASELINE.json configs[2] / [3] at their OWN size (16.8 M unknowns): the oracle runs on the scipy CSR
    matrix of the same problem (≈1.3 s per iteration with the four callbacks), the device on the
    matrix-free operator -- P1 over all 20 iterations on both residual histories and the A-norm error
    (the window of these problems is far longer: 128^2 already has 188, and it grows with the grid)."""
    kind, grid, tags, K = cfg
    A = orc.poisson3d(grid) if kind == "3d" else orc.poisson2d(grid)
    S = PoissonStencil(grid, grid, grid, dim=3) if kind == "3d" else PoissonStencil(grid, grid, 1, dim=2)
    x_true, b, x0 = orc.setup_problem(A)
    dinv = orc.jacobi_dinv(A)
    assert np.array_equal(S @ x_true, b)
    for tag in tags:
        ref = orc.solve(tag, A, b, x0, K, dinv=dinv, x_true=x_true)
        dev = _device_solve(tag, S, b, x0, K, dinv, x_true, return_info=True)
        assert dev["_info"]["path"] == 1 and dev["_info"]["kernel_launches"] > 0
        worst = 0.0
        for h in ("updated_residual_2_norm", "residual_2_norm", "error_A_norm", "error_2_norm"):
            rel = np.abs(dev[h] - ref[h]) / np.abs(ref[h])
            worst = max(worst, float(rel.max()))
            assert helpers.first_deviation(dev[h], ref[h]) >= K, (cfg, tag, h, rel.max())
        helpers.log_kd(case=f"poisson{kind}_{grid}_jacobi", variant=tag, path="stream", kd=K, window=K, kstar10=None, ensemble=None,
                       max_iter=K, iters=0, acc=float(np.log10(dev["error_A_norm"][-1] / dev["error_A_norm"][0])),
                       iters_band=[0, 0], acc_band=[0.0, 0.0], max_rel=worst)
        print(f"{kind} {grid} {tag}: max relative deviation over {K} history entries = {worst:.2e}")
