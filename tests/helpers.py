"""Shared test utilities: fixture loading and THE parity rule.

Parity rule (BASELINE.json north_star + SURVEY.md section 8c, made precise in DESIGN.md
section "Parity"):

  P1  for k < window: |dev - ref| / |ref| <= 1e-10 on updated_residual_2_norm and
      residual_2_norm, where ``window`` (tests/golden/cases.json) = min(kstar11, ensemble11):
      the iterations over which the reference has not yet left exact_pcg by 1e-11 (north_star:
      "until the reference curve departs from exact arithmetic") and still agrees to 1e-11 with
      itself under five other inner-product summation orders (oracle.DOT_ORDERS).  The device's
      1e-10 thus has one decade of margin over the reference's own sensitivity to rounding
      order; with both thresholds at 1e-10 (kstar10, ensemble -- also stored) the window would
      end where a further summation order has an even chance of having crossed the line: the
      measured first-deviation index kd of every device run is logged next to all four numbers
      (gpurun_out/parity_kd.jsonl -> profiles/parity_r02.md);
  P2  attainable accuracy: log10(min_k rel. A-norm error) not worse than log10(2) above the band the
      reference itself spans under those summation orders, widened by the band's width;
  P3  iterations to rel. A-norm error <= 1e-5 within max(1, 1 %) of that band, widened by
      the band's width.
"""
from __future__ import annotations

import json
import math
import os
import sys

import numpy as np
import scipy.sparse as sps

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
GOLDEN = os.path.join(HERE, "golden")
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from oracle import cg_oracle as orc   # noqa: E402  (tests may use the oracle)

RTOL = 1e-10
RESIDUAL_HISTS = ("updated_residual_2_norm", "residual_2_norm")

_cases = None
_hist = None


def cases():
    global _cases
    if _cases is None:
        _cases = json.load(open(os.path.join(GOLDEN, "cases.json")))
    return _cases


def tier(case):
    return cases()[case].get("tier", "full")


def cases_of(*tiers):
    return [c for c in cases() if tier(c) in tiers]


def log_kd(**rec):
    """Append one record of the measured-parity table (profiles/parity_r02.md is built from it)."""
    d = os.path.join(ROOT, "gpurun_out")
    try:
        os.makedirs(d, exist_ok=True)
        with open(os.path.join(d, "parity_kd.jsonl"), "a") as fh:
            fh.write(json.dumps(rec) + "\n")
    except OSError:
        pass


def golden_history(case, tag, name):
    global _hist
    if _hist is None:
        _hist = np.load(os.path.join(GOLDEN, "histories.npz"))
    return _hist[f"{case}/{tag}/{name}"]


def matrix_names():
    return sorted(f[:-4] for f in os.listdir(os.path.join(GOLDEN, "matrices")) if f.endswith(".npz"))


def load_matrix(name):
    """The CSR matrix ``csr_matrix(mmread(matrices/<name>.mtx))`` of the reference."""
    z = np.load(os.path.join(GOLDEN, "matrices", name + ".npz"))
    n = int(z["n"])
    row, col, val = z["row"], z["col"], z["val"]
    off = row != col
    A = sps.coo_matrix((np.concatenate([val, val[off]]),
                        (np.concatenate([row, col[off]]), np.concatenate([col, row[off]]))),
                       shape=(n, n)).tocsr()
    A.sort_indices()
    return A


def case_matrix(case):
    src = cases()[case]["matrix"]
    if isinstance(src, list):
        return getattr(orc, src[0])(*src[1:])
    return load_matrix(src)


def case_problem(case):
    meta = cases()[case]
    A = case_matrix(case)
    x_true, b, x0 = orc.setup_problem(A)
    dinv = orc.jacobi_dinv(A) if meta["preconditioner"] == "jacobi" else None
    return A, b, x0, x_true, dinv, meta["max_iter"]


def first_deviation(a, ref, tol=RTOL):
    with np.errstate(divide="ignore", invalid="ignore"):
        rel = np.abs(np.asarray(a) - ref) / np.abs(ref)
    bad = np.nonzero(~(rel <= tol))[0]
    return int(bad[0]) if len(bad) else int(len(ref))


def check_window(dev, ref, band, label=""):
    """P1 only (ref may be a stored prefix of the reference's histories)."""
    w = band["window"]
    for h in RESIDUAL_HISTS:
        m = min(w, len(ref[h]))
        kd = first_deviation(dev[h][:m], ref[h][:m])
        assert kd >= m, (f"{label} {h}: deviates from the reference by more than {RTOL} at k={kd} "
                         f"(< window {w}); dev={dev[h][kd]!r} ref={ref[h][kd]!r}")


def check_parity(dev, ref, band, label=""):
    """Assert P1-P3 for one (case, variant).  dev/ref: dicts of history arrays."""
    check_window(dev, ref, band, label)
    return check_metrics(dev, band, label)


def check_metrics(dev, band, label=""):
    """P2 / P3 against the band the reference spans under re-ordered inner products."""
    it, acc = orc.convergence_metrics(dev["error_A_norm"])
    lo, hi = band["acc_band"]
    # an error at or below machine precision (incl. exactly 0: diagonal matrices terminate exactly) is
    # "converged to rounding", whatever its last digits
    floor = math.log10(2.0 ** -52)
    acc, lo, hi = max(acc, floor), max(lo, floor), max(hi, floor)
    width = hi - lo
    # Worse than the band by more than x2 (+ the band's own width) fails.  BETTER than the band is allowed
    # ("higher-accuracy device dots are allowed and help"): min_k of a quantity that
    # fluctuates around the rounding floor dips below the five-order band for some summation orders
    # (nos1_jacobi / m, n = 237: 1e-13.58 against [-12.96, -12.83] with csr_bulk_kernel's order, 1e-12.93 with
    # csr_stream_kernel's, 1e-13.07 on the persistent path) -- but more than a decade below it would mean the
    # error is not being measured.
    assert lo - width - 1.0 <= acc <= hi + width + math.log10(2), \
        f"{label}: attainable accuracy 1e{acc:.2f} outside reference band [{lo:.2f}, {hi:.2f}]"
    ilo, ihi = band["iters_band"]
    iw = ihi - ilo
    slack = lambda v: max(1, math.ceil(0.01 * v))
    if ilo == 0 or ihi == 0:      # "never reached" in (part of) the band: nothing to bound
        return it, acc
    assert ilo - iw - slack(ilo) <= it <= ihi + iw + slack(ihi), \
        f"{label}: {it} iterations to 1e-5, reference band [{ilo}, {ihi}]"
    return it, acc
