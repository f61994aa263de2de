"""GPU (-m gpu): SURVEY.md section 8 row f4 -- GV residual replacement (gv_cg.py:156-158) and the
callbacks save_x / save_r / lanczos_recurrence / updated_error_A_norm served from device capture
buffers, against the reference's own outputs (tests/golden/f4.npz, make_golden_f4.py) and the
oracle run live."""
import os

import numpy as np
import pytest

import helpers
from helpers import orc
from new_cg_variants_b200 import callbacks as cbk, cg_variants

pytestmark = pytest.mark.gpu
STD = [cbk.error_A_norm, cbk.residual_2_norm, cbk.error_2_norm, cbk.updated_residual_2_norm]
GOLD = np.load(os.path.join(helpers.GOLDEN, "f4.npz"))

PERIODIC = lambda **kw: kw["k"] % 7 == 0                                                  # noqa: E731
DRIFT = lambda **kw: np.linalg.norm(kw["w"] - kw["A"] @ kw["r"]) > 1e-9 * np.linalg.norm(kw["w"])   # noqa: E731


def _nos4():
    A = helpers.load_matrix("nos4")
    x_true, b, x0 = orc.setup_problem(A)
    return A, b, x0, x_true, 1 / A.diagonal()


@pytest.mark.parametrize("name,pred", [("periodic7", PERIODIC), ("drift", DRIFT)])
def test_gv_residual_replacement(name, pred):
    """A k-only predicate becomes a schedule executed on the GPU (one launch sequence, no host
    round trips); a vector-reading one is evaluated on the host every iteration against the device
    state.  Both follow the reference's replaced-w run: 1e-10 over the first 25 iterations, the
    same attainable accuracy (replacement is what repairs GV's accuracy loss)."""
    A, b, x0, x_true, d = _nos4()
    dev = cg_variants.gv_pcg(A, b, x0, 80, w_replace=pred, preconditioner=lambda v: d * v, callbacks=STD,
                             x_true=x_true, return_info=True)
    never = cg_variants.gv_pcg(A, b, x0, 80, preconditioner=lambda v: d * v, callbacks=STD, x_true=x_true)
    live = orc.solve("gv", A, b, x0, 80, dinv=d, x_true=x_true, w_replace=pred)
    for h in orc.HISTORIES:
        gold = GOLD[f"gv_{name}/{h}"]
        assert np.array_equal(live[h], gold)                      # oracle == reference (same BLAS)
        np.testing.assert_allclose(dev[h][:25], gold[:25], rtol=1e-10, err_msg=f"{name}/{h}")
    acc_dev = orc.convergence_metrics(dev["error_A_norm"])[1]
    acc_ref = orc.convergence_metrics(GOLD[f"gv_{name}/error_A_norm"])[1]
    acc_never = orc.convergence_metrics(never["error_A_norm"])[1]
    assert abs(acc_dev - acc_ref) <= np.log10(2) + 0.2, (acc_dev, acc_ref)
    assert not np.array_equal(dev["updated_residual_2_norm"], never["updated_residual_2_norm"])
    print(f"gv {name}: attainable accuracy 1e{acc_dev:.2f} (reference 1e{acc_ref:.2f}; without replacement 1e{acc_never:.2f}), "
          f"{dev['_info']['kernel_launches']} launches")
    if name == "periodic7":
        assert dev["_info"]["kernel_launches"] < 80 * 8           # schedule: no per-iteration host stepping


@pytest.mark.parametrize("fn,tag", [("hs_pcg", "hs"), ("pr_pcg", "pr"), ("pipe_pr_pcg", "pipe_pr")])
def test_capture_served_callbacks(fn, tag):
    """save_x, save_r, lanczos_recurrence, updated_error_A_norm: recorded on the GPU during ONE solve
    (no host round trip per iteration), post-processed by the callback bodies; against the
    reference's outputs on nos4 + Jacobi."""
    A, b, x0, x_true, d = _nos4()
    extra = [cbk.save_x, cbk.save_r, cbk.lanczos_recurrence, cbk.updated_error_A_norm]
    out = getattr(cg_variants, fn)(A, b, x0, 40, preconditioner=lambda v: d * v, callbacks=STD + extra, x_true=x_true,
                                   return_info=True)
    assert out["_info"]["kernel_launches"] < 40 * 8               # one solve, not 40 one-iteration advances
    plain = getattr(cg_variants, fn)(A, b, x0, 40, preconditioner=lambda v: d * v, callbacks=STD, x_true=x_true,
                                     path="stream")               # (capturing runs the stream kernels)
    for h in orc.HISTORIES:
        assert np.array_equal(out[h], plain[h])                   # capturing does not perturb the solve
    g = lambda k: GOLD[f"{fn}/{k}"]                               # noqa: E731
    assert out["x"].shape == (40, 100) and out["r"].shape == (40, 100)
    np.testing.assert_allclose(out["x"][:20], g("x")[:20], rtol=1e-9, atol=1e-14)
    np.testing.assert_allclose(np.linalg.norm(out["r"][:20], axis=1), np.linalg.norm(g("r")[:20], axis=1), rtol=1e-10)
    np.testing.assert_allclose(out["lanczos_alpha"][:20], g("lanczos_alpha")[:20], rtol=1e-9)
    np.testing.assert_allclose(out["lanczos_beta"][:20], g("lanczos_beta")[:20], rtol=1e-9)
    np.testing.assert_allclose(out["lanczos_z"][:, :15], g("lanczos_z")[:, :15], rtol=0, atol=1e-8)
    np.testing.assert_allclose(out["updated_error_A_norm"][:20], g("updated_error_A_norm")[:20], rtol=1e-8)
    # the recurrence-quality curves are rounding-level quantities: same order of magnitude
    for key in ("lanczos_3_term_error", "lanczos_orthogonality"):
        assert out[key].shape == g(key).shape
        assert np.all(np.isfinite(out[key]))
        assert np.median(out[key][:15]) < 50 * np.median(g(key)[:15]) + 1e-300
    # the stepwise protocol (a foreign callable forces it) gives the same capture-served outputs
    seen = []
    slow = getattr(cg_variants, fn)(A, b, x0, 40, preconditioner=lambda v: d * v,
                                    callbacks=STD + extra + [lambda **kw: seen.append(kw["k"])], x_true=x_true,
                                    path="stream")
    assert seen == list(range(40))
    assert np.array_equal(slow["x"], out["x"]) and np.array_equal(slow["lanczos_alpha"], out["lanczos_alpha"])
    assert np.array_equal(slow["lanczos_beta"], out["lanczos_beta"])
