"""CPU: host-side logic of the drop-in boundary (no compute on the GPU)."""
import numpy as np
import pytest
import scipy.sparse as sps

import helpers
from helpers import orc
from new_cg_variants_b200 import PoissonStencil, canonical_csr, callbacks, cg_variants
from new_cg_variants_b200.cg_variants import probe_preconditioner, _split_callbacks


def test_exports_match_reference_names():
    # numerical_experiments/cg_variants/__init__.py:64-74
    for stem in ("hs", "cg", "gv", "pr", "m", "pipe_p", "pipe_pr", "pipe_p_m", "pipe_pr_m"):
        for suf in ("_cg", "_pcg"):
            f = getattr(cg_variants, stem + suf)
            assert f.__name__ == stem + suf


def test_probe_identity_and_jacobi():
    A = helpers.load_matrix("bcsstk03")
    n = A.shape[0]
    assert probe_preconditioner(lambda x: x, n) is None
    assert probe_preconditioner(None, n) is None
    d = probe_preconditioner(lambda x: (1 / A.diagonal()) * x, n)       # figure_gen.py:43
    assert np.array_equal(d, 1 / A.diagonal())


def test_probe_rejects_non_diagonal():
    A = helpers.load_matrix("nos4")
    n = A.shape[0]
    with pytest.raises(NotImplementedError):
        probe_preconditioner(lambda x: A @ x, n)
    with pytest.raises(NotImplementedError):
        probe_preconditioner(lambda x: x[:-1], n)


def test_callback_classification():
    from new_cg_variants_b200.callbacks import (error_A_norm, residual_2_norm, error_2_norm,
                                                updated_residual_2_norm, print_k, save_x)
    dev, ticks, generic = _split_callbacks([error_A_norm, residual_2_norm, error_2_norm,
                                            updated_residual_2_norm, print_k(10), save_x])
    assert dev == ["error_A_norm", "residual_2_norm", "error_2_norm", "updated_residual_2_norm"]
    assert len(ticks) == 1 and generic == [save_x]

    def error_A_norm_like(**kw):      # a user callable that merely shares a reference name is CALLED,
        pass                          # not replaced by the device history
    error_A_norm_like.__name__ = "error_A_norm"
    dev, ticks, generic = _split_callbacks([error_A_norm_like])
    assert dev == [] and generic == [error_A_norm_like]
    # the reference's own callback objects (module `callbacks.error_A_norm`) do map
    import types
    ref_cb = types.FunctionType(error_A_norm_like.__code__, {}, "error_A_norm")
    ref_cb.__module__ = "callbacks.error_A_norm"
    assert _split_callbacks([ref_cb])[0] == ["error_A_norm"]


def test_host_callbacks_follow_reference_protocol():
    """The callback bodies, driven by the oracle's iterates, give the oracle's histories."""
    A = helpers.load_matrix("nos4")
    x_true, b, x0 = orc.setup_problem(A)
    out = orc.solve("hs", A, b, x0, 3, x_true=x_true, return_state=True)
    output = {"name": "hs_pcg"}
    extra = {"x_true": x_true}
    st = out["_state"]
    # state after the last iteration == history index 2; emulate k=0 allocation first
    for k, (x, r) in [(0, (x0, b - A @ x0)), (2, (st["x"], st["r"]))]:
        for cb in (callbacks.error_A_norm, callbacks.residual_2_norm, callbacks.error_2_norm,
                   callbacks.updated_residual_2_norm):
            cb(output=output, A=A, b=b, x_k=x, r_k=r, k=k, max_iter=3, kwargs=extra)
    for h in orc.HISTORIES:
        assert output[h][0] == out[h][0] and output[h][2] == out[h][2], h


@pytest.mark.parametrize("shape", [(16, 16, 1), (7, 5, 1), (6, 5, 4), (12, 12, 12), (1, 9, 3)])
def test_stencil_host_matvec_is_bitwise_csr(shape):
    nx, ny, nz = shape
    S = PoissonStencil(nx, ny, nz, dim=2 if nz == 1 else 3)
    A = S.tocsr()
    ref = orc.poisson2d(nx, ny) if nz == 1 else orc.poisson3d(nx, ny, nz)
    assert (A != ref).nnz == 0 and A.nnz == S.nnz
    v = np.random.default_rng(1).standard_normal(nx * ny * nz)
    assert np.array_equal(S @ v, A @ v)
    assert np.array_equal(S.diagonal(), A.diagonal())


def test_canonical_csr():
    A = sps.coo_matrix(([1.0, 2.0, 3.0, 4.0], ([0, 0, 1, 1], [1, 1, 0, 1])), shape=(2, 2))
    C = canonical_csr(A)
    assert C.has_canonical_format and C.indices.dtype == np.int32 and C.nnz == 3
    assert canonical_csr(np.eye(3)).nnz == 3
    with pytest.raises(ValueError):
        canonical_csr(sps.csr_matrix(np.ones((2, 3))))


def test_gv_w_replace_is_planned_not_rejected():
    """gv_cg.py:156-158: a predicate that only looks at k becomes a device schedule, one that reads
    vectors is evaluated on the host every iteration ("host"), the reference's default never fires."""
    from new_cg_variants_b200.cg_variants import _plan_w_replace, _never
    assert _plan_w_replace(_never, 50) is None
    assert _plan_w_replace(lambda **kw: False, 50) is None
    sched = _plan_w_replace(lambda **kw: kw["k"] % 10 == 0, 35)
    assert sched.dtype == np.uint8 and list(np.nonzero(sched)[0]) == [10, 20, 30]
    assert _plan_w_replace(lambda **kw: np.linalg.norm(kw["w"]) > 1.0, 20) == "host"
    # stateful predicates get the reference's scratch dict, in iteration order
    def every_third(**kw):
        f = kw["wk_replace_flags"]
        f["n"] = f.get("n", 0) + 1
        return f["n"] % 3 == 0
    assert list(np.nonzero(_plan_w_replace(every_third, 10))[0]) == [3, 6, 9]


def test_convergence_metrics_definition():
    e = np.array([1.0, 1e-2, 1e-6, 1e-8, 1e-7])
    assert orc.convergence_metrics(e) == (2, -8.0)          # figure_gen.py:80-89
    assert orc.convergence_metrics(np.array([1.0, 0.5]))[0] == 0
