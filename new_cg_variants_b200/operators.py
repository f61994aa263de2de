"""Operator objects accepted as ``A`` by the solvers.

The reference passes a scipy CSR matrix (``figure_gen.py:350``); that still works.  For
the synthetic Poisson configurations of BASELINE.json a matrix-free operator is added:
it never materialises the 84-117 M non-zeros, and its host-side ``@`` (used only by the
caller to form ``b = A @ x_true``, figure_gen.py:33) reproduces scipy's CSR product of the
equivalent matrix bit for bit (same term order, separately rounded multiply/add).
"""
from __future__ import annotations

import numpy as np
import scipy.sparse as sps


def canonical_csr(A) -> sps.csr_matrix:
    """Anything scipy can turn into CSR -> canonical fp64/int32 CSR (sorted, no duplicates),
    the form ``csr_matrix(mmread(...))`` has in the reference."""
    if isinstance(A, np.ndarray):
        A = sps.csr_matrix(A)
    elif not sps.issparse(A):
        raise TypeError(f"cannot interpret {type(A).__name__} as a matrix")
    A = A.tocsr()
    if A.dtype != np.float64:
        A = A.astype(np.float64)
    if not A.has_canonical_format:
        A = A.copy()
        A.sum_duplicates()
        A.sort_indices()
    if A.shape[0] != A.shape[1]:
        raise ValueError("A must be square")
    if A.indices.dtype != np.int32 or A.indptr.dtype != np.int32:
        if A.nnz >= 2**31 or A.shape[0] >= 2**31:
            raise ValueError("matrix too large for int32 indices")
        A = sps.csr_matrix((A.data, A.indices.astype(np.int32), A.indptr.astype(np.int32)),
                           shape=A.shape)
    return A


class PoissonStencil:
    """Matrix-free Dirichlet Laplacian, natural ordering ``i = x + nx*(y + ny*z)``.

    dim=2: 5-point (diag 4, off -1) == ``kron(I,T)+kron(T,I)``; dim=3: 7-point (diag 6).
    """

    def __init__(self, nx, ny=None, nz=None, dim=None, diag=None, off=-1.0):
        if dim is None:
            dim = 2 if nz is None else 3
        ny = nx if ny is None else ny
        nz = (nx if dim == 3 else 1) if nz is None else nz
        if dim == 2 and nz != 1:
            raise ValueError("2-D stencil needs nz == 1")
        self.dim, self.nx, self.ny, self.nz = int(dim), int(nx), int(ny), int(nz)
        self.diag = float(2 * dim if diag is None else diag)
        self.off = float(off)
        n = self.nx * self.ny * self.nz
        self.shape = (n, n)
        self.dtype = np.dtype(np.float64)

    # scipy-like surface used by figure_gen-style callers
    def get_shape(self):
        return self.shape

    @property
    def nnz(self):
        nx, ny, nz = self.nx, self.ny, self.nz
        n = nx * ny * nz
        return n + 2 * ((nx - 1) * ny * nz + nx * (ny - 1) * nz + (nx * ny * (nz - 1) if self.dim == 3 else 0))

    def diagonal(self):
        return np.full(self.shape[0], self.diag)

    def matvec(self, v):
        """Host product in canonical-CSR term order (z-1, y-1, x-1, centre, x+1, y+1, z+1)."""
        v = np.asarray(v, dtype=np.float64)
        nx, ny, nz = self.nx, self.ny, self.nz
        g = v.reshape(nz, ny, nx)
        y = np.zeros_like(g)
        if nz > 1:
            y[1:] += self.off * g[:-1]
        if ny > 1:
            y[:, 1:] += self.off * g[:, :-1]
        if nx > 1:
            y[:, :, 1:] += self.off * g[:, :, :-1]
        y += self.diag * g
        if nx > 1:
            y[:, :, :-1] += self.off * g[:, :, 1:]
        if ny > 1:
            y[:, :-1] += self.off * g[:, 1:]
        if nz > 1:
            y[:-1] += self.off * g[1:]
        return y.reshape(-1)

    def __matmul__(self, v):
        return self.matvec(v)

    def tocsr(self):
        t = lambda m: sps.diags([1.0, 0.0, 1.0], [-1, 0, 1], shape=(m, m))
        eye = sps.identity
        nx, ny, nz = self.nx, self.ny, self.nz
        off = (sps.kron(eye(nz), sps.kron(eye(ny), t(nx))) + sps.kron(eye(nz), sps.kron(t(ny), eye(nx))))
        if nz > 1:
            off = off + sps.kron(t(nz), sps.kron(eye(ny), eye(nx)))
        A = sps.csr_matrix(self.off * off + self.diag * eye(nx * ny * nz))
        A.eliminate_zeros()
        A.sort_indices()
        return A

    def __repr__(self):
        return f"PoissonStencil(dim={self.dim}, nx={self.nx}, ny={self.ny}, nz={self.nz}, diag={self.diag}, off={self.off})"


def poisson2d(nx, ny=None):
    return PoissonStencil(nx, ny, 1, dim=2)


def poisson3d(nx, ny=None, nz=None):
    return PoissonStencil(nx, ny, nx if nz is None else nz, dim=3)
