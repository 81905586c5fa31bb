"""Comparison policy of the parity tests (BASELINE.md section 4 "Parity gates").

structure : indptr / indices bit-exact after canonical (column-sorted) ordering
values    : rtol 1e-12, atol 1e-14 -- float64 sums may be associated differently on the GPU
symmetric : compare np.triu only where the reference leaves the lower triangle zero (and check it IS zero)
"""
import numpy as np
from scipy.sparse import csr_matrix

RTOL, ATOL = 1e-12, 1e-14


def canonical(c):
    """Sorted-column copy of a CSR that keeps explicit zeros and duplicates untouched otherwise."""
    c = csr_matrix(c, copy=True)
    c.sort_indices()
    return c


def assert_csr_equal(got, want, what=""):
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    g, w = canonical(got), canonical(want)
    assert g.nnz == w.nnz, f"{what}: nnz {g.nnz} != {w.nnz}"
    assert np.array_equal(np.asarray(g.indptr, dtype=np.int64), np.asarray(w.indptr, dtype=np.int64)), f"{what}: indptr differs"
    assert np.array_equal(g.indices, w.indices), f"{what}: indices differ"
    np.testing.assert_allclose(g.data, w.data, rtol=RTOL, atol=ATOL, err_msg=f"{what}: values")


def assert_dense_equal(got, want, what=""):
    assert got.shape == want.shape, f"{what}: shape {got.shape} != {want.shape}"
    assert got.dtype == np.float64 and got.flags.c_contiguous, f"{what}: dtype/layout"
    np.testing.assert_allclose(got, want, rtol=RTOL, atol=ATOL, err_msg=what)


def golden_expected(golden, name):
    if name + "/dense" in golden.files:
        return golden[name + "/dense"]
    shape = tuple(int(x) for x in golden[name + "/shape"])
    out = csr_matrix(shape)
    out.indptr, out.indices, out.data = golden[name + "/indptr"], golden[name + "/indices"], golden[name + "/data"]
    return out
