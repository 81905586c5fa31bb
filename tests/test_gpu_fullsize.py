"""Full-size parity of the BASELINE.json configs against the reference's own binaries (oracle/_ref, built by
`make -C oracle ref`; the oracle port stands in when they are absent), VERDICT r1 "weak" #1-#3:

  cfg2  full 20,000^2 result against reference dense_sym on ALL of np.triu (row blocks, to bound memory)
  cfg3  full 5,000^2 triple product against reference triple_product
  cfg4r full R-MAT scale 16 (nnz(C) = 1.64e8) against the shipped serial sparse_nosym
  cfg5  the full 40,000^2 triple product on the GPU; its leading R x R block is the independent problem
        H[:R] Q H[:R]^T, which the reference computes in seconds
  cfg4  cannot be compared (nnz(C) > 2^31 does not fit the reference's int CSR): device-resident run checked
        against SURVEY.md's exact nnz(C) -- skipped unless SPGEMM_TEST_CFG4=1 (needs ~120 GB of HBM, ~1 min of host setup)

Structure bit-exact after canonical sorting; values rtol 1e-12 / atol 1e-14 (helpers.py); symmetric modes on
np.triu with the strictly lower triangle required to be exactly zero.
"""
import os

import numpy as np
import pytest
import scipy.sparse as sp

from helpers import ATOL, RTOL, assert_csr_equal
from oracle import port, ref
from sparse_matrix_mult_b200 import sparse_matrix_multiply, synthetic
from sparse_matrix_mult_b200.matrix_ops import last_stats

pytestmark = pytest.mark.gpu


def _ref_dense(a, b, sym):
    return ref.omp().dense(a, b, sym) if ref.available() else port.spgemm_dense(a, b, sym)


def _ref_triple(h, q):
    return ref.omp().triple(h, q, 0) if ref.available() else port.triple_product(h, q, 0)


def _ref_sparse(a, b, sym):
    return ref.shipped().sparse(a, b, sym) if ref.available() else port.spgemm_csr(a, b, sym)


def _assert_upper_equal(got, want, what, block=2000):
    """np.triu(got) == np.triu(want) to tolerance and tril(got, -1) == 0, in row blocks."""
    n = got.shape[0]
    assert got.shape == want.shape
    for r0 in range(0, n, block):
        r1 = min(n, r0 + block)
        g, w = got[r0:r1], want[r0:r1]
        cols = np.arange(n)[None, :]
        rows = np.arange(r0, r1)[:, None]
        upper = cols >= rows
        np.testing.assert_allclose(np.where(upper, g, 0.0), np.where(upper, w, 0.0), rtol=RTOL, atol=ATOL,
                                   err_msg=f"{what}: rows [{r0},{r1})")
        assert not np.where(upper, 0.0, g).any(), f"{what}: non-zero below the diagonal in rows [{r0},{r1})"


def test_cfg2_full_size_vs_reference():
    w = synthetic.workload("cfg2")
    got = sparse_matrix_multiply(w["a"], w["b"], **w["kwargs"])
    want = _ref_dense(w["a"], w["b"], True)
    _assert_upper_equal(got, want, "cfg2")
    assert last_stats()["bytes_d2h"] < got.nbytes * 0.51        # only the upper trapezoids crossed PCIe


def test_cfg3_full_size_vs_reference():
    w = synthetic.workload("cfg3")
    got = sparse_matrix_multiply(w["a"], w["b"], **w["kwargs"])
    want = _ref_triple(w["a"], w["b"])
    _assert_upper_equal(got, want, "cfg3")


def test_cfg4r_full_size_vs_reference():
    w = synthetic.workload("cfg4r")
    got = sparse_matrix_multiply(w["a"], w["b"], **w["kwargs"])
    st = last_stats()
    want = _ref_sparse(w["a"], w["b"], False)
    assert st["nnz_c"] == want.nnz == got.nnz
    assert st["products"] == port.count_products(w["a"], w["b"])
    assert_csr_equal(got, want, "cfg4r")


def test_cfg5_full_size_leading_block_vs_reference():
    w = synthetic.workload("cfg5")
    h, q = w["a"], w["b"]
    got = sparse_matrix_multiply(h, q, **w["kwargs"])
    n = h.shape[0]
    assert got.shape == (n, n)
    # the reference keeps (threads + 1) private R x R copies: bound R by ~6 GB of host memory
    cores = os.cpu_count() or 1
    r = int(min(6000, np.sqrt(6e9 / (8.0 * (cores + 1)))))
    want = _ref_triple(h[:r], q)
    _assert_upper_equal(np.ascontiguousarray(got[:r, :r]), want, f"cfg5 leading {r}x{r} block")
    # the rest of the matrix through size-independent properties: zeros below the diagonal, the diagonal
    # itself (h_i Q h_i^T from SciPy), and a sample of complete rows against SciPy
    hq = (h @ q).tocsr()
    diag = np.asarray(hq.multiply(h).sum(axis=1)).ravel()
    np.testing.assert_allclose(np.diag(got), diag, rtol=1e-11, atol=1e-13)
    rng = np.random.default_rng(5)
    rows = np.sort(rng.choice(n, 24, replace=False))
    want_rows = (hq[rows] @ h.T).toarray()
    for k, i in enumerate(rows):
        np.testing.assert_allclose(got[i, i:], want_rows[k, i:], rtol=1e-11, atol=1e-13, err_msg=f"cfg5 row {i}")
        assert not got[i, :i].any()
    for r0 in range(0, n, 4000):                                 # strictly lower triangle is exactly zero
        blk = got[r0:r0 + 4000]
        assert not np.tril(blk, r0 - 1).any()


@pytest.mark.skipif(os.environ.get("SPGEMM_TEST_CFG4") != "1", reason="needs ~120 GB of HBM; set SPGEMM_TEST_CFG4=1")
def test_cfg4_device_resident_nnz():
    from sparse_matrix_mult_b200 import device as dev
    a = synthetic.rmat(20)
    A = dev.DeviceMatrix.from_scipy(a)
    res = dev.spgemm_csr(A, A)
    assert res.nnz == 9_707_207_800                              # exact chunked count
    ptr = res.indptr_host()
    assert ptr[0] == 0 and ptr[-1] == res.nnz and (np.diff(ptr) >= 0).all()
    res.free()
    A.free()
