"""GPU parity tests: the CUDA path (through the public Python API -> C ABI) against
  * the committed golden vectors of the unmodified reference,
  * the oracle on seeded inputs at sizes the oracle finishes in seconds,
  * the reference's own test expectations (NumPy / SciPy, np.triu policy).
Structure bit-exact after canonical sorting; values rtol 1e-12 / atol 1e-14 (helpers.py).
"""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
from helpers import assert_csr_equal, assert_dense_equal, golden_expected
from oracle import port
from sparse_matrix_mult_b200 import sparse_matrix_multiply
from sparse_matrix_mult_b200 import synthetic
from sparse_matrix_mult_b200.matrix_ops import last_stats

pytestmark = pytest.mark.gpu


def _ids():
    return [c[0] for c in cases.all_cases()]


@pytest.mark.parametrize("case", cases.all_cases(), ids=_ids())
def test_golden_vectors(golden, case):
    name, a, b, kw = case
    want = golden_expected(golden, name)
    got = sparse_matrix_multiply(a, b, **kw)
    if isinstance(want, np.ndarray):
        assert_dense_equal(got, want, name)
    else:
        assert sp.isspmatrix_csr(got)
        assert got.indices.dtype == np.int32 and got.indptr.dtype == np.int32 and got.data.dtype == np.float64
        assert_csr_equal(got, want, name)
        assert got.has_sorted_indices


# ---- the reference's own tests, restated (tests/test_matrix_multiply.py:89-112, test_edge_case.py:42-71) ----
def test_reference_known_answers():
    m = cases.fixed_matrices()
    cd = np.matmul(m["C"], m["D"])
    r = sparse_matrix_multiply(m["C"], m["D"], output_format='sparse', symmetric=False)
    assert r.shape == cd.shape and np.allclose(r.toarray(), cd)
    r = sparse_matrix_multiply(m["C"], m["D"], output_format='dense', symmetric=False)
    assert r.shape == cd.shape and np.allclose(r, cd)
    cf = np.matmul(m["C"], m["F"])
    r = sparse_matrix_multiply(m["C"], m["F"], output_format='dense', symmetric=True)
    assert np.allclose(np.triu(r), np.triu(cf)) and np.all(np.tril(r, -1) == 0)
    r = sparse_matrix_multiply(m["C"], m["F"], output_format='sparse', symmetric=True)
    assert np.allclose(np.triu(r.toarray()), np.triu(cf))
    r = sparse_matrix_multiply(m["one_a"], m["one_b"], output_format='dense', symmetric=True)
    assert np.allclose(r, [[10]])
    b = np.random.default_rng(0).random((3, 4))
    r = sparse_matrix_multiply(m["zero_rows"], b, output_format='dense')
    assert np.allclose(r, m["zero_rows"] @ b)
    r = sparse_matrix_multiply(sp.csr_matrix(m["zero_rows"]), sp.csr_matrix(b), output_format='sparse')
    assert np.allclose(r.toarray(), m["zero_rows"] @ b)


# ---- seeded random vs oracle, every mode (tests/test_computation_speed.py:37-87 shapes) ---------------------
@pytest.mark.parametrize("n,density", [(500, 0.3), (700, 0.02), (64, 0.9)])
def test_seeded_random_all_modes(n, density):
    a, b = cases.seeded_pair(n, density)
    assert_csr_equal(sparse_matrix_multiply(a, b), port.spgemm_csr(a, b), "sparse")
    assert_csr_equal(sparse_matrix_multiply(a, b, symmetric=True), port.spgemm_csr(a, b, True), "sparse_sym")
    assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense'), port.spgemm_dense(a, b), "dense")
    assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense', symmetric=True),
                       port.spgemm_dense(a, b, True), "dense_sym")
    assert_dense_equal(sparse_matrix_multiply(a, b, use_triple_product=True, compute_full_matrix=0),
                       port.triple_product(a, b, 0), "triple0")
    assert_dense_equal(sparse_matrix_multiply(a, b, use_triple_product=True, compute_full_matrix=1),
                       port.triple_product(a, b, 1), "triple1")


@pytest.mark.parametrize("m,k,n,da,db", [(300, 5000, 280, 0.01, 0.02), (1, 50, 1, 0.5, 0.5), (1000, 30, 30000, 0.2, 0.01),
                                         (4000, 4000, 4000, 0.002, 0.002), (33, 40000, 33, 0.05, 0.05)])
def test_rectangular_vs_oracle(m, k, n, da, db):
    rng = np.random.default_rng(m * 7 + n)
    a = sp.random(m, k, density=da, format='csr', random_state=rng)
    b = sp.random(k, n, density=db, format='csr', random_state=rng)
    assert_csr_equal(sparse_matrix_multiply(a, b), port.spgemm_csr(a, b), "sparse")
    assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense'), port.spgemm_dense(a, b), "dense")
    if m == n:
        assert_csr_equal(sparse_matrix_multiply(a, b, symmetric=True), port.spgemm_csr(a, b, True), "sparse_sym")
        assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense', symmetric=True),
                           port.spgemm_dense(a, b, True), "dense_sym")


# ---- every cost bin of the sparse path: rows from 1 product to dense rows over several windows ------------
def test_sparse_bins_mixed_rows():
    rng = np.random.default_rng(99)
    n = 40_000
    b = sp.random(n, n, density=2e-3, format='csr', random_state=rng)            # ~80 per row
    row_nnz = [0, 1, 2, 5, 20, 60, 200, 700, 3000, 9000]
    rows, cols = [], []
    for i, c in enumerate(row_nnz * 3):
        pick = rng.choice(n, size=c, replace=False)
        rows += [i] * c
        cols += list(pick)
    a = sp.csr_matrix((rng.random(len(rows)), (rows, cols)), shape=(len(row_nnz) * 3, n))
    got = sparse_matrix_multiply(a, b)
    assert_csr_equal(got, port.spgemm_csr(a, b), "mixed bins")
    nnz_rows = np.diff(got.indptr)
    assert nnz_rows.max() > 16384 and nnz_rows.min() == 0        # dense-window and empty rows both present


@pytest.mark.parametrize("n", [600_000, 1_800_000])
def test_sparse_wide_matrices(n):
    """Column counts beyond the pair-table layout: 600,000 columns take the compact rank table (one window),
    1,800,000 take two column windows in both the symbolic bitmap and the numeric rank kernel."""
    rng = np.random.default_rng(n)
    b = sp.random(n, n, density=2.5 / n, format='csr', random_state=rng)
    rows, cols = [], []
    for i, c in enumerate([0, 3, 400, 1500, 5000, 900, 2500]):
        rows += [i] * c
        cols += list(rng.choice(n, size=c, replace=False))
    a = sp.csr_matrix((rng.random(len(rows)), (rows, cols)), shape=(7, n))
    got = sparse_matrix_multiply(a, b)
    want = port.spgemm_csr(a, b)
    assert_csr_equal(got, want, f"wide n={n}")
    assert np.diff(got.indptr).max() > 768          # the block (rank) kernel was exercised
    at = sp.csr_matrix((rng.random(len(rows)), (cols, rows)), shape=(n, 7))      # tall operand, narrow result
    assert_csr_equal(sparse_matrix_multiply(a, at), port.spgemm_csr(a, at), "wide inner dimension")


def test_unsorted_b_with_duplicates_large():
    """B with shuffled columns inside rows and repeated entries: the windowed kernels must fall back to filtering."""
    rng = np.random.default_rng(5)
    n = 30_000
    b0 = sp.random(n, n, density=1e-3, format='csr', random_state=rng)
    idx = b0.indices.copy()
    for r in range(n):
        s, e = b0.indptr[r], b0.indptr[r + 1]
        idx[s:e] = idx[s:e][::-1]
    b = sp.csr_matrix((n, n))
    b.indptr, b.indices, b.data = b0.indptr, idx, b0.data
    a = sp.random(50, n, density=0.05, format='csr', random_state=rng)
    assert_csr_equal(sparse_matrix_multiply(a, b), port.spgemm_csr(a, b), "unsorted B")
    assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense'), port.spgemm_dense(a, b), "unsorted B dense")


def test_cancellation_keeps_explicit_zeros():
    a, b = cases.cancelling()
    got = sparse_matrix_multiply(a, b)
    want = port.spgemm_csr(a, b)
    assert_csr_equal(got, want)
    assert (got.data == 0.0).any()         # structural entry with value exactly 0.0 is kept, like the reference


# ---- reduced BASELINE configs against the oracle -------------------------------------------------------
@pytest.mark.parametrize("name", ["cfg1s", "cfg2s", "cfg3s", "cfg5s", "cfg4r10", "cfg4r12"])
def test_reduced_configs_vs_oracle(name):
    w = synthetic.workload(name)
    got = sparse_matrix_multiply(w["a"], w["b"], **w["kwargs"])
    want = port.sparse_matrix_multiply(w["a"], w["b"], **w["kwargs"])
    if w["kind"] == "sparse":
        assert_csr_equal(got, want, name)
        st = last_stats()
        assert st["products"] == port.count_products(w["a"], w["b"])
        assert st["nnz_c"] == want.nnz
    else:
        assert_dense_equal(got, want, name)


def test_cfg1_full_size_vs_oracle():
    w = synthetic.workload("cfg1")
    got = sparse_matrix_multiply(w["a"], w["b"], **w["kwargs"])
    assert_csr_equal(got, port.spgemm_csr(w["a"], w["b"]), "cfg1")


# ---- mirror extension and reference-full mode ----------------------------------------------------------
def test_mirror_extension():
    a, b = cases.seeded_pair(300, 0.05)
    s = sp.csr_matrix(a @ a.T)
    up = sparse_matrix_multiply(a, sp.csr_matrix(a.T), output_format='dense', symmetric=True)
    full = sparse_matrix_multiply(a, sp.csr_matrix(a.T), output_format='dense', symmetric=True, mirror=True)
    # two separate runs: products are added with atomic reductions, so sums may associate differently
    np.testing.assert_allclose(np.triu(full), np.triu(up), rtol=1e-12, atol=1e-14)
    assert np.array_equal(full, full.T)
    np.testing.assert_allclose(full, s.toarray(), rtol=1e-12, atol=1e-14)
    q = cases.banded(300)
    t_up = sparse_matrix_multiply(a, q, use_triple_product=True)
    t_full = sparse_matrix_multiply(a, q, use_triple_product=True, mirror=True)
    # two separate runs: the scatter-adds are atomic, so sums may associate differently
    np.testing.assert_allclose(np.triu(t_full), np.triu(t_up), rtol=1e-12, atol=1e-14)
    assert np.array_equal(t_full, t_full.T)


# ---- API behaviour of the reference wrapper (matrix_ops.py:288-322) -------------------------------------
def test_api_errors_and_short_circuits():
    a = sp.random(5, 7, density=0.5, format='csr', random_state=1)
    b = sp.random(6, 5, density=0.5, format='csr', random_state=2)
    with pytest.raises(ValueError, match="incompatible"):
        sparse_matrix_multiply(a, b)
    b = sp.random(7, 4, density=0.5, format='csr', random_state=2)
    with pytest.raises(ValueError, match="square"):
        sparse_matrix_multiply(a, b, symmetric=True)
    with pytest.raises(ValueError, match="compute_full_matrix"):
        sparse_matrix_multiply(a, b, compute_full_matrix=2)
    with pytest.raises(ValueError, match="imem_size"):
        sparse_matrix_multiply(a, b, imem_size="lots")
    z = sparse_matrix_multiply(np.zeros((3, 3)), np.zeros((3, 4)))
    assert sp.isspmatrix_csr(z) and z.shape == (3, 4) and z.nnz == 0
    z = sparse_matrix_multiply(np.zeros((3, 3)), np.zeros((3, 4)), output_format='dense')
    assert isinstance(z, np.ndarray) and z.shape == (3, 4) and not z.any()
    # triple product with an empty operand and the default output_format: empty CSR of shape (m, Q.cols)
    z = sparse_matrix_multiply(np.zeros((3, 4)), np.eye(4), use_triple_product=True)
    assert sp.isspmatrix_csr(z) and z.shape == (3, 4)
    # invalid output_format: prints and returns zeros (matrix_ops.py:367-387)
    z = sparse_matrix_multiply(a, b, output_format='banana')
    assert isinstance(z, np.ndarray) and z.shape == (5, 4) and not z.any()
    # use_triple_product wins over output_format (matrix_ops.py:325)
    h = sp.random(6, 9, density=0.4, format='csr', random_state=3)
    q = cases.banded(9, 2)
    t = sparse_matrix_multiply(h, q, output_format='sparse', use_triple_product=True)
    assert isinstance(t, np.ndarray) and t.shape == (6, 6)


def test_full_size_cfg2_properties():
    """BASELINE config 2 at full size: too slow for the oracle in a unit test, so size-independent properties:
    lower triangle exactly zero, diagonal = squared row norms, and a random sample of entries against SciPy."""
    w = synthetic.workload("cfg2")
    a = w["a"]
    c = sparse_matrix_multiply(a, w["b"], **w["kwargs"])
    assert c.shape == (20_000, 20_000)
    rng = np.random.default_rng(0)
    rows = rng.choice(20_000, 64, replace=False)
    diag = np.asarray(a.multiply(a).sum(axis=1)).ravel()
    np.testing.assert_allclose(np.diag(c), diag, rtol=1e-12, atol=1e-14)
    want = (a[rows] @ a.T).toarray()
    for k, r in enumerate(rows):
        np.testing.assert_allclose(c[r, r:], want[k, r:], rtol=1e-12, atol=1e-14)
        assert not c[r, :r].any()
