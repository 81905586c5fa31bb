"""Seeded structure fuzz of the CUDA path against the oracle (GPU): many small operands whose STRUCTURE is adversarial --
empty matrices, empty rows and columns, single rows, dimensions on either side of the warp / bin boundaries, dense
blocks next to empty ones, hub columns, banded Q with emptied rows, Q with runs longer than the weight table -- through
all the modes of sparse_matrix_multiply (/root/reference/sparse_matrix_mult/matrix_ops.py:253-397).  Values are
compared to rtol 1e-12, structure bit-exactly after sorting (helpers.py).
"""
import numpy as np
import pytest
import scipy.sparse as sp

from helpers import assert_csr_equal, assert_dense_equal
from oracle import port
from sparse_matrix_mult_b200 import device as dev
from sparse_matrix_mult_b200 import sparse_matrix_multiply

pytestmark = pytest.mark.gpu

_DIMS = [1, 2, 3, 31, 32, 33, 63, 64, 65, 100, 127, 128, 129, 257, 400]
_DENS = [0.0, 0.002, 0.02, 0.2, 0.9]


def _shape_up(x, rng):
    """Knock structure into a random matrix: empty rows, empty columns, one dense row, one hub column."""
    x = x.tocoo()
    m, n = x.shape
    row, col, val = x.row, x.col, x.data
    what = rng.integers(0, 6)
    if what == 0 and m > 2:                               # a third of the rows empty
        keep = ~np.isin(row, rng.choice(m, size=m // 3, replace=False))
        row, col, val = row[keep], col[keep], val[keep]
    elif what == 1 and n > 2:                             # a third of the columns empty
        keep = ~np.isin(col, rng.choice(n, size=n // 3, replace=False))
        row, col, val = row[keep], col[keep], val[keep]
    elif what == 2:                                       # one dense row
        r = int(rng.integers(0, m))
        keep = row != r
        row = np.concatenate([row[keep], np.full(n, r)])
        col = np.concatenate([col[keep], np.arange(n)])
        val = np.concatenate([val[keep], rng.random(n) + 0.5])
    elif what == 3:                                       # hub column
        c = int(rng.integers(0, n))
        keep = col != c
        row = np.concatenate([row[keep], np.arange(m)])
        col = np.concatenate([col[keep], np.full(m, c)])
        val = np.concatenate([val[keep], rng.random(m) + 0.5])
    elif what == 4 and m > 1:                             # the last rows empty (trailing empty rows in indptr)
        keep = row < m - max(1, m // 4)
        row, col, val = row[keep], col[keep], val[keep]
    x = sp.csr_matrix((val, (row, col)), shape=(m, n), dtype=np.float64)
    x.eliminate_zeros()
    x.sort_indices()
    x.indices = x.indices.astype(np.int32)
    x.indptr = x.indptr.astype(np.int32)
    return x


def _random(m, n, rng):
    x = sp.random(m, n, density=float(rng.choice(_DENS)), format='csr', random_state=rng, dtype=np.float64)
    return _shape_up(x, rng)


def _dims(rng, count):
    return [int(rng.choice(_DIMS)) for _ in range(count)]


@pytest.mark.parametrize("seed", range(40))
def test_fuzz_products(seed):
    rng = np.random.default_rng(1000 + seed)
    m, k, n = _dims(rng, 3)
    if seed % 2:
        n = m                                             # square result: the symmetric variants apply
    a, b = _random(m, k, rng), _random(k, n, rng)
    assert_csr_equal(sparse_matrix_multiply(a, b), port.spgemm_csr(a, b), f"sparse {m}x{k}x{n}")
    assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense'), port.spgemm_dense(a, b), f"dense {m}x{k}x{n}")
    if m == n:
        assert_csr_equal(sparse_matrix_multiply(a, b, symmetric=True), port.spgemm_csr(a, b, True), "sparse_sym")
        assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense', symmetric=True),
                           port.spgemm_dense(a, b, True), "dense_sym")
    # the kernels themselves through the device API (no host short-circuit for operands without entries)
    A, B = dev.DeviceMatrix.from_scipy(a), dev.DeviceMatrix.from_scipy(b)
    ref = (a @ b).toarray()
    res = dev.spgemm_csr(A, B, False, 0, m)
    np.testing.assert_allclose(res.to_scipy().toarray(), ref, rtol=1e-12, atol=1e-13)
    res.free()
    out = dev.spgemm_dense(A, B, False, 0, m)
    np.testing.assert_allclose(out.to_host(), ref, rtol=1e-12, atol=1e-13)
    out.free()
    A.free()
    B.free()


def _q_for(k, rng):
    kind = int(rng.integers(0, 5))
    if kind == 0:                                         # banded, possibly wider than the weight table (96)
        hw = int(rng.choice([0, 1, 5, 40, 110]))
        hw = min(hw, max(0, k - 1))
        offs = list(range(-hw, hw + 1))
        q = sp.diags([np.full(k - abs(o), np.exp(-abs(o) / 9.0)) for o in offs], offs, format='csr', dtype=np.float64)
    elif kind == 1:                                       # banded with emptied rows
        hw = min(int(rng.choice([1, 3, 20])), max(0, k - 1))
        offs = list(range(-hw, hw + 1))
        q = sp.diags([np.full(k - abs(o), 1.0 / (1 + abs(o))) for o in offs], offs, format='lil', dtype=np.float64)
        for r in rng.choice(k, size=max(1, k // 4), replace=False):
            q[int(r), :] = 0
        q = q.tocsr()
    elif kind == 2:                                       # general sparse
        q = sp.random(k, k, density=float(rng.choice([0.002, 0.05, 0.5])), format='csr', random_state=rng)
    elif kind == 3:                                       # empty
        q = sp.csr_matrix((k, k), dtype=np.float64)
    else:                                                 # one run per row, at random places and of random lengths
        rows, cols = [], []
        for r in range(k):
            ln = int(rng.integers(0, min(k, 130) + 1))
            c0 = int(rng.integers(0, k - ln + 1))
            rows += [r] * ln
            cols += list(range(c0, c0 + ln))
        q = sp.csr_matrix((rng.random(len(rows)) + 0.1, (rows, cols)), shape=(k, k))
    q.eliminate_zeros()
    q.sort_indices()
    q.indices = q.indices.astype(np.int32)
    q.indptr = q.indptr.astype(np.int32)
    return q


@pytest.mark.parametrize("seed", range(120))
def test_fuzz_triple_product(seed, monkeypatch):
    rng = np.random.default_rng(5000 + seed)
    if seed >= 50:                                        # more column panels than needed / the general kernel forced
        if seed % 3:
            monkeypatch.setenv("SPGEMM_B200_TRIPLE_PANELS", str(1 + seed % 5))
        if seed % 4 == 0:
            monkeypatch.setenv("SPGEMM_B200_TRIPLE_GENERIC", "1")
    n, k = _dims(rng, 2)
    if seed % 5 == 0:
        k = int(rng.choice([700, 1500, 3000]))            # long H rows: more entries than warps, ranges of many steps
    h = _random(n, k, rng)
    q = _q_for(k, rng)
    what = f"triple n={n} k={k} nnz(H)={h.nnz} nnz(Q)={q.nnz}"
    # the entry point, against the oracle's restatement of it: an operand without entries short-circuits to an empty
    # (n x k) result whatever the mode (matrix_ops.py:315-319), which the drop-in reproduces
    for kwargs in (dict(), dict(compute_full_matrix=1)):
        got = sparse_matrix_multiply(h, q, use_triple_product=True, **kwargs)
        want = port.sparse_matrix_multiply(h, q, use_triple_product=True, **kwargs)
        if sp.issparse(want):
            assert sp.issparse(got)
            assert_csr_equal(got, want, what)
        else:
            assert_dense_equal(got, want, what + str(kwargs))
    # the kernels themselves, empty operands included: row ranges through the device API, full (unfiltered) product
    # and the upper triangle
    H, Q = dev.DeviceMatrix.from_scipy(h), dev.DeviceMatrix.from_scipy(q)
    full = (h @ q @ h.T).toarray()
    cut = n // 2
    for r0, r1 in ((0, cut), (cut, n)):
        if r1 > r0:
            out = dev.triple_product(H, Q, None, False, r0, r1)
            np.testing.assert_allclose(out.to_host(), full[r0:r1], rtol=1e-12, atol=1e-13, err_msg=what)
            out.free()
            out = dev.triple_product(H, Q, None, True, r0, r1)
            np.testing.assert_allclose(out.to_host(), np.triu(full)[r0:r1], rtol=1e-12, atol=1e-13, err_msg=what + " upper")
            out.free()
    H.free()
    Q.free()


def _shuffle_rows(x, rng):
    """The same matrix with the entries of every row in random order (unsorted CSR)."""
    x = x.copy()
    for r in range(x.shape[0]):
        s, e = x.indptr[r], x.indptr[r + 1]
        perm = rng.permutation(e - s)
        x.indices[s:e] = x.indices[s:e][perm]
        x.data[s:e] = x.data[s:e][perm]
    x.has_sorted_indices = False
    return x


@pytest.mark.parametrize("seed", range(30))
def test_fuzz_products_heavier_rows(seed):
    """Rows that land in the warp and block bins of the sparse path (hundreds to tens of thousands of products, hub
    rows of B, a dense row of A), the symmetric cut through them, and unsorted operands (every third seed: B, or both,
    with the entries of each row shuffled -- canonicalised on the device, same result as for the sorted operands)."""
    rng = np.random.default_rng(9000 + seed)
    m = int(rng.choice([300, 800, 2000]))
    k = int(rng.choice([300, 1000, 4000]))
    n = m if seed % 2 else int(rng.choice([500, 3000, 70000]))
    da = float(rng.choice([0.003, 0.03, 0.2]))
    db = float(rng.choice([0.003, 0.03, 0.2])) if n < 10000 else 0.002
    a = _shape_up(sp.random(m, k, density=da, format='csr', random_state=rng), rng)
    b = _shape_up(sp.random(k, n, density=db, format='csr', random_state=rng), rng)
    a_in, b_in = a, b
    if seed % 3 == 0:
        b_in = _shuffle_rows(b, rng)
        if seed % 6 == 0:
            a_in = _shuffle_rows(a, rng)
    what = f"{m}x{k}x{n} nnz {a.nnz} {b.nnz}"
    assert_csr_equal(sparse_matrix_multiply(a_in, b_in), port.spgemm_csr(a, b), "sparse " + what)
    if m == n:
        assert_csr_equal(sparse_matrix_multiply(a_in, b_in, symmetric=True), port.spgemm_csr(a, b, True), "sparse_sym " + what)
    if m * n <= 4_000_000:
        assert_dense_equal(sparse_matrix_multiply(a_in, b_in, output_format='dense'), port.spgemm_dense(a, b), "dense " + what)
        if m == n:
            assert_dense_equal(sparse_matrix_multiply(a_in, b_in, output_format='dense', symmetric=True),
                               port.spgemm_dense(a, b, True), "dense_sym " + what)


# ---- the same through the multi-GPU driver (csrc/multi.cu): every available GPU count; on a one-GPU box n_gpus=1 is
# ---- forced through it.  Rows < GPUs, operands without entries in some row blocks, one-row matrices.
def _gpu_counts():
    from sparse_matrix_mult_b200.matrix_ops import matrix_ops
    n = matrix_ops.get_lib().spgemm_b200_device_count()
    return [c for c in (1, 2, 3, 4, 8) if c <= max(1, n)]


@pytest.mark.parametrize("seed", range(24))
def test_fuzz_multi_gpu(seed, monkeypatch):
    monkeypatch.setenv("SPGEMM_B200_FORCE_MULTI", "1")
    rng = np.random.default_rng(7000 + seed)
    m, k = _dims(rng, 2)
    a = _random(m, k, rng)
    b = _random(k, m, rng)
    q = _q_for(k, rng)
    for n_gpus in _gpu_counts():
        what = f"n_gpus={n_gpus} {m}x{k} nnz {a.nnz} {b.nnz} {q.nnz}"
        assert_csr_equal(sparse_matrix_multiply(a, b, n_gpus=n_gpus), port.sparse_matrix_multiply(a, b), "sparse " + what)
        assert_csr_equal(sparse_matrix_multiply(a, b, symmetric=True, n_gpus=n_gpus),
                         port.sparse_matrix_multiply(a, b, symmetric=True), "sparse_sym " + what)
        assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense', n_gpus=n_gpus),
                           port.sparse_matrix_multiply(a, b, output_format='dense'), "dense " + what)
        assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense', symmetric=True, n_gpus=n_gpus),
                           port.sparse_matrix_multiply(a, b, output_format='dense', symmetric=True), "dense_sym " + what)
        got = sparse_matrix_multiply(a, q, use_triple_product=True, n_gpus=n_gpus)
        want = port.sparse_matrix_multiply(a, q, use_triple_product=True)
        if sp.issparse(want):
            assert sp.issparse(got)
            assert_csr_equal(got, want, "triple " + what)
        else:
            assert_dense_equal(got, want, "triple " + what)
