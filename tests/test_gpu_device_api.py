"""GPU tests of the entry points beside the host-buffer API:
  * the device-resident row-range calls (spgemm_b200_*_dev) that bench.py's `value` and the multi-GPU paths time:
    disjoint row ranges stitched and compared with the oracle (global row numbers in the col >= row cut);
  * upper_only through the block-bin (bitmap / rank) kernels;
  * n_gpus= through the public API (single-process multi-GPU driver; runs with n_gpus=1 on a one-GPU box and
    with every available count up to 8 otherwise);
  * thread safety of concurrent calls, operand validation, device-side canonicalisation of an unsorted B.
"""
import threading

import numpy as np
import pytest
import scipy.sparse as sp

import cases
from helpers import assert_csr_equal, assert_dense_equal
from oracle import port
from sparse_matrix_mult_b200 import device as dev
from sparse_matrix_mult_b200 import sparse_matrix_multiply, synthetic
from sparse_matrix_mult_b200.matrix_ops import matrix_ops, multi_last_bounds, multi_last_stats

pytestmark = pytest.mark.gpu


def _gpus():
    return matrix_ops.get_lib().spgemm_b200_device_count()


def _ranges(n):
    cuts = [0, n // 5, n // 5, (2 * n) // 3, n]          # includes an empty range
    return [(cuts[i], cuts[i + 1]) for i in range(len(cuts) - 1)]


# ---- device API, row ranges ------------------------------------------------------------------------------
@pytest.mark.parametrize("upper", [False, True])
def test_device_csr_row_ranges(upper):
    a, b = cases.seeded_pair(700, 0.02)
    A, B = dev.DeviceMatrix.from_scipy(a), dev.DeviceMatrix.from_scipy(b)
    want = port.spgemm_csr(a, b, upper)
    blocks = []
    for r0, r1 in _ranges(a.shape[0]):
        res = dev.spgemm_csr(A, B, upper, r0, r1)
        assert res.shape == (r1 - r0, b.shape[1])
        blocks.append(res.to_scipy() if res.nnz else sp.csr_matrix((r1 - r0, b.shape[1])))
        res.free()
    assert_csr_equal(sp.vstack(blocks).tocsr(), want, f"csr_dev ranges upper={upper}")
    # row_begin with row_end=None means "to the last row" (device._rows ignored row_begin)
    tail = dev.spgemm_csr(A, B, upper, 500)
    assert tail.shape[0] == 200
    want.sort_indices()
    assert_csr_equal(tail.to_scipy(), want[500:], "csr_dev tail")
    tail.free()


@pytest.mark.parametrize("upper", [False, True])
def test_device_dense_row_ranges(upper):
    a, b = cases.seeded_pair(600, 0.03)
    A, B = dev.DeviceMatrix.from_scipy(a), dev.DeviceMatrix.from_scipy(b)
    want = port.spgemm_dense(a, b, upper)
    got = np.empty_like(want)
    for r0, r1 in _ranges(a.shape[0]):
        if r1 == r0:
            continue
        out = dev.spgemm_dense(A, B, upper, r0, r1)
        got[r0:r1] = out.to_host()
        out.free()
    assert_dense_equal(got, want, f"dense_dev ranges upper={upper}")


@pytest.mark.parametrize("name", ["cfg3s", "cfg5s"])
@pytest.mark.parametrize("upper", [True, False])
def test_device_triple_row_ranges(name, upper):
    w = synthetic.workload(name)
    h, q = w["a"], w["b"]
    H, Q = dev.DeviceMatrix.from_scipy(h), dev.DeviceMatrix.from_scipy(q)
    Ht = H.transpose()
    n = h.shape[0]
    if upper:
        want = port.triple_product(h, q, 0)
    else:
        want = (h @ q @ h.T).toarray()
    got = np.empty((n, n))
    for k, (r0, r1) in enumerate(_ranges(n)):
        if r1 == r0:
            continue
        out = dev.triple_product(H, Q, Ht if k % 2 == 0 else None, upper, r0, r1)     # with and without a cached H^T
        got[r0:r1] = out.to_host()
        out.free()
    np.testing.assert_allclose(got, want, rtol=1e-12, atol=1e-14, err_msg=f"triple_dev ranges {name} upper={upper}")
    if upper:
        assert not np.tril(got, -1).any()


def test_triple_wide_output_overflows_the_shared_window():
    """n = 30,000 output columns: rows near the top are wider than the shared-memory window (~26,800 doubles), so
    their far columns take the L2-reduction path; rows further down fit.  Checked on sampled rows against SciPy."""
    rng = np.random.default_rng(12)
    n, k = 30_000, 120_000
    h = sp.random(n, k, density=3.0 / k, format='csr', random_state=rng)
    q = synthetic.banded_cov(k, half_width=3, length=2.0)
    H, Q = dev.DeviceMatrix.from_scipy(h), dev.DeviceMatrix.from_scipy(q)
    hq = (h @ q).tocsr()
    for r0, r1 in [(0, 40), (2_000, 2_030), (29_950, 30_000)]:
        out = dev.triple_product(H, Q, None, True, r0, r1)
        got = out.to_host()
        out.free()
        want = (hq[r0:r1] @ h.T).toarray()
        for i in range(r0, r1):
            np.testing.assert_allclose(got[i - r0, i:], want[i - r0, i:], rtol=1e-12, atol=1e-14)
            assert not got[i - r0, :i].any()


def test_triple_long_rows_of_ht():
    """H with a few dense columns: rows of H^T far longer than the per-thread sort limit (block bitonic sort) and
    than the sub-warp groups of the contraction."""
    rng = np.random.default_rng(3)
    n, k = 900, 400
    h = sp.random(n, k, density=0.01, format='csr', random_state=rng).tolil()
    h[:, 7] = rng.random((n, 1))
    h[::2, 123] = rng.random((n // 2, 1))
    h = h.tocsr()
    q = cases.banded(k, 3)
    got = sparse_matrix_multiply(h, q, use_triple_product=True)
    assert_dense_equal(got, port.triple_product(h, q, 0), "long H^T rows")
    full = sparse_matrix_multiply(h, q, use_triple_product=True, compute_full_matrix=1)
    assert_dense_equal(full, port.triple_product(h, q, 1), "long H^T rows, reference full mode")


# ---- upper_only through the block-bin kernels -----------------------------------------------------------------
def test_symmetric_sparse_heavy_rows():
    """Rows with > 768 entries take k_symbolic_bitmap / k_numeric_rank; symmetric=True runs them with the column
    window [row, n)."""
    rng = np.random.default_rng(21)
    n = 3000
    a = sp.random(n, n, density=0.02, format='csr', random_state=rng)
    b = sp.random(n, n, density=0.02, format='csr', random_state=rng)
    got = sparse_matrix_multiply(a, b, symmetric=True)
    want = port.spgemm_csr(a, b, True)
    assert np.diff(want.indptr).max() > 768
    assert_csr_equal(got, want, "symmetric heavy rows")
    # wide matrix: compact rank table (> 419k columns) with the upper-triangle window
    n = 600_000
    b = sp.random(n, n, density=2.5 / n, format='csr', random_state=rng)
    rows, cols = [], []
    for i, c in enumerate([0, 3, 400, 1500, 5000, 900, 2500]):
        rows += [i * 50_000] * c
        cols += list(rng.choice(n, size=c, replace=False))
    a = sp.csr_matrix((rng.random(len(rows)), (rows, cols)), shape=(n, n))
    got = sparse_matrix_multiply(a, b, symmetric=True)
    want = port.spgemm_csr(a, b, True)
    assert np.diff(want.indptr).max() > 768
    assert_csr_equal(got, want, "symmetric heavy rows, wide")


# ---- n_gpus through the public API -------------------------------------------------------------------------
def _gpu_counts():
    n = _gpus()
    return [c for c in (1, 2, 3, 4, 8) if c <= max(1, n)]


@pytest.mark.parametrize("name", ["cfg1s", "cfg2s", "cfg3s", "cfg4r10", "cfg5s"])
def test_multi_gpu_public_api(name, monkeypatch):
    monkeypatch.setenv("SPGEMM_B200_FORCE_MULTI", "1")     # n_gpus=1 also goes through the multi-GPU driver
    w = synthetic.workload(name)
    want = port.sparse_matrix_multiply(w["a"], w["b"], **w["kwargs"])
    for n_gpus in _gpu_counts():
        got = sparse_matrix_multiply(w["a"], w["b"], n_gpus=n_gpus, **w["kwargs"])
        if w["kind"] == "sparse":
            assert_csr_equal(got, want, f"{name} n_gpus={n_gpus}")
        else:
            assert_dense_equal(got, want, f"{name} n_gpus={n_gpus}")
        bounds = multi_last_bounds()
        assert len(bounds) == n_gpus + 1 and bounds[0] == 0 and bounds[-1] == w["a"].shape[0]
        assert all(bounds[i] <= bounds[i + 1] for i in range(n_gpus))
        assert len(multi_last_stats()) == n_gpus


def test_multi_gpu_symmetric_and_rectangular(monkeypatch):
    monkeypatch.setenv("SPGEMM_B200_FORCE_MULTI", "1")
    a, b = cases.seeded_pair(500, 0.05)
    r = sp.random(300, 500, density=0.05, format='csr', random_state=np.random.default_rng(2))
    for n_gpus in _gpu_counts():
        assert_csr_equal(sparse_matrix_multiply(a, b, symmetric=True, n_gpus=n_gpus), port.spgemm_csr(a, b, True), "sym sparse")
        assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense', n_gpus=n_gpus), port.spgemm_dense(a, b), "dense")
        assert_dense_equal(sparse_matrix_multiply(r, b, output_format='dense', n_gpus=n_gpus), port.spgemm_dense(r, b), "rect dense")
        assert_csr_equal(sparse_matrix_multiply(r, b, n_gpus=n_gpus), port.spgemm_csr(r, b), "rect sparse")
        # modes that need the whole matrix fall back to one GPU and still honour their contract
        t1 = sparse_matrix_multiply(a, cases.banded(500), use_triple_product=True, compute_full_matrix=1, n_gpus=n_gpus)
        assert_dense_equal(t1, port.triple_product(a, cases.banded(500), 1), "triple1")
    with pytest.raises(RuntimeError, match="GPUs"):
        sparse_matrix_multiply(a, b, n_gpus=_gpus() + 1)


# ---- thread safety ----------------------------------------------------------------------------------------------
def test_concurrent_calls_from_python_threads():
    """ctypes.CDLL releases the GIL: two threads really are inside the library at once (api.cu global
    context).  Each thread multiplies its own pair many times and must always get its own answer."""
    pairs = [cases.seeded_pair(300 + 40 * t, 0.05) for t in range(4)]
    wants = [(port.spgemm_csr(a, b), port.spgemm_dense(a, b, True)) for a, b in pairs]
    errors = []

    def work(t):
        try:
            a, b = pairs[t]
            for _ in range(15):
                assert_csr_equal(sparse_matrix_multiply(a, b), wants[t][0], f"thread {t} sparse")
                assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense', symmetric=True), wants[t][1],
                                   f"thread {t} dense")
        except Exception as ex:          # noqa: BLE001
            errors.append(repr(ex))

    threads = [threading.Thread(target=work, args=(t,)) for t in range(4)]
    for th in threads:
        th.start()
    for th in threads:
        th.join()
    assert not errors, errors[:3]


# ---- validation and canonicalisation -------------------------------------------------------------------------------
def test_invalid_operands_are_rejected_before_any_kernel_reads_them():
    a, b = cases.seeded_pair(200, 0.05)
    bad = b.copy()
    bad.indices = bad.indices.copy()
    bad.indices[17] = b.shape[1] + 5                      # column index out of range
    for kw in (dict(), dict(output_format='dense'), dict(use_triple_product=True)):
        with pytest.raises(RuntimeError, match="out of range|monotone"):
            sparse_matrix_multiply(a, bad, **kw)
    bad = a.copy()
    bad.indices = bad.indices.copy()
    bad.indices[3] = -2
    with pytest.raises(RuntimeError, match="out of range|monotone"):
        sparse_matrix_multiply(bad, b)
    bad = b.copy()
    bad.indptr = bad.indptr.copy()
    bad.indptr[5] = bad.indptr[4] - 1                     # row 4 ends before it starts
    assert (np.diff(bad.indptr) < 0).any()
    with pytest.raises(RuntimeError, match="out of range|monotone"):
        sparse_matrix_multiply(a, bad)
    # the context is still healthy afterwards
    assert_csr_equal(sparse_matrix_multiply(a, b), port.spgemm_csr(a, b), "after rejected operands")


def test_unsorted_b_is_canonicalised_on_the_device():
    rng = np.random.default_rng(8)
    n = 5_000
    b0 = sp.random(n, n, density=4e-3, format='csr', random_state=rng)
    idx, val = b0.indices.copy(), b0.data.copy()
    for r in range(n):                                    # reversed rows; every 7th row also gets a duplicate entry
        s, e = b0.indptr[r], b0.indptr[r + 1]
        idx[s:e], val[s:e] = idx[s:e][::-1], val[s:e][::-1]
        if r % 7 == 0 and e - s >= 2:
            idx[s] = idx[e - 1]
    b = sp.csr_matrix((n, n))
    b.indptr, b.indices, b.data = b0.indptr, idx, val
    a = sp.random(n, n, density=2e-3, format='csr', random_state=rng)
    A, B = dev.DeviceMatrix.from_scipy(a), dev.DeviceMatrix.from_scipy(b)
    assert not B.is_sorted() and A.is_sorted()
    res = dev.spgemm_csr(A, B, True)                      # windowed (upper_only): wants binary search in B's rows
    assert_csr_equal(res.to_scipy(), port.spgemm_csr(a, b, True), "unsorted B, device API")
    res.free()
    assert B.is_sorted()                                   # sorted in place on the device: no filter fallback left
    out = dev.spgemm_dense(A, B, True)
    assert_dense_equal(out.to_host(), port.spgemm_dense(a, b, True), "unsorted B dense")
    out.free()
    # borrowed arrays are never modified: a sorted shadow copy is used instead
    import ctypes
    lib = matrix_ops.get_lib()
    arrs = [np.ascontiguousarray(x) for x in (b.indptr.astype(np.int32), b.indices.astype(np.int32), b.data)]
    ptrs = []
    for x in arrs:
        p = lib.spgemm_b200_device_alloc(x.nbytes)
        lib.spgemm_b200_copy_to_device(ctypes.c_void_p(p), x.ctypes.data_as(ctypes.c_void_p), x.nbytes)
        ptrs.append(p)
    W = dev.DeviceMatrix.wrap(b.shape, b.nnz, *ptrs)
    res = dev.spgemm_csr(A, W, True)
    assert_csr_equal(res.to_scipy(), port.spgemm_csr(a, b, True), "unsorted borrowed B")
    res.free()
    back = np.empty_like(arrs[1])
    lib.spgemm_b200_copy_to_host(back.ctypes.data_as(ctypes.c_void_p), ctypes.c_void_p(ptrs[1]), back.nbytes)
    assert np.array_equal(back, arrs[1])                   # the caller's indices are untouched
    W.free()
    for p in ptrs:
        lib.spgemm_b200_device_free(ctypes.c_void_p(p))


def test_trim_releases_cached_memory():
    a, b = cases.seeded_pair(400, 0.05)
    sparse_matrix_multiply(a, b, output_format='dense')
    assert matrix_ops.get_lib().spgemm_b200_trim(0) == 0
    assert_dense_equal(sparse_matrix_multiply(a, b, output_format='dense'), port.spgemm_dense(a, b), "after trim")


# ---- triple product: column panels and both kernels at sizes the oracle checks in full -------------------------
@pytest.mark.parametrize("panels", [1, 2, 3, 5])
@pytest.mark.parametrize("generic", [0, 1])
def test_triple_panels_forced(panels, generic, monkeypatch):
    """The default plan only cuts C into several column panels for outputs wider than the shared-memory segment or
    an H^T beyond the L2 budget (cfg 5); here the panel count is forced so that the multi-panel bookkeeping -- which
    panel holds a row's diagonal, who writes the zeros left of it, the (panel, row) ticket order -- runs on small
    inputs, with the lean kernel for banded Q (rows of Q are runs of consecutive columns) and with the general one."""
    monkeypatch.setenv("SPGEMM_B200_TRIPLE_PANELS", str(panels))
    monkeypatch.setenv("SPGEMM_B200_TRIPLE_GENERIC", str(generic))
    for name in ("cfg3s", "cfg5s"):
        w = synthetic.workload(name)
        h, q = w["a"], w["b"]
        assert_dense_equal(sparse_matrix_multiply(h, q, use_triple_product=True), port.triple_product(h, q, 0),
                           f"{name} panels={panels} generic={generic} upper")
        assert_dense_equal(sparse_matrix_multiply(h, q, use_triple_product=True, compute_full_matrix=1),
                           port.triple_product(h, q, 1), f"{name} panels={panels} generic={generic} reference-full")
    # row ranges that start inside a panel, through the device API
    w = synthetic.workload("cfg3s")
    h, q = w["a"], w["b"]
    H, Q = dev.DeviceMatrix.from_scipy(h), dev.DeviceMatrix.from_scipy(q)
    want = port.triple_product(h, q, 0)
    for r0, r1 in [(0, 37), (37, 250), (250, 400)]:
        out = dev.triple_product(H, Q, None, True, r0, r1)
        np.testing.assert_allclose(out.to_host(), want[r0:r1], rtol=1e-12, atol=1e-14)
        out.free()


def test_triple_q_with_wide_and_ragged_runs():
    """Banded Q whose runs are longer than one pass of the weight table (96 columns), ragged at the matrix edges, next
    to an H with empty rows and a duplicate column entry."""
    rng = np.random.default_rng(4)
    n, k = 300, 2500
    h = sp.random(n, k, density=0.01, format='csr', random_state=rng).tolil()
    h[5, :] = 0
    h[17, :] = 0
    h = h.tocsr()
    h.eliminate_zeros()
    # duplicate entry: row 3 names its first column twice (CSR inputs are used as they are)
    s, e = h.indptr[3], h.indptr[4]
    if e - s >= 2:
        h.indices[s + 1] = h.indices[s]
    q = synthetic.banded_cov(k, half_width=120, length=30.0)          # runs of up to 241 columns
    got = sparse_matrix_multiply(h, q, use_triple_product=True)
    assert_dense_equal(got, port.triple_product(h, q, 0), "wide runs")
    # a Q that is sorted but not made of runs takes the general kernel
    q2 = sp.random(k, k, density=0.004, format='csr', random_state=rng)
    assert_dense_equal(sparse_matrix_multiply(h, q2, use_triple_product=True), port.triple_product(h, q2, 0), "general Q")


def test_multi_gpu_without_peer_copies(monkeypatch):
    """The whole-operand upload path of the multi-GPU driver (taken when the GPUs cannot reach each other)."""
    monkeypatch.setenv("SPGEMM_B200_FORCE_MULTI", "1")
    monkeypatch.setenv("SPGEMM_B200_MULTI_NO_PEER", "1")
    w = synthetic.workload("cfg3s")
    want = port.sparse_matrix_multiply(w["a"], w["b"], **w["kwargs"])
    for n_gpus in _gpu_counts():
        assert_dense_equal(sparse_matrix_multiply(w["a"], w["b"], n_gpus=n_gpus, **w["kwargs"]), want, f"no-peer n_gpus={n_gpus}")


def test_triple_cached_transpose():
    """DeviceMatrix.cache_transpose: the paneled transpose is kept on the H handle and reused by later calls with the
    same panel plan, rebuilt when the plan changes (another first row), dropped on request."""
    w = synthetic.workload("cfg3s")
    h, q = w["a"], w["b"]
    H, Q = dev.DeviceMatrix.from_scipy(h), dev.DeviceMatrix.from_scipy(q)
    want = port.triple_product(h, q, 0)
    H.cache_transpose(True)
    for r0, r1 in [(0, 400), (0, 400), (100, 300), (100, 300), (0, 400)]:
        out = dev.triple_product(H, Q, None, True, r0, r1)
        np.testing.assert_allclose(out.to_host(), want[r0:r1], rtol=1e-12, atol=1e-14)
        out.free()
    H.cache_transpose(False)
    out = dev.triple_product(H, Q, None, True, 0, 400)
    np.testing.assert_allclose(out.to_host(), want, rtol=1e-12, atol=1e-14)
    out.free()


# ---- partition with a start-dependent fixed cost (triple product over several GPUs) ------------------------
def _block_cost(costs, tail, coeff, r0, r1):
    return (coeff * tail[r0] if r1 > r0 else 0.0) + float(np.sum(costs[r0:r1] + 1))


@pytest.mark.parametrize("parts", [2, 3, 8])
def test_partition_tail_minimises_the_largest_block(parts):
    """spgemm_b200_partition_tail: bounds cover the rows, and no neighbouring single-row shift of one cut lowers
    the largest block cost (block cost = coeff * entries of H from its first row on + its rows' costs); with
    coeff = 0 the largest block is no worse than that of the equal-sum partition."""
    w = synthetic.workload("cfg3s")
    h, q = w["a"], w["b"]
    H, Q = dev.DeviceMatrix.from_scipy(h), dev.DeviceMatrix.from_scipy(q)
    Ht = H.transpose()
    d_costs, total = dev.row_costs(H, Ht, Q, upper_only=True)
    n = h.shape[0]
    costs = np.zeros(n, dtype=np.int64)
    dev.copy_to_host(costs, d_costs)
    assert int(costs.sum()) == total
    tail = (h.indptr[-1] - h.indptr).astype(np.float64)
    for coeff in (0.0, 8.6, 500.0):
        b = dev.partition_rows(d_costs, n, parts, tail_indptr=h.indptr, tail_coeff=coeff)
        assert b[0] == 0 and b[-1] == n and np.all(np.diff(b) >= 0)
        worst = max(_block_cost(costs, tail, coeff, b[p], b[p + 1]) for p in range(parts))
        # greedy bottleneck: the bound is tight to one row -- removing the last row of the worst block's predecessors
        # cannot help, so compare with every partition that moves ONE cut by one row
        for p in range(1, parts):
            for step in (-1, 1):
                c = b.copy()
                c[p] += step
                if c[p] < c[p - 1] or c[p] > c[p + 1]:
                    continue
                alt = max(_block_cost(costs, tail, coeff, c[k], c[k + 1]) for k in range(parts))
                assert alt >= worst - max(costs.max() + 1, 1)
        if coeff == 0.0:
            e = dev.partition_rows(d_costs, n, parts)
            even = max(_block_cost(costs, tail, 0.0, e[p], e[p + 1]) for p in range(parts))
            assert worst <= even
        if coeff == 500.0 and parts > 1:
            e = dev.partition_rows(d_costs, n, parts)
            assert b[1] <= e[1]            # the first block pays the whole transpose, so it gets fewer rows
    matrix_ops.get_lib().spgemm_b200_device_free(d_costs)
    for x in (Ht, H, Q):
        x.free()


# ---- banded-Q kernel: stream pipeline across ranges, empty rows of Q ---------------------------------------
def _drop_rows(q, rows):
    """Q with the given rows emptied (still one run of consecutive columns per non-empty row)."""
    q = q.tolil()
    for r in rows:
        q[r, :] = 0
    q = q.tocsr()
    q.eliminate_zeros()
    q.sort_indices()
    return q


@pytest.mark.parametrize("half_width,h_density", [(0, 0.02), (2, 0.002), (2, 0.02), (20, 0.002), (20, 0.02),
                                                  (120, 0.004)])
def test_triple_runs_kernel_short_ranges_and_empty_q_rows(half_width, h_density):
    """k_triple_runs: ranges of H^T of 0, 1, 2 and many 32-entry steps, runs longer than one weight-table pass, and
    rows of Q with no entries at all -- an entry of H that points at one contributes nothing and must not disturb the
    entry its warp takes next (whose prefetched loads it used to skip: the first 96 products of that entry were lost)."""
    rng = np.random.default_rng(100 * half_width + 1)
    n, k = 260, 3000
    h = sp.random(n, k, density=h_density, format='csr', random_state=rng)
    h.sort_indices()
    q = synthetic.banded_cov(k, half_width=half_width, length=7.0)
    # empty every 7th row of Q and the rows the first entries of rows 0..39 of H point at
    drop = set(range(3, k, 7))
    for i in range(40):
        if h.indptr[i + 1] > h.indptr[i]:
            drop.add(int(h.indices[h.indptr[i]]))
    q = _drop_rows(q, sorted(drop))
    assert (np.diff(q.indptr) == 0).sum() >= len(drop)
    want = port.triple_product(h, q, 0)
    got = sparse_matrix_multiply(h, q, use_triple_product=True)
    assert_dense_equal(got, want, f"runs kernel hw={half_width}")
    # the full (non-symmetric) product through the device API takes the unfiltered loop in every panel
    H, Q = dev.DeviceMatrix.from_scipy(h), dev.DeviceMatrix.from_scipy(q)
    full = (h @ q @ h.T).toarray()
    out = dev.triple_product(H, Q, None, False, 0, n)
    np.testing.assert_allclose(out.to_host(), full, rtol=1e-12, atol=1e-13)
    out.free()
    H.free()
    Q.free()
