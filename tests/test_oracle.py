"""CPU tests that pin the oracle (oracle/spgemm_oracle.c) to the reference.

 * against the committed golden vectors produced by the unmodified reference (tests/golden/make_golden.py);
   bit-for-bit, including the first-touch column order;
 * against the reference's own test expectations (NumPy / SciPy products, np.triu for symmetric modes),
   restating tests/test_matrix_multiply.py:89-112, tests/test_edge_case.py:42-71,
   tests/test_computation_speed.py:37-87 of the reference;
 * against oracle/_ref (the reference binaries) when they are present.
"""
import numpy as np
import pytest
import scipy.sparse as sp

import cases
from helpers import golden_expected
from oracle import port, ref


def _case_ids():
    return [c[0] for c in cases.all_cases()]


@pytest.mark.parametrize("case", cases.all_cases(), ids=_case_ids())
def test_oracle_matches_golden_bit_exact(golden, case):
    name, a, b, kw = case
    want = golden_expected(golden, name)
    got = port.sparse_matrix_multiply(a, b, **kw)
    if isinstance(want, np.ndarray):
        assert isinstance(got, np.ndarray) and got.shape == want.shape
        assert np.array_equal(got, want), name
    else:
        assert got.shape == want.shape
        assert np.array_equal(got.indptr, want.indptr), name
        assert np.array_equal(got.indices, want.indices), name       # same first-touch order as the reference
        assert np.array_equal(got.data, want.data), name


def test_golden_names_match_cases(golden):
    assert list(golden["__names__"]) == _case_ids()


def test_reference_known_answers():
    m = cases.fixed_matrices()
    cd = np.matmul(m["C"], m["D"])
    assert np.allclose(port.sparse_matrix_multiply(m["C"], m["D"]).toarray(), cd)
    assert np.allclose(port.sparse_matrix_multiply(m["C"], m["D"], output_format='dense'), cd)
    cf = np.matmul(m["C"], m["F"])
    assert np.allclose(np.triu(port.sparse_matrix_multiply(m["C"], m["F"], output_format='dense', symmetric=True)), np.triu(cf))
    assert np.allclose(np.triu(port.sparse_matrix_multiply(m["C"], m["F"], symmetric=True).toarray()), np.triu(cf))
    assert np.allclose(port.sparse_matrix_multiply(m["one_a"], m["one_b"], output_format='dense', symmetric=True), [[10]])


def test_reference_seeded_random_vs_scipy():
    a, b = cases.seeded_pair(200)
    full = (a @ b).toarray()
    assert np.allclose(port.spgemm_csr(a, b).toarray(), full)
    assert np.allclose(port.spgemm_csr(a, b, True).toarray(), np.triu(full))
    assert np.allclose(port.spgemm_dense(a, b), full)
    assert np.allclose(port.spgemm_dense(a, b, True), np.triu(full))
    t = (a @ b @ a.T).toarray()
    assert np.allclose(port.triple_product(a, b, 0), np.triu(t))
    assert np.allclose(port.triple_product(a, b, 1), t + t.T - np.diag(np.diag(t)))


def test_limits_matches_reference_layout():
    assert port.limits(10, 3) == [(0, 3), (4, 6), (7, 9)]
    assert port.limits(3, 8) == [(0, 0), (1, 1), (2, 2)]
    assert port.limits(8, 4) == [(0, 1), (2, 3), (4, 5), (6, 7)]


@pytest.mark.skipif(not ref.available(), reason="oracle/_ref not built (make -C oracle ref)")
def test_oracle_matches_reference_binaries():
    rng = np.random.default_rng(5)
    a = sp.random(300, 200, density=0.05, format='csr', random_state=rng)
    b = sp.random(200, 300, density=0.05, format='csr', random_state=rng)
    for upper in (False, True):
        got, want = port.spgemm_csr(a, b, upper), ref.shipped().sparse(a, b, upper)
        assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
        assert np.array_equal(got.data, want.data)
        d = port.spgemm_dense(a, b, upper)
        assert np.array_equal(d, ref.shipped().dense(a, b, upper))
        assert np.array_equal(d, ref.omp().dense(a, b, upper))
    h = sp.random(100, 400, density=0.05, format='csr', random_state=rng)
    q = sp.random(400, 400, density=0.02, format='csr', random_state=rng)
    for full in (0, 1):
        t = port.triple_product(h, q, full)
        assert np.array_equal(t, ref.shipped().triple(h, q, full))
        np.testing.assert_allclose(t, ref.omp().triple(h, q, full), rtol=1e-13, atol=1e-15)


def test_openmp_port_is_bit_identical_to_the_serial_port():
    """The many-core CPU baseline of bench.py (oracle_spgemm_csr_omp) computes every row exactly like the serial
    restatement, whatever the number of row blocks."""
    from sparse_matrix_mult_b200 import synthetic
    a = synthetic.rmat(11)
    for upper in (False, True):
        want = port.spgemm_csr(a, a, upper)
        for blocks in (1, 3, 64, 5000):
            got = port.spgemm_csr(a, a, upper, omp_blocks=blocks)
            assert np.array_equal(got.indptr, want.indptr) and np.array_equal(got.indices, want.indices)
            assert np.array_equal(got.data, want.data)
    rng = np.random.default_rng(2)
    x = sp.random(37, 53, density=0.2, format='csr', random_state=rng)
    y = sp.random(53, 41, density=0.2, format='csr', random_state=rng)
    got, want = port.spgemm_csr(x, y, omp_blocks=8), port.spgemm_csr(x, y)
    assert np.array_equal(got.indices, want.indices) and np.array_equal(got.data, want.data)


def test_patched_openmp_reference_is_bit_identical_to_the_shipped_binary():
    """today's src/ with the six repairs applied at build time (oracle/build_patched_ref.py)
    reproduces the shipped serial binary bit for bit -- structure in first-touch order and values -- at any thread
    count, on cfg 1 and on R-MAT scale 12 (power-law rows), both sparse modes."""
    from oracle import ref
    if not (ref.available() and ref.patched_available()):
        pytest.skip("oracle/_ref not built (make -C oracle ref)")
    from sparse_matrix_mult_b200 import synthetic
    for name in ("cfg1", "cfg4r12"):
        w = synthetic.workload(name)
        for sym in (False, True):
            a = ref.shipped().sparse(w["a"], w["b"], sym)
            b = ref.omp_patched().sparse(w["a"], w["b"], sym)
            assert a.nnz == b.nnz and np.array_equal(a.indptr, b.indptr), (name, sym)
            assert np.array_equal(a.indices, b.indices) and np.array_equal(a.data, b.data), (name, sym)
