"""Deterministic inputs for the golden vectors (shared by make_golden.py and the tests).

Every case is (name, matrix_a, matrix_b, kwargs-for-sparse_matrix_multiply).  The fixed matrices restate the
data of the reference's own known-answer tests:
  tests/test_matrix_multiply.py:7-87   C (9x12), D (12x6), F (12x9) integer/decimal grids, A/B 8x8
  tests/test_edge_case.py:6-40         1x1, a 6x3 matrix with three zero rows
  sparse_matrix_mult/matrix_ops_test_script.py:28-59   4x4 / 3x4 / 4x3 demo matrices
  tests/test_computation_speed.py:8-15 seeded sparse_random(density=.3, random_state=42/43) -- here at
                                       120x120 so the fixture stays small
plus cases the reference never tests (SURVEY.md section 4 "Not tested anywhere"): unsorted CSR input with
duplicate entries, exact cancellation (explicit zeros kept), rectangular shapes, banded triple product.
Only numpy/scipy generators with fixed seeds: the same arrays come out on the GPU box.
"""
import numpy as np
from scipy.sparse import csr_matrix, diags, random as sparse_random

MODES = {
    "sparse": dict(output_format="sparse", symmetric=False),
    "sparse_sym": dict(output_format="sparse", symmetric=True),
    "dense": dict(output_format="dense", symmetric=False),
    "dense_sym": dict(output_format="dense", symmetric=True),
}


def fixed_matrices():
    m = {}
    m["C"] = np.arange(1, 109, dtype=np.int64).reshape(9, 12)
    m["D"] = np.arange(1, 73, dtype=np.float64).reshape(12, 6) / 10.0
    m["F"] = np.arange(1, 109, dtype=np.int64).reshape(12, 9)
    m["A8"] = np.array([
        [0.64, 0.99, 0.89, 0.72, 0, 0, 0, 0],
        [0, 0.67, 0.54, 0, 0.81, 0, 0, 0],
        [0, 0.32, 0, 0, 0, 0.45, 0, 0],
        [0.1, 0, 0, 0, 0, 0, 0.23, 0],
        [0, 0, 0.78, 0, 0.55, 0, 0, 0.91],
        [0.43, 0, 0, 0.12, 0, 0, 0, 0],
        [0, 0, 0.33, 0, 0, 0.68, 0, 0],
        [0, 0.21, 0, 0, 0, 0, 0.39, 0]])
    m["B8"] = np.array([
        [0.23, 0, 0, 0, 0.51, 0, 0, 0],
        [0, 0.72, 0, 0, 0, 0.38, 0, 0],
        [0, 0, 0.99, 0, 0, 0, 0.84, 0],
        [0, 0.76, 0.87, 0.97, 0, 0, 0, 0.29],
        [0.15, 0, 0, 0, 0.62, 0, 0, 0],
        [0, 0.44, 0, 0, 0, 0.75, 0, 0],
        [0, 0, 0.58, 0, 0, 0, 0.93, 0],
        [0.36, 0, 0, 0.82, 0, 0, 0, 0.47]])
    m["A4"] = np.array([[0.64, 0.99, 0.89, 0.72], [0, 0.67, 0.54, 0], [0, 0.32, 0, 0], [0.1, 0, 0, 0]])
    m["B4"] = np.array([[0.23, 0, 0, 0.51], [0, 0.72, 0, 0], [0, 0, 0.99, 0], [0, 0.76, 0.87, 0.97]])
    m["C34"] = m["A4"][:3].copy()
    m["D43"] = np.array([[0.64, 0.99, 0.89], [0, 0.67, 0.54], [0, 0.32, 0], [0.1, 0, 0]])
    m["one_a"] = np.array([[5]])
    m["one_b"] = np.array([[2]])
    m["zero_rows"] = np.array([[1, 2, 3], [4, 5, 6], [7, 8, 9], [0, 0, 0], [0, 0, 0], [0, 0, 0]])
    m["rand34"] = np.random.default_rng(7).random((3, 4))
    return m


def seeded_pair(n=120, density=0.3):
    # the reference draws the VALUES from SciPy's unseeded global stream (data_rvs=stats.uniform().rvs);
    # here they are seeded too so that the fixture is reproducible
    a = csr_matrix(sparse_random(n, n, density=density, random_state=42,
                                 data_rvs=np.random.default_rng(142).random))
    b = csr_matrix(sparse_random(n, n, density=density, random_state=43, data_rvs=np.random.default_rng(143).random))
    return a, b


def banded(n, half=4, scale=2.0):
    offs = list(range(-half, half + 1))
    return csr_matrix(diags([np.full(n - abs(o), np.exp(-abs(o) / scale)) for o in offs], offs, format="csr"))


def unsorted_with_duplicates():
    """A CSR whose rows are unsorted and contain a repeated column (never canonicalised by the reference,
    matrix_ops.py:307-310 leaves CSR inputs alone)."""
    indptr = np.array([0, 4, 6, 6, 9], dtype=np.int32)
    indices = np.array([3, 0, 3, 1, 2, 0, 4, 1, 4], dtype=np.int32)
    data = np.array([1.5, -2.0, 0.25, 3.0, 4.0, -1.0, 2.0, 0.5, 1.0])
    a = csr_matrix((4, 5))
    a.indptr, a.indices, a.data = indptr, indices, data
    b = csr_matrix(np.random.default_rng(11).integers(-2, 3, size=(5, 4)).astype(np.float64))
    return a, b


def cancelling():
    """Row 0 of A*B has an entry that cancels to exactly 0.0: the reference keeps it."""
    a = csr_matrix(np.array([[1.0, 1.0, 0.0], [0.0, 2.0, 1.0], [1.0, 0.0, 0.0]]))
    b = csr_matrix(np.array([[1.0, 3.0, 0.0], [-1.0, 0.0, 2.0], [0.0, 0.0, -4.0]]))
    return a, b


def all_cases():
    m = fixed_matrices()
    out = []

    def add(name, a, b, **kw):
        out.append((name, a, b, kw))

    add("CD/sparse", m["C"], m["D"], **MODES["sparse"])
    add("CD/dense", m["C"], m["D"], **MODES["dense"])
    add("CF/dense_sym", m["C"], m["F"], **MODES["dense_sym"])
    add("CF/sparse_sym", m["C"], m["F"], **MODES["sparse_sym"])
    for mode, kw in MODES.items():
        add(f"A8B8/{mode}", m["A8"], m["B8"], **kw)
        add(f"A4B4/{mode}", csr_matrix(m["A4"]), csr_matrix(m["B4"]), **kw)
    add("C34D43/dense", csr_matrix(m["C34"]), csr_matrix(m["D43"]), **MODES["dense"])
    add("D43C34/sparse", csr_matrix(m["D43"]), csr_matrix(m["C34"]), **MODES["sparse"])
    add("A4B4/triple0", csr_matrix(m["A4"]), csr_matrix(m["B4"]), use_triple_product=True, compute_full_matrix=0)
    add("A4B4/triple1", csr_matrix(m["A4"]), csr_matrix(m["B4"]), use_triple_product=True, compute_full_matrix=1)
    add("C34B4/triple0", csr_matrix(m["C34"]), csr_matrix(m["B4"]), use_triple_product=True)
    add("one/dense_sym", m["one_a"], m["one_b"], **MODES["dense_sym"])
    add("zero_rows/dense", m["zero_rows"], m["rand34"], **MODES["dense"])
    add("zero_rows/sparse", csr_matrix(m["zero_rows"]), csr_matrix(m["rand34"]), **MODES["sparse"])

    a, b = seeded_pair()
    for mode, kw in MODES.items():
        add(f"seeded120/{mode}", a, b, **kw)
    add("seeded120/triple0", a, b, use_triple_product=True, compute_full_matrix=0)
    add("seeded120/triple1", a, b, use_triple_product=True, compute_full_matrix=1)

    rng = np.random.default_rng(3)
    r1 = csr_matrix(sparse_random(60, 80, density=0.08, random_state=rng))
    r2 = csr_matrix(sparse_random(80, 50, density=0.1, random_state=rng))
    add("rect/sparse", r1, r2, **MODES["sparse"])
    add("rect/dense", r1, r2, **MODES["dense"])
    r3 = csr_matrix(r1.T)
    add("rectAAt/dense_sym", r1, r3, **MODES["dense_sym"])
    add("rectAAt/sparse_sym", r1, r3, **MODES["sparse_sym"])

    h = csr_matrix(sparse_random(40, 150, density=0.06, random_state=rng))
    q = banded(150)
    add("band/triple0", h, q, use_triple_product=True, compute_full_matrix=0)
    add("band/triple1", h, q, use_triple_product=True, compute_full_matrix=1)

    ua, ub = unsorted_with_duplicates()
    for mode, kw in MODES.items():
        add(f"unsorted_dup/{mode}", ua, ub, **kw)
    # the same non-canonical matrix on the right-hand side (duplicate columns inside a row of B)
    ua2 = csr_matrix(np.random.default_rng(12).integers(-2, 3, size=(6, 4)).astype(np.float64))
    for mode in ("sparse", "dense"):
        add(f"unsorted_dupB/{mode}", ua2, ua, **MODES[mode])
    ca, cb = cancelling()
    add("cancel/sparse", ca, cb, **MODES["sparse"])
    add("cancel/sparse_sym", ca, cb, **MODES["sparse_sym"])
    add("cancel/dense", ca, cb, **MODES["dense"])
    return out
