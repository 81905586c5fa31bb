"""The reference's own pytest cases, run against the CUDA path (VERDICT r1 "missing" #6).

Each test restates one reference test -- same generator calls, sizes, densities and seeds, same truth source
(SciPy `A.dot(B)` / NumPy), same comparison (`np.allclose` at its defaults, `np.triu` for the symmetric modes) --
with `.A` spelled `.toarray()` (SciPy >= 1.14 removed `.A`; that is the only reason 9 of the reference's 24 tests
fail in this image).  The timing printouts of the reference tests are not reproduced: they assert nothing.

  /root/reference/tests/test_with_dense.py:30-109      unseeded random, sparse output
  /root/reference/tests/test_basic.py:25-61            500^2 at 1 %, triple product (the reference only prints the
                                                       comparison; here it is asserted)
  /root/reference/tests/test_computation_speed.py:8-87 500^2 at 30 %, seeds 42/43, stats.uniform values, all modes
(tests/test_matrix_multiply.py and tests/test_edge_case.py are restated in test_gpu_parity.py::test_reference_known_answers
 and pinned bit for bit by the golden vectors.)
"""
import numpy as np
import pytest
from scipy import stats
from scipy.sparse import csr_matrix, eye
from scipy.sparse import random as sparse_random

from sparse_matrix_mult_b200 import sparse_matrix_multiply

pytestmark = pytest.mark.gpu


# ---- tests/test_with_dense.py ------------------------------------------------------------------------
@pytest.mark.parametrize("size,density", [(5, 0.01), (5, 0.1), (5, 0.3), (6, 0.01), (6, 0.1), (6, 0.3)])
def test_different_sparsity_levels(size, density):                       # test_with_dense.py:30-49
    a = sparse_random(size, size, density=density, format='csr')
    b = sparse_random(size, size, density=density, format='csr')
    got = sparse_matrix_multiply(a, b, output_format='sparse', symmetric=False)
    assert np.allclose(got.toarray(), a.dot(b).toarray()), f"size {size}, density {density}"


def test_non_square_sparse_matrices():                                   # test_with_dense.py:51-68
    a = sparse_random(500, 400, density=0.1, format='csr')
    b = sparse_random(400, 500, density=0.1, format='csr')
    got = sparse_matrix_multiply(a, b, output_format='sparse', symmetric=False)
    assert np.allclose(got.toarray(), a.dot(b).toarray())


def test_sparse_identity_matrix_multiplication():                        # test_with_dense.py:70-88
    size = 500
    a = sparse_random(size, size, density=0.1, format='csr')
    i = eye(size, format='csr')
    got = sparse_matrix_multiply(a, i, output_format='sparse', symmetric=False)
    assert np.allclose(got.toarray(), a.dot(i).toarray())


def test_large_sparse_matrix_multiplication():                           # test_with_dense.py:90-109
    size, density = 1000, 0.01
    a = sparse_random(size, size, density=density, format='csr')
    b = sparse_random(size, size, density=density, format='csr')
    got = sparse_matrix_multiply(a, b, output_format='sparse', symmetric=False)
    assert np.allclose(got.toarray(), a.dot(b).toarray())


# ---- tests/test_basic.py -----------------------------------------------------------------------------
def test_basic_triple_product():                                         # test_basic.py:8-11, 25-61
    a = sparse_random(500, 500, density=0.01, format='csr')
    b = sparse_random(500, 500, density=0.01, format='csr')
    got = sparse_matrix_multiply(a, b, use_triple_product=True, compute_full_matrix=0)
    want = a.dot(b).dot(a.transpose()).toarray()
    assert got.ndim == 2 and want.ndim == 2
    mask = np.triu(np.ones(got.shape, dtype=bool))
    assert np.allclose(got[mask], want[mask], rtol=1e-5, atol=1e-8)


# ---- tests/test_computation_speed.py -----------------------------------------------------------------
@pytest.fixture
def setup_matrices():                                                    # test_computation_speed.py:8-15
    def _setup(rows_a=500, cols_a=500, rows_b=500, cols_b=500, density=0.3):
        a = sparse_random(rows_a, cols_a, density=density, random_state=42, data_rvs=stats.uniform().rvs)
        b = sparse_random(rows_b, cols_b, density=density, random_state=43, data_rvs=stats.uniform().rvs)
        return csr_matrix(a), csr_matrix(b)
    return _setup


def test_sparse_sparse_non_symmetric(setup_matrices):                    # :37-44
    a, b = setup_matrices()
    got = sparse_matrix_multiply(a, b, output_format='sparse', symmetric=False)
    assert np.allclose(got.toarray(), a.dot(b).toarray())


def test_sparse_sparse_symmetric(setup_matrices):                        # :46-54
    a, b = setup_matrices()
    got = sparse_matrix_multiply(a, b, output_format='sparse', symmetric=True)
    assert np.allclose(np.triu(a.dot(b).toarray()), np.triu(got.toarray()))


def test_sparse_dense_symmetric(setup_matrices):                         # :56-64
    a, b = setup_matrices()
    got = sparse_matrix_multiply(a, b, output_format='dense', symmetric=True)
    assert np.allclose(np.triu(a.dot(b).toarray()), np.triu(got))


def test_sparse_dense_non_symmetric(setup_matrices):                     # :66-74
    a, b = setup_matrices()
    got = sparse_matrix_multiply(a, b, output_format='dense', symmetric=False)
    assert np.allclose(a.dot(b).toarray(), got)


def test_triple_product(setup_matrices):                                 # :76-87
    a, b = setup_matrices()
    got = sparse_matrix_multiply(a, b, use_triple_product=True, compute_full_matrix=0)
    want = a.dot(b).dot(a.transpose()).toarray()
    assert np.allclose(np.triu(want), np.triu(got))
