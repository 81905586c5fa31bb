"""CPU-side checks of the boundary: the library loads without a GPU and exports every symbol that
include/spgemm_b200.h declares; the Python mirror keeps the reference's signature and pre-dispatch behaviour;
and the product path fails loudly (no CPU fallback) when there is no device."""
import ctypes
import inspect
import os
import re

import numpy as np
import pytest
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "spgemm_b200.h")).read()
    return sorted(set(re.findall(r"SPGEMM_B200_API[^;]*?\b(spgemm_b200_\w+)\s*\(", text)))


def test_header_declares_the_expected_entry_points():
    names = _declared_symbols()
    for must in ("spgemm_b200_csr", "spgemm_b200_dense", "spgemm_b200_triple", "spgemm_b200_result_copy",
                 "spgemm_b200_result_free", "spgemm_b200_host_alloc", "spgemm_b200_csr_dev", "spgemm_b200_dense_dev",
                 "spgemm_b200_triple_dev", "spgemm_b200_row_costs", "spgemm_b200_partition", "spgemm_b200_last_error"):
        assert must in names
    assert len(names) >= 35


def test_library_loads_and_exports_every_declared_symbol():
    from sparse_matrix_mult_b200.matrix_ops import matrix_ops
    lib = matrix_ops.get_lib()
    missing = [n for n in _declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"declared in include/spgemm_b200.h but not exported: {missing}"
    assert lib.spgemm_b200_version().startswith(b"spgemm_b200")
    assert lib.spgemm_b200_device_count() >= 0


def test_library_name_cannot_shadow_the_reference_loader():
    # reference loader pattern: sparse_matrix_mult/matrix_ops.py:118
    from sparse_matrix_mult_b200.matrix_ops import MatrixOpsLibrary
    pat = re.compile(r'libsparse(?:[_]?(x86_64|x86_x64|arm64))?\.so')
    assert not pat.match(MatrixOpsLibrary.LIB_NAME)


def test_signature_matches_reference():
    # sparse_matrix_mult/matrix_ops.py:271-272 (mirror= is the one documented extension, keyword with default)
    from sparse_matrix_mult_b200 import sparse_matrix_multiply
    params = list(inspect.signature(sparse_matrix_multiply).parameters.items())
    names = [n for n, _ in params]
    assert names[:7] == ["matrix_a", "matrix_b", "output_format", "symmetric", "imem_size", "use_triple_product",
                         "compute_full_matrix"]
    defaults = {n: p.default for n, p in params}
    assert defaults["output_format"] == 'sparse' and defaults["symmetric"] is False and defaults["imem_size"] is None
    assert defaults["use_triple_product"] is False and defaults["compute_full_matrix"] is None
    import sparse_matrix_mult_b200
    assert sparse_matrix_mult_b200.__all__ == ['sparse_matrix_multiply']


def test_pre_dispatch_behaviour_needs_no_gpu():
    """Everything the reference does before touching the C library (matrix_ops.py:288-322)."""
    from sparse_matrix_mult_b200 import sparse_matrix_multiply
    a = sp.random(5, 7, density=0.5, format='csr', random_state=1)
    with pytest.raises(ValueError, match="incompatible"):
        sparse_matrix_multiply(a, sp.random(6, 5, density=0.5, format='csr', random_state=2))
    with pytest.raises(ValueError, match="square"):
        sparse_matrix_multiply(a, sp.random(7, 4, density=0.5, format='csr', random_state=2), symmetric=True)
    with pytest.raises(ValueError, match="compute_full_matrix"):
        sparse_matrix_multiply(a, a.T, compute_full_matrix=3)
    with pytest.raises(ValueError, match="imem_size"):
        sparse_matrix_multiply(a, a.T, imem_size="many")
    z = sparse_matrix_multiply(np.zeros((3, 3)), np.zeros((3, 4)))
    assert sp.isspmatrix_csr(z) and z.shape == (3, 4) and z.nnz == 0
    z = sparse_matrix_multiply(sp.csr_matrix((3, 3)), np.ones((3, 4)), output_format='dense')
    assert isinstance(z, np.ndarray) and z.shape == (3, 4) and not z.any()
    z = sparse_matrix_multiply(np.ones((2, 3)), np.ones((3, 2)), output_format='nonsense')
    assert isinstance(z, np.ndarray) and z.shape == (2, 2) and not z.any()


def test_no_cpu_fallback():
    from sparse_matrix_mult_b200 import sparse_matrix_multiply
    from sparse_matrix_mult_b200.matrix_ops import matrix_ops
    if matrix_ops.get_lib().spgemm_b200_device_count() > 0:
        pytest.skip("a GPU is present")
    with pytest.raises(RuntimeError, match="no CUDA device|CUDA"):
        sparse_matrix_multiply(np.eye(4), np.eye(4))
    with pytest.raises(RuntimeError):
        sparse_matrix_multiply(np.eye(4), np.eye(4), output_format='dense')


def test_product_package_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "sparse_matrix_mult_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f)).read()
                assert "oracle" not in text.replace("oracle as the per-rank compute", ""), f"{f} mentions the oracle"


def test_synthetic_generators_match_survey_counts():
    """Sizes SURVEY.md 8(d) measured for the same generators/seeds (cheap ones only)."""
    from sparse_matrix_mult_b200 import synthetic
    w = synthetic.workload("cfg1")
    a = w["a"]
    assert a.shape == (10_000, 10_000) and a.nnz == 100_000
    p = int(np.diff(a.indptr).astype(np.int64)[a.indices].sum())
    assert p == 1_000_408
    r = synthetic.rmat(10)
    assert r.shape == (1024, 1024) and r.has_sorted_indices and r.nnz <= 16 * 1024
    q = synthetic.banded_cov(1000)
    assert q.nnz == sum(1000 - abs(o) for o in range(-32, 33))
