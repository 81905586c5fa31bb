"""ctypes view of liboracle.so (oracle/spgemm_oracle.c).  TEST INFRASTRUCTURE ONLY.

Each function mirrors one reference entry point; see the C file for file:line citations.
Inputs are scipy CSR matrices (used as they are: no sort, no duplicate merge -- like
/root/reference/sparse_matrix_mult/matrix_ops.py:187-202, which only casts dtypes).
"""
import ctypes
import os
import subprocess

import numpy as np
from scipy.sparse import csr_matrix

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "liboracle.so")
_lib = None

_i32p = ctypes.POINTER(ctypes.c_int32)
_i64p = ctypes.POINTER(ctypes.c_int64)
_f64p = ctypes.POINTER(ctypes.c_double)


class _OracleCsr(ctypes.Structure):
    _fields_ = [("nnz", ctypes.c_int64), ("rows", ctypes.c_int64),
                ("indptr", _i64p), ("indices", _i32p), ("values", _f64p)]


def build(force=False):
    """Compile liboracle.so with the Makefile next to this file (gcc only, no GPU needed)."""
    if force or not os.path.exists(_LIB_PATH) or \
            os.path.getmtime(_LIB_PATH) < os.path.getmtime(os.path.join(_HERE, "spgemm_oracle.c")):
        subprocess.check_call(["make", "-s", "-C", _HERE, "liboracle.so"])


def lib():
    global _lib
    if _lib is None:
        build()
        L = ctypes.CDLL(_LIB_PATH)
        csr_args = [ctypes.c_int] * 3 + [_i32p, _i32p, _f64p] * 2
        L.oracle_spgemm_csr.argtypes = csr_args + [ctypes.c_int]
        L.oracle_spgemm_csr.restype = ctypes.POINTER(_OracleCsr)
        L.oracle_spgemm_csr_omp.argtypes = csr_args + [ctypes.c_int, ctypes.c_int]
        L.oracle_spgemm_csr_omp.restype = ctypes.POINTER(_OracleCsr)
        L.oracle_csr_free.argtypes = [ctypes.POINTER(_OracleCsr)]
        L.oracle_csr_free.restype = None
        L.oracle_spgemm_dense.argtypes = csr_args + [ctypes.c_int, _f64p]
        L.oracle_spgemm_dense.restype = None
        L.oracle_triple_product.argtypes = [ctypes.c_int] * 2 + [_i32p, _i32p, _f64p] * 2 + [ctypes.c_int, _f64p]
        L.oracle_triple_product.restype = None
        L.oracle_limits.argtypes = [ctypes.c_int, ctypes.c_int, _i32p]
        L.oracle_limits.restype = ctypes.c_int
        L.oracle_count_products.argtypes = [ctypes.c_int, _i32p, _i32p, _i32p]
        L.oracle_count_products.restype = ctypes.c_int64
        _lib = L
    return _lib


def _csr_arrays(x):
    if not isinstance(x, csr_matrix):
        x = csr_matrix(x)
    ptr = np.ascontiguousarray(x.indptr, dtype=np.int32)
    idx = np.ascontiguousarray(x.indices, dtype=np.int32)
    val = np.ascontiguousarray(x.data, dtype=np.float64)
    return x.shape, ptr, idx, val


def _p(a, t):
    return a.ctypes.data_as(t)


def spgemm_csr(a, b, upper_only=False, omp_blocks=0, copy=True):
    """A*B as csr_matrix, columns in first-touch order, int64 indptr when nnz >= 2**31 else int32.
    omp_blocks > 0 runs the multi-threaded variant (OpenMP, that many row blocks, dynamic schedule) -- the
    many-core CPU baseline of bench.py; the result is bit-identical.  copy=False only times the C call."""
    (m, k), ap, ai, av = _csr_arrays(a)
    (k2, n), bp, bi, bv = _csr_arrays(b)
    assert k == k2
    L = lib()
    if omp_blocks > 0:
        h = L.oracle_spgemm_csr_omp(m, k, n, _p(ap, _i32p), _p(ai, _i32p), _p(av, _f64p),
                                    _p(bp, _i32p), _p(bi, _i32p), _p(bv, _f64p), int(bool(upper_only)),
                                    int(omp_blocks))
    else:
        h = L.oracle_spgemm_csr(m, k, n, _p(ap, _i32p), _p(ai, _i32p), _p(av, _f64p),
                                _p(bp, _i32p), _p(bi, _i32p), _p(bv, _f64p), int(bool(upper_only)))
    if h and not copy:
        L.oracle_csr_free(h)
        return None
    if not h:
        raise MemoryError("oracle_spgemm_csr")
    try:
        nnz = h.contents.nnz
        indptr = np.ctypeslib.as_array(h.contents.indptr, shape=(m + 1,)).copy()
        if nnz:
            indices = np.ctypeslib.as_array(h.contents.indices, shape=(nnz,)).copy()
            values = np.ctypeslib.as_array(h.contents.values, shape=(nnz,)).copy()
        else:
            indices = np.zeros(0, np.int32)
            values = np.zeros(0, np.float64)
    finally:
        L.oracle_csr_free(h)
    if nnz < 2 ** 31:
        indptr = indptr.astype(np.int32)
    out = csr_matrix((m, n))
    out.indptr, out.indices, out.data = indptr, indices, values
    return out


def spgemm_dense(a, b, upper_only=False):
    (m, k), ap, ai, av = _csr_arrays(a)
    (k2, n), bp, bi, bv = _csr_arrays(b)
    assert k == k2
    c = np.empty((m, n), dtype=np.float64)
    lib().oracle_spgemm_dense(m, k, n, _p(ap, _i32p), _p(ai, _i32p), _p(av, _f64p),
                              _p(bp, _i32p), _p(bi, _i32p), _p(bv, _f64p), int(bool(upper_only)), _p(c, _f64p))
    return c


def triple_product(h, q, full=0):
    (n, k), hp, hi, hv = _csr_arrays(h)
    (k2, k3), qp, qi, qv = _csr_arrays(q)
    assert k == k2 == k3, "triple product needs square Q with H.cols rows"
    c = np.empty((n, n), dtype=np.float64)
    lib().oracle_triple_product(n, k, _p(hp, _i32p), _p(hi, _i32p), _p(hv, _f64p),
                                _p(qp, _i32p), _p(qi, _i32p), _p(qv, _f64p), int(full), _p(c, _f64p))
    return c


def limits(rows, parts):
    """[(first, last_inclusive)] per partition, as src/workdivision.cpp:16-89 lays them out."""
    out = np.zeros(2 * max(1, min(parts, rows)), dtype=np.int32)
    got = lib().oracle_limits(rows, parts, _p(out, _i32p))
    return [(int(out[p]), int(out[p + got])) for p in range(got)]


def count_products(a, b):
    (m, _), ap, ai, _ = _csr_arrays(a)
    _, bp, _, _ = _csr_arrays(b)
    return int(lib().oracle_count_products(m, _p(ap, _i32p), _p(ai, _i32p), _p(bp, _i32p)))


def sparse_matrix_multiply(matrix_a, matrix_b, output_format='sparse', symmetric=False, imem_size=None,
                           use_triple_product=False, compute_full_matrix=None):
    """The reference's dispatch (matrix_ops.py:271-387) over the C restatement -- checker only."""
    full = 0 if compute_full_matrix is None else int(compute_full_matrix)
    a = matrix_a if isinstance(matrix_a, csr_matrix) else csr_matrix(matrix_a)
    b = matrix_b if isinstance(matrix_b, csr_matrix) else csr_matrix(matrix_b)
    if a.shape[1] != b.shape[0]:
        raise ValueError("Matrix dimensions are incompatible for multiplication.")
    if a.nnz == 0 or b.nnz == 0:
        return csr_matrix((a.shape[0], b.shape[1])) if output_format == 'sparse' \
            else np.zeros((a.shape[0], b.shape[1]))
    if symmetric and a.shape[0] != b.shape[1]:
        raise ValueError("For symmetric output, the resulting matrix must be square.")
    if use_triple_product:
        return triple_product(a, b, full)
    if output_format == 'sparse':
        return spgemm_csr(a, b, symmetric)
    if output_format == 'dense':
        return spgemm_dense(a, b, symmetric)
    raise ValueError("Invalid output_format. Choose 'sparse' or 'dense'.")
