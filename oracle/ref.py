"""ctypes view of the reference's own binaries in oracle/_ref/.  TEST INFRASTRUCTURE ONLY.

Two libraries, built by `make -C oracle ref` (needs /root/reference; the GPU box only ever sees the
prebuilt files):

  libsparse_ref_shipped.so  the reference's prebuilt Linux binary: serial, `int` struct fields
                            (matches matrix_ops.py:26-33,44-48).  All five entry points work; this is the
                            authoritative oracle for the sparse-output path.
  libsparse_ref_omp.so      today's src/*.cpp with setup.py's flags (+OpenMP): `size_t` struct fields
                            (include/matrix_def.h:17-31).  dense_nosym / dense_sym / triple_product work
                            and are bit-identical to the shipped binary; sparse_* are defective.

  libsparse_ref_omp_patched.so  the same sources with the repairs of SURVEY.md Appendix B applied in memory at
                            build time (oracle/build_patched_ref.py): the reference's multi-threaded sparse_nosym /
                            sparse_sym with a working body -- bit-identical to the shipped binary
                            (tests/test_oracle.py), used as the many-core CPU baseline of the sparse-output path.

Raw C calls only (no reference Python on the path) so that timing measures the C routine itself.
"""
import ctypes
import os

import numpy as np
from scipy.sparse import csr_matrix

_HERE = os.path.dirname(os.path.abspath(__file__))
_DIR = os.path.join(_HERE, "_ref")
_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)


def _structs(size_t_fields):
    T = ctypes.c_size_t if size_t_fields else ctypes.c_int

    class SparseMat(ctypes.Structure):
        _fields_ = [("nzmax", T), ("rows", T), ("cols", T),
                    ("rowPtr", _i32p), ("colInd", _i32p), ("values", _f64p)]

    class DArray(ctypes.Structure):
        _fields_ = [("array", _f64p), ("rows", T), ("cols", T)]

    return SparseMat, DArray


class RefLib:
    def __init__(self, name, size_t_fields):
        self.path = os.path.join(_DIR, name)
        self.lib = ctypes.CDLL(self.path)
        self.SparseMat, self.DArray = _structs(size_t_fields)
        self.lib.destroy_darray.argtypes = [ctypes.POINTER(self.DArray)]
        self.lib.destroy_sparsemat.argtypes = [ctypes.POINTER(self.SparseMat)]
        for f in ("dense_nosym", "dense_sym"):
            getattr(self.lib, f).argtypes = [ctypes.POINTER(self.SparseMat)] * 2 + [ctypes.POINTER(self.DArray)]
            getattr(self.lib, f).restype = None
        for f in ("sparse_nosym", "sparse_sym"):
            getattr(self.lib, f).argtypes = [ctypes.POINTER(self.SparseMat)] * 3 + [ctypes.c_int]
            getattr(self.lib, f).restype = None
        self.lib.triple_product.argtypes = [ctypes.POINTER(self.SparseMat)] * 2 + \
            [ctypes.POINTER(self.DArray), ctypes.c_int]
        self.lib.triple_product.restype = None

    def _mat(self, x):
        """Borrow numpy buffers (kept alive on the returned struct) instead of create_sparsemat+memmove."""
        if not isinstance(x, csr_matrix):
            x = csr_matrix(x)
        ptr = np.ascontiguousarray(x.indptr, dtype=np.int32)
        idx = np.ascontiguousarray(x.indices, dtype=np.int32)
        val = np.ascontiguousarray(x.data, dtype=np.float64)
        s = self.SparseMat(x.nnz, x.shape[0], x.shape[1], ptr.ctypes.data_as(_i32p),
                           idx.ctypes.data_as(_i32p), val.ctypes.data_as(_f64p))
        s._keep = (ptr, idx, val)
        return s

    def dense(self, a, b, upper_only=False, copy=True):
        sa, sb, out = self._mat(a), self._mat(b), self.DArray()
        (self.lib.dense_sym if upper_only else self.lib.dense_nosym)(ctypes.byref(sa), ctypes.byref(sb),
                                                                     ctypes.byref(out))
        res = None
        if copy:
            res = np.ctypeslib.as_array(out.array, shape=(int(out.rows), int(out.cols))).copy()
        self.lib.destroy_darray(ctypes.byref(out))
        return res

    def triple(self, h, q, full=0, copy=True):
        sh, sq, out = self._mat(h), self._mat(q), self.DArray()
        self.lib.triple_product(ctypes.byref(sh), ctypes.byref(sq), ctypes.byref(out), int(full))
        res = None
        if copy:
            res = np.ctypeslib.as_array(out.array, shape=(int(out.rows), int(out.cols))).copy()
        self.lib.destroy_darray(ctypes.byref(out))
        return res

    def sparse(self, a, b, upper_only=False, imem_size=5, copy=True):
        sa, sb, out = self._mat(a), self._mat(b), self.SparseMat()
        (self.lib.sparse_sym if upper_only else self.lib.sparse_nosym)(ctypes.byref(sa), ctypes.byref(sb),
                                                                       ctypes.byref(out), int(imem_size))
        m, n, nnz = int(out.rows), int(out.cols), int(out.nzmax)
        res = None
        if copy:
            if nnz == 0:
                res = csr_matrix((m, n))
            else:
                res = csr_matrix((m, n))
                res.indptr = np.ctypeslib.as_array(out.rowPtr, shape=(m + 1,)).copy()
                res.indices = np.ctypeslib.as_array(out.colInd, shape=(nnz,)).copy()
                res.data = np.ctypeslib.as_array(out.values, shape=(nnz,)).copy()
        if nnz:
            self.lib.destroy_sparsemat(ctypes.byref(out))
        return res


_cache = {}


def available():
    return all(os.path.exists(os.path.join(_DIR, n))
               for n in ("libsparse_ref_shipped.so", "libsparse_ref_omp.so"))


def shipped():
    """Serial prebuilt binary; valid for every mode."""
    if "shipped" not in _cache:
        _cache["shipped"] = RefLib("libsparse_ref_shipped.so", size_t_fields=False)
    return _cache["shipped"]


def patched_available():
    return os.path.exists(os.path.join(_DIR, "libsparse_ref_omp_patched.so"))


def omp_patched():
    """Appendix-B-repaired OpenMP build: every mode usable, multi-threaded."""
    if "patched" not in _cache:
        _cache["patched"] = RefLib("libsparse_ref_omp_patched.so", size_t_fields=True)
    return _cache["patched"]


def omp():
    """From-source OpenMP build; only .dense() and .triple() are usable."""
    if "omp" not in _cache:
        _cache["omp"] = RefLib("libsparse_ref_omp.so", size_t_fields=True)
    return _cache["omp"]
