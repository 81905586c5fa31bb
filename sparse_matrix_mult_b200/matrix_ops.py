"""Host side of the B200 SpGEMM path: the reference's Python API over libspgemm_b200.so.

Mirror of /root/reference/sparse_matrix_mult/matrix_ops.py.  Same public function, same argument meaning,
same pre-dispatch ValueErrors and short-circuits; the native library is loaded with ctypes.CDLL from
<package>/lib/ the way MatrixOpsLibrary does (matrix_ops.py:51-181) -- but it is a CUDA library with a
pointer-and-size ABI (include/spgemm_b200.h) instead of struct pointers, and it has no CPU fallback:
import works anywhere, the first compute call raises RuntimeError when no B200 is present.

Deliberate deviations from the reference (all documented in DESIGN.md):
  * no stdout chatter at import or on all-zero results except the one "zero matrix" notice when no
    product was formed (the reference scans the whole dense result for it, matrix_ops.py:370);
  * CUDA / allocation failures raise RuntimeError instead of being swallowed into a zero matrix
    (matrix_ops.py:377-387); an invalid `output_format` still prints and returns zeros as the reference does;
  * sparse results have sorted column indices (the reference's are in first-touch order);
  * `mirror=True` (new keyword, default off) fills the lower triangle of symmetric dense results on the GPU;
  * `n_gpus=N` (new keyword; default: the SPGEMM_NUM_GPUS environment variable, else 1) shards the rows of the
    product over N GPUs of the box inside the call -- one host thread per GPU in the library, the way the
    reference sizes its OpenMP team inside the call (src/sparse_sparse_sparse.cpp:188-197).

Thread safety: the library serialises concurrent calls per device (include/spgemm_b200.h, "Thread safety"), so
sparse_matrix_multiply may be called from several Python threads.
"""
import ctypes
import os
import threading

import numpy as np
from scipy.sparse import csr_matrix, isspmatrix_csr

_i32p = ctypes.POINTER(ctypes.c_int32)
_f64p = ctypes.POINTER(ctypes.c_double)
_vp = ctypes.c_void_p

TRIPLE_UPPER, TRIPLE_REF_FULL, TRIPLE_MIRROR = 0, 1, 2


class Stats(ctypes.Structure):
    """spgemm_b200_stats (include/spgemm_b200.h)."""
    _fields_ = [("ms_h2d", ctypes.c_double), ("ms_analysis", ctypes.c_double), ("ms_symbolic", ctypes.c_double),
                ("ms_numeric", ctypes.c_double), ("ms_post", ctypes.c_double), ("ms_d2h", ctypes.c_double),
                ("ms_total", ctypes.c_double), ("products", ctypes.c_int64), ("nnz_c", ctypes.c_int64),
                ("bytes_min", ctypes.c_int64), ("launches", ctypes.c_int32), ("device", ctypes.c_int32),
                ("bytes_h2d", ctypes.c_int64), ("bytes_d2h", ctypes.c_int64)]

    def as_dict(self):
        return {name: getattr(self, name) for name, _ in self._fields_}


class MatrixOpsLibrary:
    """Singleton loader of lib/libspgemm_b200.so (reference: MatrixOpsLibrary, matrix_ops.py:51-181)."""
    _instance = None
    _lock = threading.Lock()
    LIB_NAME = "libspgemm_b200.so"     # must not match the reference's libsparse*.so pattern (matrix_ops.py:118)

    def __new__(cls):
        with cls._lock:
            if cls._instance is None:
                inst = super().__new__(cls)
                inst._lib = None
                cls._instance = inst
        return cls._instance

    @property
    def lib_path(self):
        return os.path.join(os.path.dirname(os.path.abspath(__file__)), 'lib', self.LIB_NAME)

    def _load_library(self):
        path = self.lib_path
        if not os.path.exists(path):
            raise OSError(f"{path} not found: build it with `make -C {os.path.dirname(os.path.dirname(path))}` "
                          f"(nvcc, sm_100a); there is no CPU fallback")
        try:
            self._lib = ctypes.CDLL(path)
        except OSError as e:
            raise OSError(f"Failed to load library: {path}. Error: {e}")
        self._setup_function_prototypes()

    def _setup_function_prototypes(self):
        L = self._lib
        csr = [_i32p, _i32p, _f64p]
        L.spgemm_b200_version.restype = ctypes.c_char_p
        L.spgemm_b200_last_error.restype = ctypes.c_char_p
        L.spgemm_b200_device_count.restype = ctypes.c_int
        L.spgemm_b200_init.argtypes = [ctypes.c_int]
        L.spgemm_b200_shutdown.restype = None
        L.spgemm_b200_get_stats.argtypes = [ctypes.POINTER(Stats)]
        L.spgemm_b200_host_alloc.argtypes = [ctypes.c_size_t]
        L.spgemm_b200_host_alloc.restype = _vp
        L.spgemm_b200_host_free.argtypes = [_vp]
        L.spgemm_b200_host_free.restype = None
        L.spgemm_b200_csr.argtypes = [ctypes.c_int] * 3 + csr + csr + [ctypes.c_int, ctypes.POINTER(_vp)]
        L.spgemm_b200_result_nnz.argtypes = [_vp]
        L.spgemm_b200_result_nnz.restype = ctypes.c_int64
        L.spgemm_b200_result_rows.argtypes = [_vp]
        L.spgemm_b200_result_cols.argtypes = [_vp]
        L.spgemm_b200_result_copy.argtypes = [_vp, _vp, ctypes.c_int, _i32p, _f64p]
        L.spgemm_b200_result_free.argtypes = [_vp]
        L.spgemm_b200_result_free.restype = None
        L.spgemm_b200_result_device_ptrs.argtypes = [_vp, ctypes.POINTER(_vp), ctypes.POINTER(_vp), ctypes.POINTER(_vp)]
        L.spgemm_b200_dense.argtypes = [ctypes.c_int] * 3 + csr + csr + [ctypes.c_int, ctypes.c_int, _f64p]
        L.spgemm_b200_triple.argtypes = [ctypes.c_int] * 2 + csr + csr + [ctypes.c_int, _f64p]
        L.spgemm_b200_mat_upload.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64] + csr + [ctypes.POINTER(_vp)]
        L.spgemm_b200_mat_wrap.argtypes = [ctypes.c_int, ctypes.c_int, ctypes.c_int64, _vp, _vp, _vp, ctypes.POINTER(_vp)]
        L.spgemm_b200_mat_transpose.argtypes = [_vp, ctypes.POINTER(_vp)]
        L.spgemm_b200_mat_free.argtypes = [_vp]
        L.spgemm_b200_mat_free.restype = None
        L.spgemm_b200_csr_dev.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.POINTER(_vp)]
        L.spgemm_b200_dense_dev.argtypes = [_vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]
        L.spgemm_b200_triple_dev.argtypes = [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, ctypes.c_int, _vp]
        L.spgemm_b200_mirror_dev.argtypes = [_vp, ctypes.c_int]
        L.spgemm_b200_symmetrize_dev.argtypes = [_vp, ctypes.c_int]
        L.spgemm_b200_row_costs.argtypes = [_vp, _vp, _vp, ctypes.c_int, ctypes.c_int, _vp, ctypes.POINTER(ctypes.c_int64)]
        L.spgemm_b200_partition.argtypes = [_vp, ctypes.c_int, ctypes.c_int, _i32p]
        L.spgemm_b200_partition_tail.argtypes = [_vp, _i32p, ctypes.c_double, ctypes.c_int, ctypes.c_int, _i32p]
        L.spgemm_b200_device_alloc.argtypes = [ctypes.c_size_t]
        L.spgemm_b200_device_alloc.restype = _vp
        L.spgemm_b200_device_free.argtypes = [_vp]
        L.spgemm_b200_device_free.restype = None
        L.spgemm_b200_copy_to_host.argtypes = [_vp, _vp, ctypes.c_size_t]
        L.spgemm_b200_copy_to_device.argtypes = [_vp, _vp, ctypes.c_size_t]
        L.spgemm_b200_copy_on_device.argtypes = [_vp, _vp, ctypes.c_size_t]
        L.spgemm_b200_copy_upper_to_host.argtypes = [_vp, _vp, ctypes.c_int]
        L.spgemm_b200_shared_alloc.argtypes = [ctypes.c_size_t]
        L.spgemm_b200_shared_alloc.restype = _vp
        L.spgemm_b200_shared_free.argtypes = [_vp]
        L.spgemm_b200_shared_free.restype = None
        L.spgemm_b200_ipc_export.argtypes = [_vp, ctypes.c_char_p]
        L.spgemm_b200_ipc_open.argtypes = [ctypes.c_char_p, ctypes.POINTER(_vp)]
        L.spgemm_b200_ipc_close.argtypes = [_vp]
        L.spgemm_b200_set_stream.argtypes = [_vp]
        L.spgemm_b200_timer_stop.argtypes = [ctypes.POINTER(ctypes.c_double)]
        L.spgemm_b200_trim.argtypes = [ctypes.c_size_t]
        L.spgemm_b200_mat_cache_transpose.argtypes = [_vp, ctypes.c_int]
        L.spgemm_b200_mat_sort.argtypes = [_vp]
        L.spgemm_b200_mat_is_sorted.argtypes = [_vp]
        L.spgemm_b200_multi_dense.argtypes = [ctypes.c_int] * 4 + csr + csr + [ctypes.c_int, _f64p]
        L.spgemm_b200_multi_triple.argtypes = [ctypes.c_int] * 3 + csr + csr + [_f64p]
        L.spgemm_b200_multi_csr.argtypes = [ctypes.c_int] * 4 + csr + csr + [ctypes.c_int, ctypes.POINTER(_vp)]
        L.spgemm_b200_multi_result_nnz.argtypes = [_vp]
        L.spgemm_b200_multi_result_nnz.restype = ctypes.c_int64
        L.spgemm_b200_multi_result_copy.argtypes = [_vp, _vp, ctypes.c_int, _i32p, _f64p]
        L.spgemm_b200_multi_result_free.argtypes = [_vp]
        L.spgemm_b200_multi_result_free.restype = None
        L.spgemm_b200_multi_last_bounds.argtypes = [_i32p, ctypes.c_int]
        L.spgemm_b200_multi_last_stats.argtypes = [ctypes.c_int, ctypes.POINTER(Stats)]

    def get_lib(self):
        if self._lib is None:
            with self._lock:
                if self._lib is None:
                    self._load_library()
        return self._lib


matrix_ops = MatrixOpsLibrary()      # reference: module-level singleton, matrix_ops.py:184 (loading is lazy here)


def _check(rc, what):
    if rc != 0:
        msg = matrix_ops.get_lib().spgemm_b200_last_error().decode(errors="replace")
        raise RuntimeError(f"{what} failed (status {rc}): {msg}")


def last_stats():
    """Timing/size record of the most recent library call as a dict (see include/spgemm_b200.h)."""
    s = Stats()
    _check(matrix_ops.get_lib().spgemm_b200_get_stats(ctypes.byref(s)), "spgemm_b200_get_stats")
    return s.as_dict()


# ------------------------------------------------------------------------------------------------------
# result storage: pinned host memory owned by the library's cache, exposed as ordinary ndarrays
_PINNED_MIN_BYTES = 1 << 16


class _PinnedOwner:
    __slots__ = ("ptr", "_free")

    def __init__(self, lib, nbytes):
        self.ptr = lib.spgemm_b200_host_alloc(nbytes)
        self._free = lib.spgemm_b200_host_free
        if not self.ptr:
            raise MemoryError(f"pinned allocation of {nbytes} bytes failed: "
                              f"{lib.spgemm_b200_last_error().decode(errors='replace')}")

    def __del__(self):
        if self.ptr:
            self._free(self.ptr)
            self.ptr = None


def _result_array(shape, dtype):
    """np.empty(shape, dtype) whose storage is page-locked (so the copy out of HBM is one DMA) and returns
    to the library's cache when the array is garbage collected.  Replaces the calloc'd C array +
    `.copy()` of darray_to_numpy / sparsemat_to_csr (matrix_ops.py:205-240)."""
    dtype = np.dtype(dtype)
    count = int(np.prod(shape, dtype=np.int64))
    nbytes = count * dtype.itemsize
    if nbytes < _PINNED_MIN_BYTES:
        return np.empty(shape, dtype=dtype)
    owner = _PinnedOwner(matrix_ops.get_lib(), nbytes)
    buf = (ctypes.c_char * nbytes).from_address(owner.ptr)
    buf._owner = owner                       # lifetime: ndarray -> buf -> owner -> host_free
    return np.frombuffer(buf, dtype=dtype, count=count).reshape(shape)


def csr_to_arrays(csr):
    """(indptr int32, indices int32, data float64) contiguous views/copies of a scipy CSR, cast exactly like
    csr_to_sparsemat (matrix_ops.py:187-202); no canonicalisation."""
    if csr.nnz >= 2 ** 31 or max(csr.shape) >= 2 ** 31:
        raise OverflowError("operands with nnz or a dimension >= 2**31 do not fit the int32 CSR of this interface "
                            "(the reference has the same limit: include/matrix_def.h:21-22, matrix_ops.py:196-197)")
    return (np.ascontiguousarray(csr.indptr, dtype=np.int32),
            np.ascontiguousarray(csr.indices, dtype=np.int32),
            np.ascontiguousarray(csr.data, dtype=np.float64))


def _ptrs(arrs):
    p, i, v = arrs
    return p.ctypes.data_as(_i32p), i.ctypes.data_as(_i32p), v.ctypes.data_as(_f64p)


def result_to_csr(lib, handle, shape, multi=False):
    """Device result -> scipy CSR (reference: sparsemat_to_csr, matrix_ops.py:205-228).  multi: the handle is a
    spgemm_b200_multi_result (row blocks on several GPUs, copied out in parallel)."""
    nnz = (lib.spgemm_b200_multi_result_nnz if multi else lib.spgemm_b200_result_nnz)(handle)
    if nnz == 0:
        return csr_matrix(shape)
    if nnz >= 2 ** 31:
        raise OverflowError(f"nnz(C) = {nnz} does not fit SciPy's int32 CSR; use the device-resident API "
                            f"(sparse_matrix_mult_b200.device) for products of this size")
    indptr = _result_array((shape[0] + 1,), np.int32)
    indices = _result_array((nnz,), np.int32)
    data = _result_array((nnz,), np.float64)
    copy = lib.spgemm_b200_multi_result_copy if multi else lib.spgemm_b200_result_copy
    _check(copy(handle, indptr.ctypes.data_as(_vp), 0, indices.ctypes.data_as(_i32p), data.ctypes.data_as(_f64p)),
           "spgemm_b200_result_copy")
    out = csr_matrix((data, indices, indptr), shape=shape, copy=False)
    out.has_sorted_indices = True
    return out


def multi_last_bounds():
    """Row bounds (len n_gpus + 1) the last n_gpus > 1 call sharded by."""
    buf = (ctypes.c_int32 * 64)()
    n = matrix_ops.get_lib().spgemm_b200_multi_last_bounds(buf, 64)
    return [int(buf[i]) for i in range(min(n, 64))]


def multi_last_stats():
    """Per-GPU stats dicts of the last n_gpus > 1 call."""
    lib, out = matrix_ops.get_lib(), []
    for part in range(max(0, len(multi_last_bounds()) - 1)):
        s = Stats()
        _check(lib.spgemm_b200_multi_last_stats(part, ctypes.byref(s)), "spgemm_b200_multi_last_stats")
        out.append(s.as_dict())
    return out


def _default_gpus():
    try:
        return max(1, int(os.environ.get("SPGEMM_NUM_GPUS", "1")))
    except ValueError:
        return 1


def sparse_matrix_multiply(matrix_a, matrix_b, output_format='sparse', symmetric=False, imem_size=None,
                           use_triple_product=False, compute_full_matrix=None, mirror=False, n_gpus=None):
    """Multiply two sparse matrices on a B200.  Signature and semantics of
    /root/reference/sparse_matrix_mult/matrix_ops.py:271-387.

    matrix_a, matrix_b : scipy CSR (used as is) or anything csr_matrix() accepts.
    output_format      : 'sparse' -> scipy.sparse.csr_matrix, 'dense' -> C-contiguous float64 ndarray.
    symmetric          : keep only col >= row (zeros below the diagonal), result must be square.
    imem_size          : accepted and validated like the reference; unused (the symbolic phase sizes C exactly).
    use_triple_product : return matrix_a @ matrix_b @ matrix_a.T as a dense ndarray (wins over output_format).
    compute_full_matrix: triple product only.  None/0 -> upper triangle; 1 -> what the reference returns for 1,
                         i.e. T + T.T - diag(T) of the full product T.
    mirror             : extension.  With symmetric=True, output_format='dense' (or the triple product with
                         compute_full_matrix in (None, 0)) also fill the lower triangle with the mirror image.
    n_gpus             : extension.  Shard the rows of the product over this many GPUs of the box inside the call
                         (default: $SPGEMM_NUM_GPUS, else 1).  The modes that need the whole matrix on one device
                         (mirror=True, compute_full_matrix=1) always run on one GPU.
    """
    # -- argument handling: matrix_ops.py:288-305 ----------------------------------------------------
    if imem_size is None:
        imem_size = 5
    else:
        try:
            imem_size = int(imem_size)
        except ValueError:
            raise ValueError(f"imem_size must be an integer or None, got {type(imem_size)}")
    if compute_full_matrix is None:
        compute_full_matrix = 0
    else:
        if compute_full_matrix not in (0, 1):
            raise ValueError("compute_full_matrix must be None, 0, or 1")
        compute_full_matrix = int(compute_full_matrix)

    # -- coercion, shape checks, short-circuits: matrix_ops.py:307-322 ---------------------------------
    if not isspmatrix_csr(matrix_a):
        matrix_a = csr_matrix(matrix_a)
    if not isspmatrix_csr(matrix_b):
        matrix_b = csr_matrix(matrix_b)
    if matrix_a.shape[1] != matrix_b.shape[0]:
        raise ValueError("Matrix dimensions are incompatible for multiplication.")
    if matrix_a.nnz == 0 or matrix_b.nnz == 0:
        if output_format == 'sparse':
            return csr_matrix((matrix_a.shape[0], matrix_b.shape[1]))
        else:
            return np.zeros((matrix_a.shape[0], matrix_b.shape[1]))
    if symmetric and (matrix_a.shape[0] != matrix_b.shape[1]):
        raise ValueError("For symmetric output, the resulting matrix must be square.")

    # -- dispatch: matrix_ops.py:324-368; precedence triple > sparse > dense -----------------------------
    m, k = matrix_a.shape
    n = matrix_b.shape[1]
    if not use_triple_product and output_format not in ('sparse', 'dense'):
        # the reference raises inside its try block and swallows it (matrix_ops.py:367-387)
        print("An error occurred during matrix multiplication: Invalid output_format. Choose 'sparse' or 'dense'.")
        return np.zeros((m, n))

    lib = matrix_ops.get_lib()
    a_arr, b_arr = csr_to_arrays(matrix_a), csr_to_arrays(matrix_b)
    n_gpus = _default_gpus() if n_gpus is None else int(n_gpus)
    if n_gpus < 1:
        raise ValueError("n_gpus must be a positive integer")
    force_multi = bool(os.environ.get("SPGEMM_B200_FORCE_MULTI"))        # tests: multi-GPU driver on a one-GPU box
    if (n_gpus > 1 or force_multi) and not (mirror or (use_triple_product and compute_full_matrix)):
        return _multiply_multi(lib, n_gpus, a_arr, b_arr, (m, k, n), output_format, symmetric, use_triple_product)
    if use_triple_product:
        # like the reference (matrix_ops.py:312-313 is the only check) Q is assumed square with H.cols rows;
        # unlike it, a violation is reported instead of reading out of bounds
        if matrix_b.shape[0] != matrix_b.shape[1]:
            raise ValueError("Triple product needs a square second matrix (H @ Q @ H.T).")
        mode = TRIPLE_REF_FULL if compute_full_matrix else (TRIPLE_MIRROR if mirror else TRIPLE_UPPER)
        result = _result_array((m, m), np.float64)
        _check(lib.spgemm_b200_triple(m, k, *_ptrs(a_arr), *_ptrs(b_arr), mode, result.ctypes.data_as(_f64p)),
               "spgemm_b200_triple")
    elif output_format == 'sparse':
        handle = _vp()
        _check(lib.spgemm_b200_csr(m, k, n, *_ptrs(a_arr), *_ptrs(b_arr), 1 if symmetric else 0,
                                   ctypes.byref(handle)), "spgemm_b200_csr")
        try:
            result = result_to_csr(lib, handle, (m, n))
        finally:
            lib.spgemm_b200_result_free(handle)
    else:
        result = _result_array((m, n), np.float64)
        _check(lib.spgemm_b200_dense(m, k, n, *_ptrs(a_arr), *_ptrs(b_arr), 1 if symmetric else 0,
                                     1 if (mirror and symmetric) else 0, result.ctypes.data_as(_f64p)),
               "spgemm_b200_dense")

    if isinstance(result, csr_matrix) and result.nnz == 0:
        print("Multiplication resulted in a zero matrix.")
    return result


def _multiply_multi(lib, n_gpus, a_arr, b_arr, dims, output_format, symmetric, use_triple_product):
    """The n_gpus > 1 legs of sparse_matrix_multiply: same dispatch precedence, multi-GPU entry points."""
    m, k, n = dims
    if use_triple_product:
        if k != n:
            raise ValueError("Triple product needs a square second matrix (H @ Q @ H.T).")
        result = _result_array((m, m), np.float64)
        _check(lib.spgemm_b200_multi_triple(n_gpus, m, k, *_ptrs(a_arr), *_ptrs(b_arr), result.ctypes.data_as(_f64p)),
               "spgemm_b200_multi_triple")
        return result
    if output_format == 'sparse':
        handle = _vp()
        _check(lib.spgemm_b200_multi_csr(n_gpus, m, k, n, *_ptrs(a_arr), *_ptrs(b_arr), 1 if symmetric else 0,
                                         ctypes.byref(handle)), "spgemm_b200_multi_csr")
        try:
            result = result_to_csr(lib, handle, (m, n), multi=True)
        finally:
            lib.spgemm_b200_multi_result_free(handle)
        if result.nnz == 0:
            print("Multiplication resulted in a zero matrix.")
        return result
    result = _result_array((m, n), np.float64)
    _check(lib.spgemm_b200_multi_dense(n_gpus, m, k, n, *_ptrs(a_arr), *_ptrs(b_arr), 1 if symmetric else 0,
                                       result.ctypes.data_as(_f64p)), "spgemm_b200_multi_dense")
    return result
