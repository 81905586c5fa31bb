"""Device-resident handles over libspgemm_b200.so: operands stay in HBM across calls.

Used by bench.py's kernel-only leg, by the multi-GPU row-sharded driver (distributed.py) and by callers that
iterate on the same matrices.  Thin ctypes wrappers; every method maps to one entry point of
include/spgemm_b200.h.
"""
import ctypes

import numpy as np
from scipy.sparse import csr_matrix, isspmatrix_csr

from .matrix_ops import _check, _f64p, _i32p, _ptrs, _vp, csr_to_arrays, last_stats, matrix_ops, result_to_csr

__all__ = ["DeviceMatrix", "DeviceResult", "DeviceDense", "spgemm_csr", "spgemm_dense", "triple_product",
           "row_costs", "partition_rows", "last_stats", "set_stream", "synchronize", "init"]


def init(device=0):
    _check(matrix_ops.get_lib().spgemm_b200_init(int(device)), "spgemm_b200_init")


CUDA_STREAM_LEGACY = 1      # cudaStreamLegacy: the handle that names the legacy default stream explicitly


def set_stream(stream_ptr):
    """Launch on a caller stream; None restores the library's own stream.

    torch.cuda.current_stream().cuda_stream is 0 for torch's default stream -- the LEGACY default stream.  0 is also
    how the C ABI spells "use your own stream", so 0 is translated to cudaStreamLegacy here: the library then
    really runs on torch's stream and is ordered with the NCCL collectives torch issues around it."""
    if stream_ptr is None:
        handle = 0
    else:
        handle = int(stream_ptr) or CUDA_STREAM_LEGACY
    _check(matrix_ops.get_lib().spgemm_b200_set_stream(_vp(handle)), "spgemm_b200_set_stream")


def copy_on_device(dst_ptr, src_ptr, nbytes):
    """Device-to-device copy on the library stream (asynchronous)."""
    _check(matrix_ops.get_lib().spgemm_b200_copy_on_device(_vp(dst_ptr), _vp(src_ptr), int(nbytes)),
           "spgemm_b200_copy_on_device")


def copy_to_host(host_array, src_ptr):
    """Device -> host ndarray on the library stream; returns when the copy is done."""
    _check(matrix_ops.get_lib().spgemm_b200_copy_to_host(host_array.ctypes.data_as(_vp), _vp(src_ptr),
                                                         host_array.nbytes), "spgemm_b200_copy_to_host")


def copy_upper_to_host(host_array, src_ptr):
    """n x n device matrix with a zero strictly-lower triangle -> host ndarray, sending only the upper trapezoids."""
    n = host_array.shape[0]
    _check(matrix_ops.get_lib().spgemm_b200_copy_upper_to_host(host_array.ctypes.data_as(_vp), _vp(src_ptr), n),
           "spgemm_b200_copy_upper_to_host")


def synchronize():
    _check(matrix_ops.get_lib().spgemm_b200_synchronize(), "spgemm_b200_synchronize")


class DeviceMatrix:
    """A CSR matrix in HBM (int32 indices, float64 values)."""

    def __init__(self, handle, shape, nnz, keep=None):
        self._h, self.shape, self.nnz, self._keep = handle, tuple(shape), int(nnz), keep

    @classmethod
    def from_scipy(cls, x):
        if not isspmatrix_csr(x):
            x = csr_matrix(x)
        arrs = csr_to_arrays(x)
        h = _vp()
        _check(matrix_ops.get_lib().spgemm_b200_mat_upload(x.shape[0], x.shape[1], x.nnz, *_ptrs(arrs),
                                                           ctypes.byref(h)), "spgemm_b200_mat_upload")
        return cls(h, x.shape, x.nnz)

    @classmethod
    def wrap(cls, shape, nnz, indptr_ptr, indices_ptr, values_ptr, keep=None):
        """Borrow device arrays (raw addresses, e.g. tensor.data_ptr()); `keep` holds their owners alive."""
        h = _vp()
        _check(matrix_ops.get_lib().spgemm_b200_mat_wrap(shape[0], shape[1], nnz, _vp(indptr_ptr), _vp(indices_ptr),
                                                         _vp(values_ptr), ctypes.byref(h)), "spgemm_b200_mat_wrap")
        return cls(h, shape, nnz, keep)

    def transpose(self):
        """X^T on the device, rows sorted by descending column (what the triple product wants for H^T)."""
        h = _vp()
        _check(matrix_ops.get_lib().spgemm_b200_mat_transpose(self._h, ctypes.byref(h)), "spgemm_b200_mat_transpose")
        return DeviceMatrix(h, self.shape[::-1], self.nnz)

    def cache_transpose(self, enable=True):
        """Keep the paneled transpose triple_product() builds from this matrix (as H) across calls."""
        _check(matrix_ops.get_lib().spgemm_b200_mat_cache_transpose(self._h, int(bool(enable))),
               "spgemm_b200_mat_cache_transpose")

    def is_sorted(self):
        r = matrix_ops.get_lib().spgemm_b200_mat_is_sorted(self._h)
        if r < 0:
            _check(1, "spgemm_b200_mat_is_sorted")
        return bool(r)

    def sort_rows(self):
        """Device-side canonicalisation: rows sorted by ascending column (duplicates stay)."""
        _check(matrix_ops.get_lib().spgemm_b200_mat_sort(self._h), "spgemm_b200_mat_sort")

    def free(self):
        if self._h:
            matrix_ops.get_lib().spgemm_b200_mat_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceResult:
    """A sparse product in HBM: int64 indptr, int32 indices (sorted), float64 values."""

    def __init__(self, handle):
        lib = matrix_ops.get_lib()
        self._h = handle
        self.shape = (lib.spgemm_b200_result_rows(handle), lib.spgemm_b200_result_cols(handle))
        self.nnz = int(lib.spgemm_b200_result_nnz(handle))

    def device_ptrs(self):
        p, i, v = _vp(), _vp(), _vp()
        _check(matrix_ops.get_lib().spgemm_b200_result_device_ptrs(self._h, ctypes.byref(p), ctypes.byref(i),
                                                                   ctypes.byref(v)), "result_device_ptrs")
        return p.value, i.value, v.value

    def to_scipy(self):
        return result_to_csr(matrix_ops.get_lib(), self._h, self.shape)

    def indptr_host(self):
        out = np.empty(self.shape[0] + 1, dtype=np.int64)
        p, _, _ = self.device_ptrs()
        _check(matrix_ops.get_lib().spgemm_b200_copy_to_host(out.ctypes.data_as(_vp), _vp(p), out.nbytes), "copy_to_host")
        return out

    def free(self):
        if self._h:
            matrix_ops.get_lib().spgemm_b200_result_free(self._h)
            self._h = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class DeviceDense:
    """rows x cols float64 row-major buffer in HBM from the library's pool."""

    def __init__(self, rows, cols):
        self.shape = (int(rows), int(cols))
        self.nbytes = self.shape[0] * self.shape[1] * 8
        self.ptr = matrix_ops.get_lib().spgemm_b200_device_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(matrix_ops.get_lib().spgemm_b200_last_error().decode(errors="replace"))

    def to_host(self, out=None):
        if out is None:
            out = np.empty(self.shape, dtype=np.float64)
        _check(matrix_ops.get_lib().spgemm_b200_copy_to_host(out.ctypes.data_as(_vp), _vp(self.ptr), self.nbytes),
               "copy_to_host")
        return out

    def free(self):
        if self.ptr:
            matrix_ops.get_lib().spgemm_b200_device_free(_vp(self.ptr))
            self.ptr = None

    def __del__(self):
        try:
            self.free()
        except Exception:
            pass


class SharedDense:
    """rows x cols float64 buffer that other ranks on the box can map (CUDA IPC) and write into over NVLink."""

    def __init__(self, rows, cols, ptr=None, owner=True):
        lib = matrix_ops.get_lib()
        self.shape, self.nbytes, self.owner = (int(rows), int(cols)), int(rows) * int(cols) * 8, owner
        self.ptr = ptr if ptr is not None else lib.spgemm_b200_shared_alloc(self.nbytes)
        if not self.ptr:
            raise MemoryError(lib.spgemm_b200_last_error().decode(errors="replace"))

    def export(self):
        buf = ctypes.create_string_buffer(64)
        _check(matrix_ops.get_lib().spgemm_b200_ipc_export(_vp(self.ptr), buf), "spgemm_b200_ipc_export")
        return buf.raw

    @classmethod
    def open(cls, handle, rows, cols):
        p = _vp()
        _check(matrix_ops.get_lib().spgemm_b200_ipc_open(handle, ctypes.byref(p)), "spgemm_b200_ipc_open")
        return cls(rows, cols, ptr=p.value, owner=False)

    def row_ptr(self, r):
        return self.ptr + int(r) * self.shape[1] * 8

    def close(self):
        lib = matrix_ops.get_lib()
        if self.ptr:
            if self.owner:
                lib.spgemm_b200_shared_free(_vp(self.ptr))
            else:
                lib.spgemm_b200_ipc_close(_vp(self.ptr))
            self.ptr = None


def _rows(a, row_begin, row_end):
    if row_end is None:
        return int(row_begin), a.shape[0]
    return int(row_begin), int(row_end)


def spgemm_csr(a, b, upper_only=False, row_begin=0, row_end=None):
    r0, r1 = _rows(a, row_begin, row_end)
    h = _vp()
    _check(matrix_ops.get_lib().spgemm_b200_csr_dev(a._h, b._h, int(bool(upper_only)), r0, r1, ctypes.byref(h)),
           "spgemm_b200_csr_dev")
    return DeviceResult(h)


def spgemm_dense(a, b, upper_only=False, row_begin=0, row_end=None, out=None):
    r0, r1 = _rows(a, row_begin, row_end)
    if out is None:
        out = DeviceDense(r1 - r0, b.shape[1])
    ptr = out.ptr if isinstance(out, DeviceDense) else int(out)
    _check(matrix_ops.get_lib().spgemm_b200_dense_dev(a._h, b._h, int(bool(upper_only)), r0, r1, _vp(ptr)),
           "spgemm_b200_dense_dev")
    return out


def triple_product(h, q, ht=None, upper_only=True, row_begin=0, row_end=None, out=None):
    r0, r1 = _rows(h, row_begin, row_end)
    if out is None:
        out = DeviceDense(r1 - r0, h.shape[0])
    ptr = out.ptr if isinstance(out, DeviceDense) else int(out)
    _check(matrix_ops.get_lib().spgemm_b200_triple_dev(h._h, q._h, ht._h if ht is not None else None,
                                                       int(bool(upper_only)), r0, r1, _vp(ptr)),
           "spgemm_b200_triple_dev")
    return out


def mirror(out, n):
    ptr = out.ptr if isinstance(out, DeviceDense) else int(out)
    _check(matrix_ops.get_lib().spgemm_b200_mirror_dev(_vp(ptr), int(n)), "spgemm_b200_mirror_dev")


def symmetrize(out, n):
    ptr = out.ptr if isinstance(out, DeviceDense) else int(out)
    _check(matrix_ops.get_lib().spgemm_b200_symmetrize_dev(_vp(ptr), int(n)), "spgemm_b200_symmetrize_dev")


def row_costs(a, b, q=None, upper_only=False, dense_cols=0):
    """(device buffer of int64 per-row costs, total).  Sparse/dense: a=A, b=B (dense_cols = columns of a dense output
    row, 0 for sparse output).  Triple: a=H, b=H^T, q=Q."""
    lib = matrix_ops.get_lib()
    d = lib.spgemm_b200_device_alloc(max(1, a.shape[0]) * 8)
    total = ctypes.c_int64(0)
    _check(lib.spgemm_b200_row_costs(a._h, b._h, q._h if q is not None else None, int(bool(upper_only)), int(dense_cols),
                                     _vp(d), ctypes.byref(total)), "spgemm_b200_row_costs")
    return d, int(total.value)


TRIPLE_TAIL_COEFF = 8.6      # SPGEMM_B200_TRIPLE_TAIL_COEFF (include/spgemm_b200.h)


def partition_rows(d_costs, rows, parts, tail_indptr=None, tail_coeff=TRIPLE_TAIL_COEFF):
    """Cost-balanced contiguous row bounds (len parts+1) -- the multi-GPU replacement of limits().  tail_indptr (the
    host row pointers of H) selects the triple-product variant, where the block that starts at row r also pays for
    transposing rows r.. of H."""
    out = np.zeros(parts + 1, dtype=np.int32)
    lib = matrix_ops.get_lib()
    if tail_indptr is not None:
        ptr = np.ascontiguousarray(tail_indptr, dtype=np.int32)
        _check(lib.spgemm_b200_partition_tail(_vp(d_costs), ptr.ctypes.data_as(_i32p), float(tail_coeff), int(rows),
                                              int(parts), out.ctypes.data_as(_i32p)), "spgemm_b200_partition_tail")
    else:
        _check(lib.spgemm_b200_partition(_vp(d_costs), int(rows), int(parts), out.ctypes.data_as(_i32p)),
               "spgemm_b200_partition")
    return out
