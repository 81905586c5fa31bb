"""Where does the host time of one sparse end-to-end call go?  (cfg4r by default)"""
import ctypes, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sparse_matrix_mult_b200 import synthetic, sparse_matrix_multiply
from sparse_matrix_mult_b200 import matrix_ops as mo
w = synthetic.workload(sys.argv[1] if len(sys.argv) > 1 else "cfg4r")
a, b = w["a"], w["b"]
lib = mo.matrix_ops.get_lib()
for it in range(4):
    t0 = time.perf_counter()
    arr_a, arr_b = mo.csr_to_arrays(a), mo.csr_to_arrays(b)
    h = mo._vp()
    t1 = time.perf_counter()
    mo._check(lib.spgemm_b200_csr(a.shape[0], a.shape[1], b.shape[1], *mo._ptrs(arr_a), *mo._ptrs(arr_b), 0, ctypes.byref(h)), "csr")
    t2 = time.perf_counter()
    nnz = lib.spgemm_b200_result_nnz(h)
    indptr = mo._result_array((a.shape[0] + 1,), np.int32); indices = mo._result_array((nnz,), np.int32); data = mo._result_array((nnz,), np.float64)
    t3 = time.perf_counter()
    mo._check(lib.spgemm_b200_result_copy(h, indptr.ctypes.data_as(mo._vp), 0, indices.ctypes.data_as(mo._i32p), data.ctypes.data_as(mo._f64p)), "copy")
    t4 = time.perf_counter()
    lib.spgemm_b200_result_free(h)
    from scipy.sparse import csr_matrix
    out = csr_matrix((data, indices, indptr), shape=(a.shape[0], b.shape[1]), copy=False)
    t5 = time.perf_counter()
    del out, data, indices, indptr
    t6 = time.perf_counter()
    print(f"it{it}: arrays {1e3*(t1-t0):.2f}  csr call {1e3*(t2-t1):.2f}  alloc {1e3*(t3-t2):.2f}  copy {1e3*(t4-t3):.2f}  wrap {1e3*(t5-t4):.2f}  free {1e3*(t6-t5):.2f} ms")
for it in range(3):
    t0 = time.perf_counter(); r = sparse_matrix_multiply(a, b); t1 = time.perf_counter(); del r; t2 = time.perf_counter()
    print(f"api it{it}: call {1e3*(t1-t0):.2f} del {1e3*(t2-t1):.2f} ms")
