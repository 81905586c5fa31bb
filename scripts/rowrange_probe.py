"""Time the dense kernel on 1/N of cfg2's rows (what one rank of an N-GPU strong-scaling run executes)."""
import ctypes, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from sparse_matrix_mult_b200 import device as dev, synthetic
from sparse_matrix_mult_b200.matrix_ops import matrix_ops
lib = matrix_ops.get_lib()
w = synthetic.workload("cfg2")
A = dev.DeviceMatrix.from_scipy(w["a"]); B = dev.DeviceMatrix.from_scipy(w["b"])
ms = ctypes.c_double()
for parts in (1, 2, 4, 8):
    rows = 20000 // parts
    out = dev.DeviceDense(rows, 20000)
    for _ in range(3):
        dev.spgemm_dense(A, B, True, 0, rows, out=out)
    ts, ks = [], []
    for _ in range(10):
        lib.spgemm_b200_flush_l2(); lib.spgemm_b200_synchronize()
        lib.spgemm_b200_timer_start(); dev.spgemm_dense(A, B, True, 0, rows, out=out); lib.spgemm_b200_timer_stop(ctypes.byref(ms))
        ts.append(ms.value); ks.append(dev.last_stats()["ms_numeric"])
    print(f"rows={rows:6d} step_ms={np.mean(ts):.4f} kernel_ms={np.mean(ks):.4f} ideal_at_6.9TB/s={rows*20000*8/6.9e9:.4f}")
    out.free()
