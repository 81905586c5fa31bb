"""End-to-end probe of the single-process multi-GPU driver: sparse_matrix_multiply(..., n_gpus=N) from pinned host
operands to a host result, for several N / host zero-fill thread counts / upload modes.  One JSON line per setting.
   python scripts/e2e_multi_probe.py cfg5 "1 2 4 8" "2 4 8 16" """
import gc
import json
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
os.environ.setdefault("SPGEMM_B200_POOL_KEEP_GB", "170")
import bench  # noqa: E402
from sparse_matrix_mult_b200 import sparse_matrix_multiply, synthetic  # noqa: E402
from sparse_matrix_mult_b200.matrix_ops import last_stats, matrix_ops, multi_last_stats  # noqa: E402


def main():
    name = sys.argv[1] if len(sys.argv) > 1 else "cfg5"
    gpus = [int(x) for x in (sys.argv[2] if len(sys.argv) > 2 else "1 2 4 8").split()]
    zts = [int(x) for x in (sys.argv[3] if len(sys.argv) > 3 else "4").split()]
    have = matrix_ops.get_lib().spgemm_b200_device_count()
    print("host cores", os.cpu_count(), "gpus", have, flush=True)
    w = synthetic.workload(name)
    ap = bench.pinned_csr(w["a"])
    bp = ap if w["b"] is w["a"] else bench.pinned_csr(w["b"])
    for n in gpus:
        if n > have:
            continue
        for no_peer in ((0, 1) if n > 1 else (0,)):
            for zt in zts:
                os.environ["SPGEMM_B200_ZERO_THREADS"] = str(zt)
                os.environ["SPGEMM_B200_MULTI_NO_PEER"] = str(no_peer)
                ms = []
                for it in range(4):
                    t0 = time.perf_counter()
                    r = sparse_matrix_multiply(ap, bp, n_gpus=n, **w["kwargs"])
                    ms.append((time.perf_counter() - t0) * 1e3)
                    del r
                    gc.collect()
                per = multi_last_stats() if n > 1 else [last_stats()]
                print(json.dumps({"workload": name, "n_gpus": n, "zero_threads": zt, "no_peer": no_peer,
                                  "ms": [round(x, 1) for x in ms],
                                  "h2d": [round(s["ms_h2d"], 1) for s in per], "kernel": [round(s["ms_numeric"], 2) for s in per],
                                  "d2h": [round(s["ms_d2h"], 1) for s in per]}), flush=True)
                if no_peer and zt != zts[0]:
                    break


if __name__ == "__main__":
    main()
