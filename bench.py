#!/usr/bin/env python
"""bench.py -- SpGEMM throughput of the B200 path on the BASELINE.json configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg2] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one synthetic input:
  value    : whole-job GFLOP/s (2 * intermediate products / time) with the operands already in HBM,
             timed per step with CUDA events on the library stream (max over ranks for N > 1)
  e2e      : the same metric through the public API sparse_matrix_multiply() with HOST operands in
             pinned memory and a HOST result: H2D + kernels + D2H all inside the timed region
  roofline : the dominant kernel of the workload against the measured HBM copy bandwidth
  cpu_baseline : the reference's own C routine (oracle/_ref) on this box's host cores, bounded sample

Default workload = BASELINE.json configs[1] (A 20,000 x 50,000, density 5e-4, A*A^T -> symmetric dense):
the largest config the reference itself can also run on the host.  Other workloads: cfg1 cfg3 cfg4r cfg4 cfg5
(and the reduced cfg1s cfg2s cfg3s cfg5s cfg4r<scale>).  --impl reference times the reference's CPU
implementation of the same workload (rank 0 only).
"""
import argparse
import gc
import json
import os
import subprocess
import sys
import threading
import time

import numpy as np
import scipy.sparse as sp

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "SpGEMM GFLOP/s (2 x intermediate products / s)"
UNIT = "GFLOP/s"


# ------------------------------------------------------------------------------------------------------
# algorithmic denominators; recomputed from the generated matrices on every run
def count_products(a, b):
    return int(np.diff(b.indptr).astype(np.int64)[a.indices].sum())


def triple_flops(h, q, upper=True, chunk=2000):
    """2 * (P1 + P2): P1 = products of H*Q, P2 = sum_i sum_{c in cols(T_i)} |{r >= i : H[r,c] != 0}| with
    T = H*Q merged (structure only), computed in row chunks."""
    p1 = count_products(h, q)
    hp = sp.csr_matrix((np.ones(h.nnz, np.int32), h.indices, h.indptr), shape=h.shape)
    qp = sp.csr_matrix((np.ones(q.nnz, np.int32), q.indices, q.indptr), shape=q.shape)
    ht = sp.csr_matrix(h.T)
    ht.sort_indices()
    n = h.shape[0]
    ht_ptr = ht.indptr.astype(np.int64)
    ht_len = np.diff(ht_ptr)
    # global sort key of every entry (c, r) of H^T: c * n + r  (rows of H^T are sorted, so keys are sorted)
    keys = np.repeat(np.arange(ht.shape[0], dtype=np.int64), ht_len) * n + ht.indices
    p2 = 0
    for r0 in range(0, n, chunk):
        t = (hp[r0:r0 + chunk] @ qp).tocsr()
        cols = t.indices.astype(np.int64)
        if not upper:
            p2 += int(ht_len[cols].sum())
            continue
        rows = np.repeat(np.arange(r0, r0 + t.shape[0], dtype=np.int64), np.diff(t.indptr))
        first_ge = np.searchsorted(keys, cols * n + rows, side="left")      # first entry of H^T[c,:] with r >= i
        p2 += int((ht_ptr[cols + 1] - first_ge).sum())
    return 2 * (p1 + p2), p1, p2


def csr_bytes(x):
    return 12 * x.nnz + 4 * (x.shape[0] + 1)


def cached_counts(name, a, b):
    """Product counts that take minutes to recount on the host (cfg5: ~4 min) may come from
    profiles/flop_counts.json, keyed by workload AND operand sizes; --recount ignores the file."""
    try:
        with open(os.path.join(ROOT, "profiles", "flop_counts.json")) as f:
            e = json.load(f).get(name)
        if e and e["a_nnz"] == int(a.nnz) and e["b_nnz"] == int(b.nnz) and e["a_shape"] == list(a.shape):
            return e
    except Exception:
        pass
    return None


def describe(w, name, recount=False):
    a, b = w["a"], w["b"]
    info = {"workload": name, "kind": w["kind"], "a_shape": list(a.shape), "a_nnz": int(a.nnz),
            "b_shape": list(b.shape), "b_nnz": int(b.nnz)}
    if w["kind"] == "triple":
        c = None if recount else cached_counts(name, a, b)
        if c:
            p1, p2 = int(c["p1"]), int(c["p2_upper"])
            assert p1 == count_products(a, b), "flop_counts.json does not match the generated matrices"
            flops = 2 * (p1 + p2)
            info["counts"] = "P2 from profiles/flop_counts.json (P1 recounted and matched)"
        else:
            flops, p1, p2 = triple_flops(a, b, upper=True)
        info.update(p1=p1, p2_upper=p2)
    else:
        p = count_products(a, b)
        flops = 2 * p
        info.update(products=p)
    info["flops"] = int(flops)
    return info, flops


# ------------------------------------------------------------------------------------------------------
class ClockSampler:
    """nvidia-smi clocks / throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index=0):
        self.index, self.proc, self.lines = index, None, []

    def __enter__(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), f"--query-gpu={self.Q}",
                                          "--format=csv,noheader,nounits", "-lms", "200"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except OSError:
            self.proc = None
        return self

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def __exit__(self, *exc):
        if self.proc:
            time.sleep(0.25)
            self.proc.terminate()
            try:
                self.proc.wait(timeout=2)
            except subprocess.TimeoutExpired:
                self.proc.kill()

    def summary(self):
        sm, mx, reasons = [], 0.0, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 7:
                continue
            try:
                sm.append(float(f[0]))
                mx = max(mx, float(f[1]))
            except ValueError:
                continue
            for nm, v in zip(names, f[3:7]):
                if v.lower().startswith("active"):
                    reasons.add(nm)
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": mx or None,
                "reasons": sorted(reasons), "samples": len(sm)}


def measured_peak():
    path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    try:
        with open(path) as f:
            return float(json.load(f)["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs)"
    except Exception:
        return 6650.0, "fallback (B200_PROFILING.md 6.65 TB/s)"


def traffic_for(name):
    """dram bytes per launch of the dominant kernel from the committed ncu capture, if one exists."""
    try:
        with open(os.path.join(ROOT, "profiles", "roofline_traffic.json")) as f:
            return json.load(f).get(name)
    except Exception:
        return None


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own C routines (oracle/_ref) or the oracle port
def cpu_sample(w, name, max_seconds=25.0):
    """Returns (callable running one bounded sample, flops of the sample, kind, cores, description)."""
    from oracle import port, ref
    a, b, kind = w["a"], w["b"], w["kind"]
    cores = os.cpu_count() or 1
    have_ref = ref.available()
    if kind == "dense":
        sym = bool(w["kwargs"].get("symmetric"))
        flops = 2 * count_products(a, b)
        if have_ref:
            return (lambda: ref.omp().dense(a, b, sym, copy=False)), flops, "reference", cores, \
                f"full {name}: reference dense_{'sym' if sym else 'nosym'} (src/sparse_sparse_dense.cpp, -O3 -fopenmp), C call only"
        return (lambda: port.spgemm_dense(a, b, sym)), flops, "port", 1, f"full {name}: oracle port, 1 thread"
    if kind == "sparse":
        sym = bool(w["kwargs"].get("symmetric"))
        # the only working reference build of the sparse-output path is the shipped serial binary
        p_rows = np.add.reduceat(np.diff(b.indptr).astype(np.int64)[a.indices], a.indptr[:-1][np.diff(a.indptr) > 0]) \
            if a.nnz else np.zeros(0)
        total = int(p_rows.sum())
        budget = int(2.0e7 * max_seconds)              # ~20 M products/s serial
        if total <= budget:
            sub, desc = a, f"full {name}"
        else:
            per_row = np.zeros(a.shape[0], dtype=np.int64)
            per_row[np.diff(a.indptr) > 0] = p_rows
            rows = int(np.searchsorted(np.cumsum(per_row), budget)) + 1
            sub, desc = a[:rows], f"rows [0,{rows}) of {name} ({budget / total:.1%} of the products)"
        flops = 2 * count_products(sub, b)
        if have_ref:
            return (lambda: ref.shipped().sparse(sub, b, sym, copy=False)), flops, "reference", 1, \
                desc + ": reference shipped libsparse_x86_64.so sparse_nosym (serial build, 1 thread)"
        return (lambda: port.spgemm_csr(sub, b, sym)), flops, "port", 1, desc + ": oracle port, 1 thread"
    # triple: cost of the reference is ~ rows^2/2 * nnz/row gathers; bound the row count
    n = a.shape[0]
    gathers_per_s = 4.0e8 * cores / 8.0
    nnz_row = max(1.0, a.nnz / max(1, n))
    rows = int(min(n, np.sqrt(2.0 * gathers_per_s * max_seconds / nnz_row)))
    # memory of the reference: (threads + 1) * rows^2 * 8 bytes
    while rows > 64 and (cores + 1) * rows * rows * 8 > 24e9:
        rows = int(rows * 0.8)
    sub = a[:rows] if rows < n else a
    desc = (f"full {name}" if rows >= n else f"H rows [0,{rows}) of {name} (own {rows}x{rows} upper-triangle problem)")
    flops, _, _ = triple_flops(sub, b, upper=True)
    if have_ref:
        return (lambda: ref.omp().triple(sub, b, 0, copy=False)), flops, "reference", cores, \
            desc + ": reference triple_product (src/sparse_sparse_dense.cpp, -O3 -fopenmp), C call only"
    return (lambda: port.triple_product(sub, b, 0)), flops, "port", 1, desc + ": oracle port, 1 thread"


def run_reference_arm(args, w, name, info):
    fn, flops, kind, cores, desc = cpu_sample(w, name)
    for _ in range(args.warmup):
        fn()
    t0 = time.perf_counter()
    for _ in range(args.steps):
        fn()
    dt = (time.perf_counter() - t0) / args.steps
    val = flops / dt / 1e9
    line = {"impl": "reference", "metric": METRIC, "value": val, "unit": UNIT, "n_gpus": args.gpus,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": dt * 1e3, "higher_is_better": True,
            "scaling": "strong", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(info, sample=desc),
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": cores, "kind": kind, "sample": desc},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}
    emit(line)


# ------------------------------------------------------------------------------------------------------
def pinned_csr(x):
    """Same CSR with its three arrays in page-locked memory (the e2e leg copies its inputs from pinned memory)."""
    from sparse_matrix_mult_b200.matrix_ops import _result_array
    out = sp.csr_matrix(x.shape)
    for nm, dt in (("indptr", np.int32), ("indices", np.int32), ("data", np.float64)):
        src = getattr(x, nm)
        dst = _result_array(src.shape, dt)
        dst[...] = src
        setattr(out, nm, dst)
    return out


def run_ours(args, w, name, info, flops, rank, world):
    from sparse_matrix_mult_b200 import device as dev
    from sparse_matrix_mult_b200 import sparse_matrix_multiply
    from sparse_matrix_mult_b200.matrix_ops import matrix_ops

    lib = matrix_ops.get_lib()
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev.init(local)
    dist = None
    if world > 1:
        import torch
        import torch.distributed as dist_mod
        torch.cuda.set_device(local)
        dist_mod.init_process_group("nccl", device_id=torch.device("cuda", local))
        dist = dist_mod
        box = [info, flops]
        dist.broadcast_object_list(box, src=0)
        info, flops = box

    a, b, kind, kw = w["a"], w["b"], w["kind"], w["kwargs"]
    sym = bool(kw.get("symmetric"))
    n_rows = a.shape[0]

    # ---- operands resident in HBM; rows sharded by the flop-balanced partition for N > 1 -------------
    A = dev.DeviceMatrix.from_scipy(a)
    B = A if (b is a) else dev.DeviceMatrix.from_scipy(b)
    Ht = A.transpose() if kind == "triple" else None
    if world > 1:
        costs, _ = dev.row_costs(A, Ht if kind == "triple" else B, B if kind == "triple" else None,
                                 upper_only=(kind == "triple" or sym))
        bounds = dev.partition_rows(costs, n_rows, world)
        lib.spgemm_b200_device_free(costs)
        r0, r1 = int(bounds[rank]), int(bounds[rank + 1])
    else:
        r0, r1 = 0, n_rows
    out = None
    if kind == "dense":
        out = dev.DeviceDense(r1 - r0, b.shape[1])
    elif kind == "triple":
        out = dev.DeviceDense(r1 - r0, n_rows)

    def step_device():
        if kind == "dense":
            dev.spgemm_dense(A, B, sym, r0, r1, out=out)
        elif kind == "sparse":
            dev.spgemm_csr(A, B, sym, r0, r1).free()
        else:
            dev.triple_product(A, B, None, True, r0, r1, out=out)      # includes building H^T on the device

    small = (csr_bytes(a) + csr_bytes(b)) < (256 << 20)               # operands could sit in the 126 MB L2
    if os.environ.get("SPGEMM_BENCH_NO_FLUSH"):                      # experiments only
        small = False
    ms_c = ctypes_double()

    def barrier():
        lib.spgemm_b200_synchronize()
        if dist:
            dist.barrier()

    n_warm = max(3, args.warmup)                                    # timing rules: at least 3 warm-up steps
    for _ in range(n_warm):
        step_device()
    barrier()
    kernel_ms, step_ms, launches, bytes_min, last_stats = [], [], 0, 0, {}
    # clocks / throttle reasons are sampled from here to the end of the end-to-end leg (the resident leg alone
    # lasts a few milliseconds: too short for nvidia-smi's sampling period)
    clocks = ClockSampler(local)
    clocks.__enter__()
    for _ in range(args.steps):
        if small:
            lib.spgemm_b200_flush_l2()
        barrier()
        lib.spgemm_b200_timer_start()
        step_device()
        lib.spgemm_b200_timer_stop(ms_c)
        step_ms.append(ms_c.value)
        st = dev.last_stats()
        last_stats = st
        kernel_ms.append(st["ms_numeric"])
        launches += st["launches"]
        bytes_min = st["bytes_min"]
    barrier()
    t_local = float(np.sum(step_ms))
    if dist:
        import torch
        t = torch.tensor([t_local], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        t_job = float(t.item())
        bm = torch.tensor([float(bytes_min), float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(bm)
        bytes_min_job, launches = float(bm[0].item()), int(bm[1].item())
    else:
        t_job, bytes_min_job = t_local, float(bytes_min)
    ms_per_step = t_job / args.steps
    value = flops / (ms_per_step * 1e-3) / 1e9

    # ---- end to end through the public API (host operands in pinned memory, host result) ---------------
    e2e = None
    if args.no_e2e:
        pass
    elif world == 1:
        ap = pinned_csr(a)
        bp = ap if (b is a) else pinned_csr(b)
        try:
            for _ in range(min(2, args.warmup)):
                r = sparse_matrix_multiply(ap, bp, **kw)
                del r
                gc.collect()
            e2e_ms, d2h_bytes = [], 0
            for _ in range(args.steps):
                t0 = time.perf_counter()
                r = sparse_matrix_multiply(ap, bp, **kw)
                e2e_ms.append((time.perf_counter() - t0) * 1e3)
                d2h_bytes = r.nbytes if isinstance(r, np.ndarray) else (r.data.nbytes + r.indices.nbytes + r.indptr.nbytes)
                del r
                gc.collect()            # outside the timed region: result storage goes back to the pinned cache
            e2e_t = float(np.mean(e2e_ms))
            st = dev.last_stats()           # of the last end-to-end call: bytes that actually crossed PCIe
            e2e = {"value": flops / (e2e_t * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_t,
                   "h2d_bytes_per_step": int(st["bytes_h2d"]), "d2h_bytes_per_step": int(st["bytes_d2h"]),
                   "result_bytes": int(d2h_bytes),
                   "device_ms": {k: round(st[k], 3) for k in ("ms_h2d", "ms_analysis", "ms_symbolic", "ms_numeric",
                                                               "ms_post", "ms_d2h", "ms_total")},
                   "timing": "host wall clock around sparse_matrix_multiply()"}
        except OverflowError as ex:       # nnz(C) >= 2^31 cannot be returned as a SciPy int32 CSR (BASELINE cfg4)
            e2e = {"value": None, "unit": UNIT, "unavailable": str(ex)}
    else:
        from sparse_matrix_mult_b200 import distributed as sd
        e2e = sd.bench_e2e(args, w, flops, rank, world, csr_bytes)

    clocks.__exit__(None, None, None)
    if dist:
        dist.barrier()
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    k_ms = float(np.mean(kernel_ms))
    achieved = (bytes_min / 1e9) / (k_ms * 1e-3) if k_ms > 0 else 0.0
    dominant = {"dense": "k_dense_rows_red", "sparse": "numeric phase (k_numeric_rank + k_numeric_warp<*>)",
                "triple": "k_triple_rows_red"}[kind]
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": dict(info, parallelism=f"rows sharded over {world} GPU(s), flop-balanced",
                           l2="flushed between timed steps (512 MB written, then 256 MB of it read back so L2 holds no "
                              "dirty lines of the flush buffer)" if small else
                              "no flush: each step streams more bytes than the 126 MB L2"),
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": dominant, "achieved": achieved, "peak": peak, "unit": "GB/s",
                         "frac": achieved / peak, "traffic": traffic_for(name),
                         "algorithmic_bytes": int(bytes_min), "kernel_ms": k_ms, "peak_source": peak_src},
            "phases_ms": {k: round(last_stats.get(k, 0.0), 4) for k in
                          ("ms_analysis", "ms_symbolic", "ms_numeric", "ms_post")},
            "nnz_c": int(last_stats.get("nnz_c", 0)),
            "clocks": dict(clocks.summary(), window="timed steps of the resident leg + the end-to-end leg")}
    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------------
    if world == 1 and not args.no_cpu:
        fn, cflops, ckind, cores, desc = cpu_sample(w, name)
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": cflops / dt / 1e9, "unit": UNIT, "cores": cores, "kind": ckind,
                                "sample": desc, "seconds": dt}
        if kind == "sparse":
            # the reference has no working multi-threaded sparse path; for scale, the oracle's
            # OpenMP port of it (bit-identical rows, dynamic row blocks) on every host core
            from oracle import port
            ncores = os.cpu_count() or 1
            sub = a if "full" in desc else a[:int(desc.split("[0,")[1].split(")")[0])]
            port.spgemm_csr(sub, b, sym, omp_blocks=16 * ncores, copy=False)          # warm-up (threads, pages)
            t0 = time.perf_counter()
            port.spgemm_csr(sub, b, sym, omp_blocks=16 * ncores, copy=False)
            dt = time.perf_counter() - t0
            line["cpu_baseline_omp_port"] = {"value": cflops / dt / 1e9, "unit": UNIT, "cores": ncores, "kind": "port",
                                             "sample": desc.split(":")[0] + ": oracle_spgemm_csr_omp, C call only",
                                             "seconds": dt}
    emit(line)


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def ctypes_double():
    import ctypes
    return ctypes.c_double(0.0)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default="cfg2")
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--recount", action="store_true", help="recount products on the host even if cached")
    args = ap.parse_args()
    rank = int(os.environ.get("RANK", 0))
    world = int(os.environ.get("WORLD_SIZE", 1))
    if args.impl == "reference" and rank != 0:
        return 0
    # exactly ONE line on stdout: libraries (NCCL prints its version banner there) are sent to stderr
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)

    from sparse_matrix_mult_b200 import synthetic
    w = synthetic.workload(args.workload)
    if rank == 0:
        info, flops = describe(w, args.workload, args.recount)
    else:
        info, flops = None, None
    if args.impl == "reference":
        run_reference_arm(args, w, args.workload, info)
    else:
        run_ours(args, w, args.workload, info, flops, rank, world)
    return 0


if __name__ == "__main__":
    sys.exit(main())
