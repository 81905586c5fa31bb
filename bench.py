#!/usr/bin/env python
"""bench.py -- SpGEMM throughput of the B200 path on the BASELINE.json configs.

    python bench.py [--gpus N] [--steps K] [--warmup W] [--workload cfg5] [--impl ours|reference]

One JSON line on stdout (rank 0).  A "step" is one pass of the hot path over one synthetic input:
  value    : whole-job GFLOP/s (2 * intermediate products / time) with the operands already in HBM,
             timed per step with CUDA events on the library stream (max over ranks for N > 1)
  e2e      : the same metric through the public API sparse_matrix_multiply(..., n_gpus=N) with HOST operands in
             pinned memory and a HOST result: H2D + kernels + D2H all inside the timed region
  roofline : the dominant kernel of the workload against the measured HBM copy bandwidth
  cpu_baseline : the reference's own C routine (oracle/_ref) on this box's host cores, bounded sample
  per_config   : (N = 1, default workload only) value / roofline / e2e of the other BASELINE configs
  parity       : (N > 1) checksums of every rank's row block against the same rows of a one-GPU run

Default workload = BASELINE.json configs[4] (cfg5: H 40,000 x 1,000,000, Q banded 1M^2 -> symmetric dense 40,000^2
triple product): the largest single-GPU configuration whose metric north_star targets for multi-GPU scaling
(cfg2 is a zero fill).  Other workloads: cfg1 cfg2 cfg3 cfg4r cfg4 (and the reduced cfg1s cfg2s cfg3s
cfg5s cfg4r<scale>).  --impl reference times the reference's CPU implementation of the same workload (rank 0 only).
"""
import argparse
import ctypes
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
# The benchmark owns its GPUs: let the library's private pool keep every freed workspace between steps (the default
# bound of 32 GB would hand the 117 GB result of cfg4 back to the driver after every step).
os.environ.setdefault("SPGEMM_B200_POOL_KEEP_GB", "170")

METRIC = "SpGEMM GFLOP/s (2 x intermediate products / s)"
UNIT = "GFLOP/s"
DEFAULT_WORKLOAD = "cfg5"
PER_CONFIG = ("cfg1", "cfg2", "cfg3", "cfg4r", "cfg4")


# ------------------------------------------------------------------------------------------------------
# host threads of the CPU legs: torchrun exports OMP_NUM_THREADS=1 to its workers
def use_all_host_threads():
    """Make OpenMP regions use every host core and return the team size the reference routines will really get."""
    n = os.cpu_count() or 1
    os.environ["OMP_NUM_THREADS"] = str(n)            # for a libgomp not loaded yet
    try:
        gomp = ctypes.CDLL("libgomp.so.1", mode=ctypes.RTLD_GLOBAL)
        gomp.omp_set_num_threads(n)                    # for one that is (numpy / torch may have pulled it in)
        return int(gomp.omp_get_max_threads())
    except OSError:
        return n


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


def config_of(info, w, n_gpus):
    """`config` of a bench line -- identical on both arms: the workload and how the GPU arm treats it (whether L2 is
    flushed between timed steps, what a step contains, how many GPUs share it)."""
    kind = w["kind"]
    out = dict(info)
    out["gpu_l2"] = L2_NOTE[operands_small(w["a"], w["b"])]
    out["gpu_step"] = ("H^T is rebuilt on the device inside every timed step" if kind == "triple" else
                       "analysis + symbolic + numeric phases" if kind == "sparse" else "one kernel")
    out["n_gpus"] = int(n_gpus)
    return out


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


DOMINANT = {"dense": "k_dense_rows_red", "sparse": "numeric phase (k_numeric_rank + k_numeric_warp<*>)",
            "triple": "k_triple_runs"}


# ------------------------------------------------------------------------------------------------------
# reference arm / cpu baseline: the reference's own C routines (oracle/_ref) or the oracle port
def cpu_sample(w, name, threads, max_seconds=25.0):
    """Returns (callable running one bounded sample, flops of the sample, kind, cores, description)."""
    from oracle import port, ref
    a, b, kind = w["a"], w["b"], w["kind"]
    cores = threads
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
        # today's src/ cannot run its sparse-output path.  Preferred: the same sources with the
        # Appendix-B repairs (oracle/build_patched_ref.py), multi-threaded and bit-identical to the shipped binary;
        # else the shipped serial binary; else the oracle port.
        patched = have_ref and ref.patched_available()
        p_rows = np.add.reduceat(np.diff(b.indptr).astype(np.int64)[a.indices], a.indptr[:-1][np.diff(a.indptr) > 0]) \
            if a.nnz else np.zeros(0)
        total = int(p_rows.sum())
        budget = int(2.0e7 * max_seconds * (max(1, cores // 2) if patched else 1))    # ~20 M products/s per core
        if total <= budget:
            sub, desc = a, f"full {name}"
        else:
            per_row = np.zeros(a.shape[0], dtype=np.int64)
            per_row[np.diff(a.indptr) > 0] = p_rows
            rows = int(np.searchsorted(np.cumsum(per_row), budget)) + 1
            sub, desc = a[:rows], f"rows [0,{rows}) of {name} ({budget / total:.1%} of the products)"
        flops = 2 * count_products(sub, b)
        if patched:
            return (lambda: ref.omp_patched().sparse(sub, b, sym, copy=False)), flops, "reference", cores, \
                desc + ": reference sparse_" + ("sym" if sym else "nosym") + " from src/ with the SURVEY Appendix-B " \
                "repairs (-O3 -fopenmp; bit-identical to the shipped binary), C call only"
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


def run_reference_arm(args, w, name, info, threads):
    fn, flops, kind, cores, desc = cpu_sample(w, name, threads)
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
            "config": config_of(info, w, args.gpus),
            "run": {"sample": desc, "omp_threads": threads, "step": "one call of the reference C routine on the sample"},
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


class Resident:
    """Operands of one workload resident in HBM, rows [r0, r1) of the product per step."""

    def __init__(self, w, rank=0, world=1, torch_out=False):
        from sparse_matrix_mult_b200 import device as dev
        from sparse_matrix_mult_b200.matrix_ops import matrix_ops
        self.dev, self.lib = dev, matrix_ops.get_lib()
        a, b = w["a"], w["b"]
        self.kind, self.sym = w["kind"], bool(w["kwargs"].get("symmetric"))
        self.n_rows = a.shape[0]
        self.ncols = self.n_rows if self.kind == "triple" else b.shape[1]
        self.A = dev.DeviceMatrix.from_scipy(a)
        self.B = self.A if (b is a) else dev.DeviceMatrix.from_scipy(b)
        self.bounds = [0, self.n_rows]
        if world > 1:
            # flop-balanced partition from the GPU cost pass (the multi-GPU replacement of limits())
            ht = self.A.transpose() if self.kind == "triple" else None
            costs, _ = dev.row_costs(self.A, ht if ht is not None else self.B, self.B if ht is not None else None,
                                     upper_only=(self.kind == "triple" or self.sym),
                                     dense_cols=(self.ncols if self.kind == "dense" else 0))
            self.bounds = [int(x) for x in dev.partition_rows(costs, self.n_rows, world,
                                                             tail_indptr=a.indptr if self.kind == "triple" else None)]
            self.lib.spgemm_b200_device_free(costs)
            if ht is not None:
                ht.free()
        self.r0, self.r1 = self.bounds[rank], self.bounds[rank + 1]
        self.out, self.out_t = None, None
        if self.kind != "sparse":
            if torch_out:
                import torch
                self.out_t = torch.empty((self.r1 - self.r0, self.ncols), dtype=torch.float64, device="cuda")
                self.out_ptr = self.out_t.data_ptr()
            else:
                self.out = dev.DeviceDense(self.r1 - self.r0, self.ncols)
                self.out_ptr = self.out.ptr

    def step(self, keep=False, rows=None):
        dev = self.dev
        r0, r1 = rows if rows else (self.r0, self.r1)
        if self.kind == "dense":
            dev.spgemm_dense(self.A, self.B, self.sym, r0, r1, out=self.out_ptr)
        elif self.kind == "sparse":
            res = dev.spgemm_csr(self.A, self.B, self.sym, r0, r1)
            if keep:
                return res
            res.free()
        else:
            # includes building H^T on the device (with its rows sorted) every step, at every N
            dev.triple_product(self.A, self.B, None, True, r0, r1, out=self.out_ptr)
        return None

    def free(self):
        if self.out is not None:
            self.out.free()
        self.out_t = None
        if self.B is not self.A:
            self.B.free()
        self.A.free()


def time_resident(res, steps, warm, small, barrier):
    """-> dict(step_ms[], kernel_ms[], launches, bytes_min, stats) of `steps` timed steps after `warm` warm-ups."""
    lib, dev = res.lib, res.dev
    ms_c = ctypes.c_double(0.0)
    for _ in range(warm):
        res.step()
    barrier()
    step_ms, kernel_ms, launches, stats = [], [], 0, {}
    for _ in range(steps):
        if small:
            lib.spgemm_b200_flush_l2()
        barrier()
        lib.spgemm_b200_timer_start()
        res.step()
        lib.spgemm_b200_timer_stop(ctypes.byref(ms_c))
        step_ms.append(ms_c.value)
        stats = dev.last_stats()
        kernel_ms.append(stats["ms_numeric"])
        launches += stats["launches"]
    barrier()
    return dict(step_ms=step_ms, kernel_ms=kernel_ms, launches=launches, bytes_min=stats.get("bytes_min", 0), stats=stats)


def operands_small(a, b):
    """Operands that could sit in the 126 MB L2 between steps: flush it (timing rules)."""
    if os.environ.get("SPGEMM_BENCH_NO_FLUSH"):                      # experiments only
        return False
    return (csr_bytes(a) + csr_bytes(b)) < (256 << 20)


L2_NOTE = {True: "flushed between timed steps (512 MB written, then 256 MB of it read back so L2 holds no dirty lines "
                 "of the flush buffer)",
           False: "no flush: each step streams more bytes than the 126 MB L2"}


def e2e_single_process(w, flops, steps, warm, n_gpus):
    """sparse_matrix_multiply(..., n_gpus=N) from pinned HOST operands to a HOST result, host wall clock."""
    from sparse_matrix_mult_b200 import device as dev
    from sparse_matrix_mult_b200 import sparse_matrix_multiply
    from sparse_matrix_mult_b200.matrix_ops import multi_last_stats
    a, b, kw = w["a"], w["b"], w["kwargs"]
    ap = pinned_csr(a)
    bp = ap if (b is a) else pinned_csr(b)
    try:
        t0 = time.perf_counter()
        r = sparse_matrix_multiply(ap, bp, n_gpus=n_gpus, **kw)
        first_ms = (time.perf_counter() - t0) * 1e3      # cold pinned-result cache, cold device pools / contexts
        del r
        gc.collect()
        for _ in range(max(0, warm - 1)):
            r = sparse_matrix_multiply(ap, bp, n_gpus=n_gpus, **kw)
            del r
            gc.collect()
        e2e_ms, d2h_bytes = [], 0
        for _ in range(steps):
            t0 = time.perf_counter()
            r = sparse_matrix_multiply(ap, bp, n_gpus=n_gpus, **kw)
            e2e_ms.append((time.perf_counter() - t0) * 1e3)
            d2h_bytes = r.nbytes if isinstance(r, np.ndarray) else (r.data.nbytes + r.indices.nbytes + r.indptr.nbytes)
            del r
            gc.collect()            # outside the timed region: result storage goes back to the pinned cache
        e2e_t = float(np.mean(e2e_ms))
        if n_gpus > 1:
            per = multi_last_stats()
            h2d = sum(int(s["bytes_h2d"]) for s in per)
            d2h = sum(int(s["bytes_d2h"]) for s in per) or d2h_bytes
            device_ms = {k: round(max(s[k] for s in per), 3) for k in
                         ("ms_h2d", "ms_analysis", "ms_symbolic", "ms_numeric", "ms_post", "ms_d2h", "ms_total")}
            device_ms["per_gpu"] = [{k: round(s[k], 3) if k.startswith("ms") else int(s[k]) for k in
                                     ("ms_h2d", "ms_analysis", "ms_numeric", "ms_d2h", "bytes_d2h")} for s in per]
        else:
            st = dev.last_stats()           # of the last end-to-end call: bytes that actually crossed PCIe
            h2d, d2h = int(st["bytes_h2d"]), int(st["bytes_d2h"])
            device_ms = {k: round(st[k], 3) for k in ("ms_h2d", "ms_analysis", "ms_symbolic", "ms_numeric", "ms_post",
                                                      "ms_d2h", "ms_total")}
        return {"value": flops / (e2e_t * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": e2e_t, "first_call_ms": first_ms,
                "h2d_bytes_per_step": h2d, "d2h_bytes_per_step": d2h, "result_bytes": int(d2h_bytes),
                "device_ms": device_ms, "n_gpus": n_gpus,
                "path": "sparse_matrix_multiply(n_gpus=%d): one host thread per GPU inside the library, every GPU "
                        "uploads over its own PCIe link and writes its rows of the host result" % n_gpus if n_gpus > 1
                        else "sparse_matrix_multiply()",
                "timing": "host wall clock around the call"}
    except OverflowError as ex:       # nnz(C) >= 2^31 cannot be returned as a SciPy int32 CSR (BASELINE cfg4)
        return {"value": None, "unit": UNIT, "unavailable": str(ex)}
    finally:
        del ap, bp
        gc.collect()


def roofline_of(kind, name, t, peak, peak_src):
    k_ms = float(np.mean(t["kernel_ms"]))
    achieved = (t["bytes_min"] / 1e9) / (k_ms * 1e-3) if k_ms > 0 else 0.0
    return {"bound": "hbm", "kernel": DOMINANT[kind], "achieved": achieved, "peak": peak, "unit": "GB/s",
            "frac": achieved / peak, "traffic": traffic_for(name), "algorithmic_bytes": int(t["bytes_min"]),
            "kernel_ms": k_ms, "peak_source": peak_src}


def run_per_config(names, steps, peak, peak_src, lib):
    """value / roofline / e2e of the other BASELINE configs on one GPU (short runs: each step is < 1 s of GPU time)."""
    from sparse_matrix_mult_b200 import synthetic
    out = {}
    for name in names:
        t_setup = time.perf_counter()
        try:
            w = synthetic.workload(name)
            info, flops = describe(w, name)
            res = Resident(w)
            small = operands_small(w["a"], w["b"])
            t = time_resident(res, steps, 3, small, lib.spgemm_b200_synchronize)
            res.free()
            ms = float(np.mean(t["step_ms"]))
            entry = {"value": flops / (ms * 1e-3) / 1e9, "unit": UNIT, "ms_per_step": ms, "steps": steps, "warmup": 3,
                     "flops": int(flops), "roofline": roofline_of(w["kind"], name, t, peak, peak_src),
                     "phases_ms": {k: round(t["stats"].get(k, 0.0), 4) for k in
                                   ("ms_analysis", "ms_symbolic", "ms_numeric", "ms_post")},
                     "nnz_c": int(t["stats"].get("nnz_c", 0)), "gpu_launches": int(t["launches"]),
                     "l2": L2_NOTE[small]}
            if name == "cfg4":
                entry["e2e"] = {"value": None, "unit": UNIT,
                                "unavailable": "nnz(C) = 9.7e9 does not fit SciPy's int32 CSR (nor the reference's)"}
            else:
                entry["e2e"] = e2e_single_process(w, flops, max(2, steps // 2), 1, 1)
            lib.spgemm_b200_trim(0)
            del w
            gc.collect()
        except Exception as ex:                # noqa: BLE001 -- a side entry must not lose the headline line
            entry = {"error": repr(ex)}
        entry["setup_s"] = round(time.perf_counter() - t_setup, 1)
        out[name] = entry
    return out


# ------------------------------------------------------------------------------------------------------
# N > 1: checksums of every rank's row block against the same rows of a one-GPU run on rank 0
def block_checksums(kind, res, rows, torch):
    """(count, sum, column-weighted sum, sum of squares) of rows [r0, r1) computed by THIS rank."""
    lib, dev = res.lib, res.dev
    if kind == "sparse":
        h = res.step(keep=True, rows=rows)
        p, i, v = h.device_ptrs()
        idx = torch.empty(max(1, h.nnz), dtype=torch.int32, device="cuda")
        val = torch.zeros(max(1, h.nnz), dtype=torch.float64, device="cuda")
        if h.nnz:
            dev.copy_on_device(idx.data_ptr(), i, h.nnz * 4)
            dev.copy_on_device(val.data_ptr(), v, h.nnz * 8)
        lib.spgemm_b200_synchronize()
        nnz = h.nnz
        h.free()
        if not nnz:
            return [0.0, 0.0, 0.0, 0.0]
        wgt = torch.cos(idx[:nnz].to(torch.float64) * 1e-3)
        return [float(nnz), float(val[:nnz].sum()), float((val[:nnz] * wgt).sum()), float((val[:nnz] ** 2).sum())]
    r0, r1 = rows
    out = torch.empty((r1 - r0, res.ncols), dtype=torch.float64, device="cuda")
    if r1 > r0:
        if kind == "dense":
            dev.spgemm_dense(res.A, res.B, res.sym, r0, r1, out=out.data_ptr())
        else:
            dev.triple_product(res.A, res.B, None, True, r0, r1, out=out.data_ptr())
    lib.spgemm_b200_synchronize()
    wgt = torch.cos(torch.arange(res.ncols, device="cuda", dtype=torch.float64) * 1e-3)
    return [float(torch.count_nonzero(out)), float(out.sum()), float((out * wgt).sum()), float((out ** 2).sum())]


def multi_gpu_parity(res, rank, world, dist, torch):
    """Every rank checksums its own block; rank 0 recomputes every block alone and compares."""
    mine = torch.tensor(block_checksums(res.kind, res, (res.r0, res.r1), torch), device="cuda", dtype=torch.float64)
    allc = [torch.zeros(4, device="cuda", dtype=torch.float64) for _ in range(world)]
    dist.all_gather(allc, mine)
    if rank != 0:
        return None
    worst, ok = 0.0, True
    for r in range(world):
        ref = block_checksums(res.kind, res, (res.bounds[r], res.bounds[r + 1]), torch)
        got = [float(x) for x in allc[r].tolist()]
        ok = ok and got[0] == ref[0]
        for g, w in zip(got[1:], ref[1:]):
            rel = abs(g - w) / max(abs(w), 1e-300)
            worst = max(worst, rel)
    ok = ok and worst < 1e-9
    return {"ok": bool(ok), "max_rel_diff": worst,
            "what": "per-rank row block (non-zero count, sum, cos-weighted sum, sum of squares) vs the same rows "
                    "computed by rank 0 alone"}


def run_ours(args, w, name, info, flops, rank, world, threads):
    from sparse_matrix_mult_b200 import device as dev
    from sparse_matrix_mult_b200.matrix_ops import matrix_ops

    lib = matrix_ops.get_lib()
    local = int(os.environ.get("LOCAL_RANK", 0))
    dev.init(local)
    dist, torch, cpu_group = None, None, None
    if world > 1:
        import torch
        import torch.distributed as dist
        torch.cuda.set_device(local)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
        # host-side waits go through gloo: an NCCL barrier would park a spinning kernel on every GPU while rank 0's
        # end-to-end leg is using them
        cpu_group = dist.new_group(backend="gloo")
        box = [info, flops]
        dist.broadcast_object_list(box, src=0)
        info, flops = box

    a, b, kind = w["a"], w["b"], w["kind"]
    res = Resident(w, rank, world)
    small = operands_small(a, b)

    def barrier():
        lib.spgemm_b200_synchronize()
        if dist:
            dist.barrier(group=cpu_group)

    n_warm = max(3, args.warmup)                                    # timing rules: at least 3 warm-up steps
    # clocks / throttle reasons are sampled from here to the end of the end-to-end leg (the resident leg alone
    # lasts a few milliseconds: too short for nvidia-smi's sampling period)
    clocks = ClockSampler(local)
    clocks.__enter__()
    t = time_resident(res, args.steps, n_warm, small, barrier)
    t_local = float(np.sum(t["step_ms"]))
    launches, bytes_min = t["launches"], t["bytes_min"]
    rank_ms, parity = None, None
    if dist:
        tt = torch.tensor([t_local], device="cuda", dtype=torch.float64)
        every = [torch.zeros(1, device="cuda", dtype=torch.float64) for _ in range(world)]
        dist.all_gather(every, tt)
        rank_ms = [float(x.item()) / args.steps for x in every]
        t_job = max(rank_ms) * args.steps
        bm = torch.tensor([float(launches)], device="cuda", dtype=torch.float64)
        dist.all_reduce(bm)
        launches = int(bm[0].item())
        parity = multi_gpu_parity(res, rank, world, dist, torch)
    else:
        t_job = t_local
    ms_per_step = t_job / args.steps
    value = flops / (ms_per_step * 1e-3) / 1e9
    # beside it (not the headline): the same steps with the paneled transpose of H kept on the device handle, the
    # way a caller iterating with the same H would run (spgemm_b200_mat_cache_transpose)
    cached = None
    if kind == "triple":
        res.A.cache_transpose(True)
        tc = time_resident(res, args.steps, 2, small, barrier)
        res.A.cache_transpose(False)
        tc_local = float(np.sum(tc["step_ms"]))
        if dist:
            tt = torch.tensor([tc_local], device="cuda", dtype=torch.float64)
            dist.all_reduce(tt, op=dist.ReduceOp.MAX)
            tc_local = float(tt.item())
        cached = {"ms_per_step": tc_local / args.steps, "value": flops / (tc_local / args.steps * 1e-3) / 1e9, "unit": UNIT,
                  "what": "H^T (paneled transpose) kept on the H handle across steps instead of rebuilt in every step"}

    # ---- end to end through the public API (host operands in pinned memory, host result) ---------------
    e2e = None
    if not args.no_e2e:
        if world == 1:
            e2e = e2e_single_process(w, flops, args.steps, min(2, args.warmup), 1)
        elif args.e2e_mode == "nccl":
            from sparse_matrix_mult_b200 import distributed as sd
            e2e = sd.bench_e2e(args, w, flops, rank, world, csr_bytes)
        else:
            # rank 0 drives all N GPUs through the drop-in API; the other ranks keep their GPUs idle and wait on CPU
            barrier()
            if rank == 0:
                e2e = e2e_single_process(w, flops, args.steps, min(2, args.warmup), world)
            barrier()

    clocks.__exit__(None, None, None)
    res.free()
    if dist:
        dist.barrier(group=cpu_group)
        dist.destroy_process_group()
    if rank != 0:
        return
    peak, peak_src = measured_peak()
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": n_warm, "ms_per_step": ms_per_step, "higher_is_better": True, "scaling": "strong",
            "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": config_of(info, w, world),
            "run": {"sample": "full workload", "parallelism": f"rows sharded over {world} GPU(s), cost-balanced"},
            "e2e": e2e, "gpu_launches": int(launches),
            "roofline": roofline_of(kind, name, t, peak, peak_src),
            "phases_ms": {k: round(t["stats"].get(k, 0.0), 4) for k in
                          ("ms_analysis", "ms_symbolic", "ms_numeric", "ms_post")},
            "nnz_c": int(t["stats"].get("nnz_c", 0)),
            "clocks": dict(clocks.summary(), window="timed steps of the resident leg + the end-to-end leg")}
    if cached:
        line["cached_transpose"] = cached
    if world > 1:
        mean = float(np.mean(rank_ms))
        line["parity"] = parity
        line["residual"] = {"rank_ms": [round(x, 4) for x in rank_ms], "bounds": res.bounds,
                            "imbalance_max_over_mean": max(rank_ms) / mean if mean > 0 else None,
                            "rank0_fixed_ms": round(t["stats"].get("ms_analysis", 0.0), 4),
                            "note": "rank_ms = mean CUDA-event time of a step on each rank; rank0_fixed_ms = part of "
                                    "rank 0's step that does not shrink with N (H^T build / analysis)"}
    # ---- CPU baseline beside it (rank 0, N = 1 only) -----------------------------------------------------
    if world == 1 and not args.no_cpu:
        fn, cflops, ckind, cores, desc = cpu_sample(w, name, threads)
        t0 = time.perf_counter()
        fn()                                   # first call: page faults of the result, thread team start-up
        first = time.perf_counter() - t0
        t0 = time.perf_counter()
        fn()
        dt = time.perf_counter() - t0
        line["cpu_baseline"] = {"value": cflops / dt / 1e9, "unit": UNIT, "cores": cores, "kind": ckind,
                                "sample": desc, "seconds": dt, "first_call_seconds": first}
        if kind == "sparse" and ckind == "reference" and cores > 1:
            # beside it: the reference's shipped (serial) binary on the same sample
            from oracle import ref
            sym = bool(w["kwargs"].get("symmetric"))
            sub = a if "full" in desc else a[:int(desc.split("[0,")[1].split(")")[0])]
            t0 = time.perf_counter()
            ref.shipped().sparse(sub, b, sym, copy=False)
            dt = time.perf_counter() - t0
            line["cpu_baseline_shipped_serial"] = {"value": cflops / dt / 1e9, "unit": UNIT, "cores": 1,
                                                   "kind": "reference", "seconds": dt,
                                                   "sample": desc.split(":")[0] + ": shipped libsparse_x86_64.so, 1 thread"}
    if world == 1 and name == DEFAULT_WORKLOAD and not args.no_per_config:
        del w, a, b
        gc.collect()
        lib.spgemm_b200_trim(0)
        line["per_config"] = run_per_config(PER_CONFIG, max(3, min(5, args.steps)), peak, peak_src, lib)
    emit(line)


_REAL_STDOUT = None


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--workload", default=DEFAULT_WORKLOAD)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--no-cpu", action="store_true", help="skip the cpu_baseline leg")
    ap.add_argument("--no-e2e", action="store_true", help="skip the end-to-end leg (profiling runs)")
    ap.add_argument("--no-per-config", action="store_true", help="skip the per_config block of the default run")
    ap.add_argument("--e2e-mode", default="api", choices=["api", "nccl"],
                    help="N > 1 end-to-end leg: the drop-in API with n_gpus=N on rank 0 (default) or the "
                         "one-process-per-GPU NCCL broadcast/gather path of distributed.py")
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
    threads = use_all_host_threads() if (rank == 0) else 1

    from sparse_matrix_mult_b200 import synthetic
    w = synthetic.workload(args.workload)
    if rank == 0:
        info, flops = describe(w, args.workload, args.recount)
    else:
        info, flops = None, None
    if args.impl == "reference":
        run_reference_arm(args, w, args.workload, info, threads)
    else:
        run_ours(args, w, args.workload, info, flops, rank, world, threads)
    return 0


if __name__ == "__main__":
    sys.exit(main())
