#!/usr/bin/env python
"""Build profiles/<round>/SUMMARY.md from the bench lines in gpurun_out/final_*.json (+ 2/8-GPU lines if present)."""
import glob
import json
import os
import sys

rnd = sys.argv[1] if len(sys.argv) > 1 else "r1"
root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
src, dst = os.path.join(root, "gpurun_out"), os.path.join(root, "profiles", rnd)
os.makedirs(dst, exist_ok=True)


def load(path):
    try:
        with open(path) as f:
            txt = [l for l in f.read().splitlines() if l.startswith("{")]
        return json.loads(txt[-1])
    except Exception:
        return None


rows = []
for w in ("cfg1", "cfg2", "cfg3", "cfg4r", "cfg4", "cfg5"):
    d = load(os.path.join(src, f"final_{w}.json"))
    if not d:
        continue
    with open(os.path.join(dst, f"bench_{w}.json"), "w") as f:
        json.dump(d, f)
    r, e, c = d["roofline"], d.get("e2e") or {}, d.get("cpu_baseline") or {}
    rows.append((w, d, r, e, c))
ref = load(os.path.join(src, "final_reference_cfg2.json"))
if ref:
    json.dump(ref, open(os.path.join(dst, "bench_reference_cfg2.json"), "w"))

out = [f"# Measured results, round {rnd[1:]} (one B200, `scripts/final_round.sh`)", "",
       "`value` = operands resident in HBM, CUDA events per step; `e2e` = `sparse_matrix_multiply()` from pinned host",
       "operands to a host result (wall clock); roofline = algorithmic bytes (SURVEY.md 8d) / CUDA-event time of the",
       "compute phase, against MEASURED_PEAKS.json `hbm_gbs` = 6538.9 GB/s (\"of measured\"); CPU = the reference's own C",
       "routine on the box's host cores (`kind`, `cores` in the JSON).", "",
       "| workload | ms/step (resident) | GFLOP/s | compute-phase ms | bytes_min | achieved GB/s | frac of measured peak | e2e ms | e2e GFLOP/s | CPU reference s (cores) | CPU GFLOP/s | launches/step |",
       "|---|---|---|---|---|---|---|---|---|---|---|---|"]
for w, d, r, e, c in rows:
    out.append(f"| {w} | {d['ms_per_step']:.3f} | {d['value']:.1f} | {r['kernel_ms']:.3f} | {r['algorithmic_bytes'] / 1e6:.1f} MB | "
               f"{r['achieved']:.0f} | {r['frac']:.3f} | {e.get('ms_per_step') and round(e['ms_per_step'], 2)} | "
               f"{e.get('value') and round(e['value'], 2)} | {c.get('seconds') and round(c['seconds'], 3)} ({c.get('cores')}) | "
               f"{c.get('value') and round(c['value'], 3)} | {d['gpu_launches'] / d['steps']:.0f} |")
out.append("")
for w, d, r, e, c in rows:
    ph = d.get("phases_ms", {})
    out.append(f"* **{w}**: phases (last step) analysis {ph.get('ms_analysis')} ms, symbolic {ph.get('ms_symbolic')} ms, "
               f"numeric/compute {ph.get('ms_numeric')} ms; nnz(C) or cells {d.get('nnz_c')}; DRAM traffic of the dominant "
               f"kernel (ncu) {r.get('traffic')}; clocks {d.get('clocks')}; CPU sample: {c.get('sample')}")
if ref:
    out += ["", f"Reference arm (`bench.py --impl reference`, default workload): {ref['value']:.4f} GFLOP/s, "
            f"{ref['ms_per_step']:.1f} ms/step, {ref['cpu_baseline']['cores']} cores, {ref['cpu_baseline']['sample']}"]
multi = []
for path in sorted(glob.glob(os.path.join(src, "bench[248]_cfg*.json"))):
    d = load(path)
    if d:
        json.dump(d, open(os.path.join(dst, os.path.basename(path)), "w"))
        multi.append(d)
if multi:
    out += ["", "## Multi-GPU (strong scaling: same workload, rows sharded by the flop-balanced partition)", "",
            "| workload | GPUs | ms/step (max over ranks) | GFLOP/s | e2e ms (host operands on rank 0 -> host result on rank 0) |", "|---|---|---|---|---|"]
    for d in multi:
        e = d.get("e2e") or {}
        out.append(f"| {d['config']['workload']} | {d['n_gpus']} | {d['ms_per_step']:.3f} | {d['value']:.1f} | {e.get('ms_per_step') and round(e['ms_per_step'], 1)} |")
notes = os.path.join(dst, "multi_gpu_notes.md")
if os.path.exists(notes):
    out += ["", open(notes).read().rstrip()]
open(os.path.join(dst, "SUMMARY.md"), "w").write("\n".join(out) + "\n")
print("\n".join(out))
