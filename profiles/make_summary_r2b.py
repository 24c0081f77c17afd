#!/usr/bin/env python
"""profiles/r2b/ncu_r2b_{substep,prologue}_frame40.json (summarise_ncu.py) + profiles/r2b/bench_*.json -> profiles/r2b/ncu_summary.md"""
import json
import os

HERE = os.path.join(os.path.dirname(os.path.abspath(__file__)), "r2b")
d = json.load(open(os.path.join(HERE, "ncu_r2b_substep_frame40.json")))
pr = json.load(open(os.path.join(HERE, "ncu_r2b_prologue_frame40.json")))
old = json.load(open(os.path.join(HERE, "..", "r2", "ncu_r2_substep_frame40.json")))
b20 = json.load(open(os.path.join(HERE, "bench_bench20.json")))
b60 = json.load(open(os.path.join(HERE, "bench_bench60.json")))
ref = json.load(open(os.path.join(HERE, "bench_ref20.json")))
COLS = [("us", "duration us"), ("share", "share of the capture"), ("regs", "regs/thread"), ("occupancy_pct", "achieved occupancy %"),
        ("lanes", "active lanes / instr"), ("fp64_pipe_pct", "FP64 pipe %"), ("issue_pct", "issue slots busy %"),
        ("l1_hit_pct", "L1 hit %"), ("l2_hit_pct", "L2 hit %"), ("stall_long_scoreboard", "stall long_scoreboard / issue"), ("stall_wait", "stall wait / issue"),
        ("stall_barrier", "stall barrier / issue"), ("stall_no_instruction", "stall no_instruction / issue"), ("warp_insts", "warp instructions")]


def table(dd):
    names = list(dd["kernels"].keys())
    out = ["| metric | " + " | ".join("`%s`" % n for n in names) + " |", "|---|" + "---|" * len(names)]
    for key, label in COLS:
        row = []
        for n in names:
            k = dd["kernels"][n]
            v = k["share"] if key == "share" else k["rows"][0].get(key)
            row.append("" if v is None else ("%.3g" % v))
        out.append("| %s | " % label + " | ".join(row) + " |")
    out.append("| DRAM read MB | " + " | ".join("%.1f" % (dd["kernels"][n]["dram_read"] / 1e6) for n in names) + " |")
    out.append("| DRAM written MB | " + " | ".join("%.1f" % (dd["kernels"][n]["dram_write"] / 1e6) for n in names) + " |")
    return out


def us(dd, k):
    return dd["kernels"][k]["rows"][0]["us"]


def lanes(dd, k):
    return dd["kernels"][k]["rows"][0]["lanes"]


tot_new = sum(v["us"] for v in d["kernels"].values())
tot_old = sum(v["us"] for v in old["kernels"].values())
lines = ["# ncu summary, round 2, second session (commit %s)\n" % d["commit"],
         "`ncu --set full --clock-control none --import-source on --profile-from-start off -k regex:k_(integrate|cull|gjk|epa|manifold|solve_pos|solve_vel) -c 7 python bench.py --ncu-frame 40`",
         "on one B200 (gpurun, `scripts/gpu_final.sh`): the seven kernels of substep 0 of frame 40 of the north-star workload (4096 worlds x 257 bodies); a second pass",
         "captured the per-frame prologue. Per-launch numbers under ncu are cold-cache and serialised: shares and diagnosis only. Bench values come from `bench.py`",
         "without a profiler: `bench_bench20.json` %.2f ms/frame, %.4g body-substeps/s on frames 40..59 (e2e %.4g); `bench_bench60.json` %.2f ms/frame, %.4g over the window;" % (
             b20["ms_per_step"], b20["value"], b20["e2e"]["value"], b60["ms_per_step"], b60["value"]),
         "`bench_ref20.json` (the reference on all host cores) %.4g. Raw tables: `ncu_r2b_*.json`; `launches_bench20.csv` lists the first 1200 launches of a" % ref["value"],
         "`bench.py --steps 20 --warmup 3 --no-extras` run (warm-up frames and the first timed ones: light frames, so its shares are those of the window's START).",
         "Stall samples by source line: `profiles/sass_lines.py` (ncu SASS view matched with `nvdisasm -g`).\n",
         "## One substep (frame 40)\n"] + table(d) + ["\n## Per-frame prologue (frame 40)\n"] + table(pr) + ["""
## Reading

* **Sweeps, dataflow form (per-body chains).** `k_solve_pos` %.0f -> %.0f us, `k_solve_vel` %.0f -> %.0f us against the capture of the first session (`profiles/r2/`), same frame, same work:
  active lanes %.1f -> %.1f of 32 (`k_solve_pos`), barrier stalls 4.2 -> %.2f per issued instruction (%.2f: no grid barrier is left); the release fences before the
  chain updates show up as membar stalls instead (~1 cycle per issue); the substep as a whole %.0f -> %.0f us under ncu. `k_manifold` pays two atomics per unit for the live-level masks.
* What is left is the narrowphase: `k_gjk` + `k_epa` + `k_manifold` = %.0f of %.0f us. By source line (`sass_lines.py` on this capture), `k_gjk` spends its samples
  in the support scans (the dot products and staged-vertex loads of `support_index`, 45 %%), `k_epa` in re-deriving geometry from the poses (`to_mat3`,
  `model_matrix`, `transform_point`: 20 %%) and in the dependent loads that start a hit (hit record -> collider -> pose -> simplex: 16 %%) with 8 warps per SM to
  hide them. Both are the price of not storing transformed hulls (DESIGN.md 1), which bought more than it costs.
* DRAM traffic of the substep: %.2f GB.
""" % (us(old, "k_solve_pos<0>"), us(d, "k_solve_pos<0>"), us(old, "k_solve_vel"), us(d, "k_solve_vel"), lanes(old, "k_solve_pos<0>"), lanes(d, "k_solve_pos<0>"),
       d["kernels"]["k_solve_pos<0>"]["rows"][0]["stall_barrier"], d["kernels"]["k_solve_pos<0>"]["rows"][0]["stall_barrier"], tot_old, tot_new,
       us(d, "k_gjk") + us(d, "k_epa") + us(d, "k_manifold"), tot_new,
       sum(v["dram_read"] + v["dram_write"] for v in d["kernels"].values()) / 1e9)]
open(os.path.join(HERE, "ncu_summary.md"), "w").write("\n".join(lines))
print("\n".join(lines[-12:]))
