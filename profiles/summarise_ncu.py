#!/usr/bin/env python
"""ncu --page raw --csv export of one substep capture -> profiles/ncu_<tag>.json (per-kernel table) and
profiles/ncu_traffic.json (DRAM bytes per launch, what bench.py's roofline.traffic reports).

    ncu --set full --clock-control none --import-source on --profile-from-start off \\
        -k regex:'k_(integrate|cull|gjk|epa|manifold|solve_pos|solve_vel)' -c 7 -o gpurun_out/substep python bench.py --ncu-frame 40
    ncu -i gpurun_out/substep.ncu-rep --page raw --csv > gpurun_out/substep.raw.csv
    python profiles/summarise_ncu.py gpurun_out/substep.raw.csv r1_final <commit>
"""
import collections
import csv
import json
import os
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
SCALE = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "usecond": 1.0, "us": 1.0, "nsecond": 1e-3, "ns": 1e-3, "msecond": 1e3, "ms": 1e3}
COLS = dict(regs="launch__registers_per_thread", occupancy_pct="sm__warps_active.avg.pct_of_peak_sustained_active",
            lanes="smsp__thread_inst_executed_per_inst_executed.ratio", fp64_pipe_pct="sm__inst_executed_pipe_fp64.avg.pct_of_peak_sustained_active",
            issue_pct="smsp__issue_active.avg.pct_of_peak_sustained_active", lsu_wavefront_pct="l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed",
            stall_long_scoreboard="smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio",
            stall_wait="smsp__average_warps_issue_stalled_wait_per_issue_active.ratio",
            stall_no_instruction="smsp__average_warps_issue_stalled_no_instruction_per_issue_active.ratio",
            stall_barrier="smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio",
            l1_hit_pct="l1tex__t_sector_hit_rate.pct", l2_hit_pct="lts__t_sector_hit_rate.pct", grid="launch__grid_size", block="launch__block_size",
            warp_insts="smsp__inst_executed.sum", local_load_sectors="l1tex__t_sectors_pipe_lsu_mem_local_op_ld.sum")


def main():
    src, tag, commit = sys.argv[1], sys.argv[2], (sys.argv[3] if len(sys.argv) > 3 else "")
    rows = list(csv.reader(open(src)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    idx = {h: i for i, h in enumerate(hdr)}

    def val(r, k):
        try:
            return float(r[idx[k]].replace(",", "")) * SCALE.get(units[idx[k]], 1.0)
        except (KeyError, ValueError):
            return None

    agg = collections.OrderedDict()
    for r in data:
        name = r[idx["Kernel Name"]].split("(")[0].replace("void ", "")
        d = agg.setdefault(name, dict(launches=0, us=0.0, dram_read=0.0, dram_write=0.0, rows=[]))
        d["launches"] += 1
        d["us"] += val(r, "gpu__time_duration.sum")
        d["dram_read"] += val(r, "dram__bytes_read.sum")
        d["dram_write"] += val(r, "dram__bytes_write.sum")
        row = dict(us=val(r, "gpu__time_duration.sum"), dram_read=val(r, "dram__bytes_read.sum"), dram_write=val(r, "dram__bytes_write.sum"))
        row.update({k: val(r, c) for k, c in COLS.items()})
        d["rows"].append(row)
    total = sum(d["us"] for d in agg.values())
    for d in agg.values():
        d["share"] = d["us"] / total
    out = dict(source=os.path.basename(src), commit=commit, kernels=agg,
               what="one substep of frame 40 of the north-star workload (4096 worlds x 257 bodies) on one B200 under ncu --set full; "
                    "per-launch numbers are cold-cache and serialised: shares and diagnosis, not bench values")
    json.dump(out, open(os.path.join(HERE, "ncu_%s.json" % tag), "w"), indent=1)
    json.dump(dict(source="profiles/ncu_%s.json" % tag, commit=commit,
                   kernels={k: dict(dram_read=d["dram_read"] / d["launches"], dram_write=d["dram_write"] / d["launches"], us=d["us"] / d["launches"])
                            for k, d in agg.items()}), open(os.path.join(HERE, "ncu_traffic.json"), "w"), indent=1)
    for k, d in agg.items():
        print("%-16s x%-3d %8.1f us  share %.3f  dram %7.1f MB read %7.1f MB written" % (k, d["launches"], d["us"], d["share"], d["dram_read"] / 1e6, d["dram_write"] / 1e6))


if __name__ == "__main__":
    main()
