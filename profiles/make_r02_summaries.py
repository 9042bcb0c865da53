#!/usr/bin/env python
"""Turns one record run (profiles/run_record_r02.sh -> gpurun_out/<run>/) into the committed round-2 evidence:
ncu summaries (counters, stall reasons, hottest source lines), the counters bench.py quotes next to its roofline
(track_kernel_counters.json, xs_lookup_counters.json), the launch-list shares and the bench lines themselves.

  python profiles/make_r02_summaries.py gpurun_out/r2m r02m
"""
import csv
import json
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
run, tag = sys.argv[1], sys.argv[2]


def sh(*a):
    return subprocess.run([sys.executable, *a], capture_output=True, text=True).stdout


def raw(path):
    rows = list(csv.reader(open(path)))
    return rows[0], rows[1], rows[2:]


def col(h, rows, name, k=0):
    return float(rows[k][h.index(name)].replace(",", ""))


def unit(h, u, name):
    return u[h.index(name)]


def to_bytes(v, u):
    return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]


def to_s(v, u):
    return v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1.0}[u]


def summary(name, header, extra=""):
    r, s = os.path.join(run, name + "_raw.csv"), os.path.join(run, name + "_src.csv")
    txt = header + "\n" + sh(os.path.join(HERE, "ncu_raw_summary.py"), r) + sh(os.path.join(HERE, "ncu_stalls.py"), r) + extra
    if os.path.exists(s):
        txt += "## hottest source lines\n" + "\n".join(l[:190] for l in sh(os.path.join(HERE, "ncu_line_hotspots.py"), s, "40").splitlines())
    open(os.path.join(HERE, f"ncu_{tag}_{name}.txt"), "w").write(txt + "\n")


# ---- k_track, collision tally (the bench headline's configuration)
h, u, rows = raw(os.path.join(run, "track_raw.csv"))
t = to_s(col(h, rows, "gpu__time_duration.sum"), unit(h, u, "gpu__time_duration.sum"))
dram = to_bytes(col(h, rows, "dram__bytes_read.sum"), unit(h, u, "dram__bytes_read.sum")) + \
    to_bytes(col(h, rows, "dram__bytes_write.sum"), unit(h, u, "dram__bytes_write.sum"))
counters = {"beavrs_assembly": {
    "dram_bytes_per_launch": dram, "dram_gbs": dram / t / 1e9, "launch_us": t * 1e6,
    "fp64_pipe_frac": col(h, rows, "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active") / 100,
    "issue_active_frac": col(h, rows, "smsp__issue_active.avg.pct_of_peak_sustained_active") / 100,
    "warps_active_frac": col(h, rows, "sm__warps_active.avg.pct_of_peak_sustained_active") / 100,
    "l2_hit_rate": col(h, rows, "lts__t_sector_hit_rate.pct") / 100,
    "active_lanes_per_instruction": col(h, rows, "smsp__thread_inst_executed_per_inst_executed.ratio"),
    "source": f"profiles/ncu_{tag}_track.txt (ncu --set full, event round 4 of a 10^6-history batch)"}}
json.dump(counters, open(os.path.join(HERE, "track_kernel_counters.json"), "w"), indent=1)
summary("track", f"# k_track<QUAD4, event> on beavrs_assembly, 10^6 histories, round 4, per-element COLLISION tally: ncu --set full ({run})")

# ---- k_track with the track-length tally on: atomics
h, u, rows = raw(os.path.join(run, "track_tl_raw.csv"))
t_tl = to_s(col(h, rows, "gpu__time_duration.sum"), unit(h, u, "gpu__time_duration.sum"))
red_ops = col(h, rows, "l1tex__t_sectors_pipe_lsu_mem_global_op_red.sum")
red_inst = col(h, rows, "smsp__inst_executed_op_global_red.sum")
l2_red = col(h, rows, "lts__t_sectors_srcunit_tex_op_red.sum")
l2_red_pct = col(h, rows, "lts__t_sectors_srcunit_tex_op_red.sum.pct_of_peak_sustained_elapsed")
extra = (f"## atomics of the track-length tally (one FP64 RED per scored segment and distinct bin in the warp)\n"
         f"RED warp instructions {red_inst:.0f}; RED operations (L1 sectors) {red_ops:.0f} = {red_ops / red_inst:.1f} per instruction;\n"
         f"L2 RED sectors {l2_red:.0f} in {t_tl * 1e6:.0f} us = {l2_red / t_tl / 1e9:.1f} G atomics/s = {l2_red_pct:.1f} % of the L2 RED peak;\n"
         f"same launch with the collision tally: {t * 1e6:.0f} us -> the track-length tally costs {100 * (t_tl / t - 1):.0f} % of the kernel,\n"
         f"through the warp pre-reduction (__match_any_sync + shuffles: short_scoreboard) rather than atomic throughput.\n")
summary("track_tl", f"# k_track<QUAD4, event> on beavrs_assembly, 10^6 histories, round 4, per-element TRACKLENGTH tally: ncu --set full ({run})", extra)
json.dump({"red_ops_per_launch": red_ops, "red_warp_instructions": red_inst, "atomics_per_s": l2_red / t_tl,
           "l2_red_pct_of_peak": l2_red_pct, "launch_us": t_tl * 1e6, "launch_us_collision_tally": t * 1e6,
           "source": f"profiles/ncu_{tag}_track_tl.txt"},
          open(os.path.join(HERE, "tracklength_atomics.json"), "w"), indent=1)

# ---- stand-alone lookup, direct kernel (the tile kernel's figures: profiles/ncu_r01j_xs_lookup_plan.txt)
h, u, rows = raw(os.path.join(run, "xs_raw.csv"))
k = len(rows) - 1
t = to_s(col(h, rows, "gpu__time_duration.sum", k), unit(h, u, "gpu__time_duration.sum"))
dram = to_bytes(col(h, rows, "dram__bytes_read.sum", k), unit(h, u, "dram__bytes_read.sum")) + \
    to_bytes(col(h, rows, "dram__bytes_write.sum", k), unit(h, u, "dram__bytes_write.sum"))
xs = {"direct": {"dram_bytes_per_launch": dram, "dram_gbs": dram / t / 1e9, "launch_ms": t * 1e3,
                 "l2_hit_rate": col(h, rows, "lts__t_sector_hit_rate.pct", k) / 100,
                 "active_lanes_per_instruction": col(h, rows, "smsp__thread_inst_executed_per_inst_executed.ratio", k),
                 "source": f"profiles/ncu_{tag}_xs.txt"},
      "tile_presorted": {"dram_bytes_per_launch": 505.0e6, "dram_gbs": 505.0e6 / 0.938e-3 / 1e9, "launch_ms": 0.938,
                         "issue_active_frac": 0.666, "source": "profiles/ncu_r01j_xs_lookup_plan.txt (kernel unchanged since)"}}
json.dump(xs, open(os.path.join(HERE, "xs_lookup_counters.json"), "w"), indent=1)
summary("xs", f"# k_xs_lookup_array (direct lookup) of the bench xs_lookup leg: 91 nuclides, 170 MB, 4M points: ncu --set full ({run})")

# ---- launch list, bench lines, A/B logs
open(os.path.join(HERE, f"launches_{tag}.txt"), "w").write(sh(os.path.join(HERE, "summarize_launches.py"), os.path.join(run, "launches.csv")))
for src, dst in (("bench.json", f"bench_{tag}.json"), ("ref.json", f"bench_{tag}_reference_arm.json"),
                 ("bench_tracklength.json", f"bench_{tag}_tracklength_tally.json"), ("ab_variants.txt", f"ab_{tag}_variants.txt")):
    if os.path.exists(os.path.join(run, src)):
        shutil.copy(os.path.join(run, src), os.path.join(HERE, dst))
for w in ("brick_tet_100", "brick_tet_100_morton", "brick_tet_160", "brick_tet_160_morton"):
    if os.path.exists(os.path.join(run, f"bench_{w}.json")):
        shutil.copy(os.path.join(run, f"bench_{w}.json"), os.path.join(HERE, f"bench_{tag}_{w}.json"))
print(json.dumps(counters), json.dumps(xs["direct"]))
