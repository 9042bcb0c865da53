#!/usr/bin/env python
"""profiles/scale_r02_summary.txt + copies of the multi-GPU bench lines (profiles/run_scale_r02.sh on N = 2 / 4 / 8,
one-GPU points from the record runs).  Efficiency is plain arithmetic on the `value`s: weak = value_N / (N value_1),
strong = value_N / (N value_1) at fixed global size."""
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
OUT = os.path.join(HERE, "..", "gpurun_out")


def line(path):
    try:
        return json.loads([l for l in open(path) if l.startswith("{")][0])
    except Exception:
        return None


one = {"weak": line(os.path.join(OUT, "r2m", "bench.json")), "c4": line(os.path.join(OUT, "r2o", "c4_core_hex_n1.json")),
       "c5": line(os.path.join(OUT, "r2m", "bench_brick_tet_100.json"))}
rows = []


def add(label, key1, fname, strong, v1=None):
    base = v1 if v1 is not None else one[key1]["value"]
    cells = [f"{base / 1e6:8.2f}"]
    for n in (2, 4, 8):
        d = line(os.path.join(OUT, f"scale{n}", fname.format(n=n)))
        if d is None:
            cells.append("       -")
            continue
        shutil.copy(os.path.join(OUT, f"scale{n}", fname.format(n=n)), os.path.join(HERE, "scale_r02_" + fname.format(n=n)))
        eff = d["value"] / (n * base) if not strong else d["value"] / (n * base)
        cells.append(f"{d['value'] / 1e6:8.2f} ({eff:4.2f}, {d['ms_per_step']:6.1f} ms)")
    rows.append(f"{label:58s} | " + " | ".join(cells))


add("beavrs_assembly weak, 1e6 / GPU / batch", "weak", "weak_n{n}.json", False)
add("beavrs_assembly strong, 1e6 global / batch", "weak", "strong1m_n{n}.json", True)
# the deck's own size (2e5 global): the reference_size leg of the weak files
v1 = one["weak"]["reference_size"]["value"]
cells = [f"{v1 / 1e6:8.2f}"]
for n in (2, 4, 8):
    d = line(os.path.join(OUT, f"scale{n}", f"weak_n{n}.json"))
    r = d and d.get("reference_size")
    cells.append(f"{r['value'] / 1e6:8.2f} ({r['value'] / (n * v1):4.2f}, {r['ms_per_batch']:6.1f} ms)" if r else "       -")
rows.append(f"{'beavrs_assembly strong, 2e5 global x 5 batches (the deck)':58s} | " + " | ".join(cells))
add("C4 core_hex_15_3_17 (9.95 M HEX8) weak, 1e6 / GPU", "c4", "c4_core_hex_weak_n{n}.json", False)
add("C5 brick_tet_100 (6 M TET4) weak, 2e6 / GPU", "c5", "c5_brick_tet_100_weak_n{n}.json", False)
add("C5 brick_tet_100 (6 M TET4) strong, 2e6 global", "c5", "c5_brick_tet_100_strong_n{n}.json", True)
hdr = f"{'M histories/s (efficiency = value_N / (N x value_1), ms per step)':58s} | {'N = 1':>8s} | {'N = 2':>26s} | {'N = 4':>26s} | {'N = 8':>26s}"
txt = "# round 2, one box of B200s, NCCL over NVLink; profiles/run_scale_r02.sh; e2e (host banks in the timed region) is in the JSON files\n" + hdr + "\n" + "\n".join(rows) + "\n"
e2e = []
for n in (2, 4, 8):
    d = line(os.path.join(OUT, f"scale{n}", f"weak_n{n}.json"))
    if d:
        e2e.append(f"N = {n}: {d['e2e']['value'] / 1e6:.1f}")
txt += "e2e of the weak headline (M histories/s; every rank moves 96 MB up and 96 MB down per step through one host): N = 1: " + \
    f"{one['weak']['e2e']['value'] / 1e6:.1f}, " + ", ".join(e2e) + "\n"
open(os.path.join(HERE, "scale_r02_summary.txt"), "w").write(txt)
print(txt)
