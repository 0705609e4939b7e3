#!/usr/bin/env python3
"""Aggregate an `ncu --csv --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,
smsp__inst_executed.sum` log (one row per launch and metric) per kernel; writes the markdown table and the JSON that
bench.py reads for `roofline.traffic` / `issue`.
usage: launch_profile.py <ncu.csv> <sequence_evaluations> <first_n_launches> <out.md> <out.json> "<command>" """
import collections
import csv
import json
import re
import sys

path, nse, first, out_md, out_json, cmd = sys.argv[1], int(sys.argv[2]), int(sys.argv[3]), sys.argv[4], sys.argv[5], sys.argv[6]
rows = [r for r in csv.reader(l for l in open(path) if l.startswith('"'))]
h = rows[0]
iid, ik, im, iv = h.index("ID"), h.index("Kernel Name"), h.index("Metric Name"), h.index("Metric Value")
per = collections.OrderedDict()
for r in rows[1:]:
    k = int(r[iid])
    if k >= first:
        continue
    d = per.setdefault(k, {"kernel": re.sub(r"relem::|\(.*$|^void ", "", r[ik])})
    d[r[im]] = float(r[iv].replace(",", ""))
agg = collections.OrderedDict()
for d in per.values():
    a = agg.setdefault(d["kernel"], [0, 0.0, 0.0, 0.0, 0.0, 0.0])
    a[0] += 1
    a[1] += d.get("gpu__time_duration.sum", 0.0)
    a[2] += d.get("dram__bytes_read.sum", 0.0)
    a[3] += d.get("dram__bytes_write.sum", 0.0)
    a[4] += d.get("smsp__inst_executed.sum", 0.0)
    # fp64 arithmetic thread-instructions (fused multiply-adds, multiplies, adds), when the pass collected them
    a[5] += sum(d.get("smsp__sass_thread_inst_executed_op_%s_pred_on.sum" % op, 0.0) for op in ("dfma", "dmul", "dadd"))
tot = [sum(a[k] for a in agg.values()) for k in range(6)]
with open(out_md, "w") as f:
    f.write("| kernel | launches | ms | share | DRAM read GB | DRAM write GB | G warp-instr |\n|---|---|---|---|---|---|---|\n")
    for k, a in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        f.write("| `%s` | %d | %.1f | %.1f %% | %.1f | %.1f | %.2f |\n" % (k, a[0], a[1] / 1e6, 100 * a[1] / tot[1], a[2] / 1e9, a[3] / 1e9, a[4] / 1e9))
    f.write("| total | %d | %.1f | | %.1f | %.1f | %.2f |\n" % (tot[0], tot[1] / 1e6, tot[2] / 1e9, tot[3] / 1e9, tot[4] / 1e9))
json.dump({"command": cmd, "ncu": "--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum,smsp__inst_executed.sum --clock-control none",
           "sequence_evaluations": nse, "dram_bytes_read": tot[2], "dram_bytes_write": tot[3],
           "dram_bytes_per_sequence_evaluation": (tot[2] + tot[3]) / nse,
           "warp_instructions_per_sequence_evaluation": tot[4] / nse, "launches": tot[0],
           "dfma_per_sequence_evaluation": tot[5] / nse,
           "sum_kernel_ms_under_ncu": tot[1] / 1e6}, open(out_json, "w"), indent=1)
print(open(out_md).read())
