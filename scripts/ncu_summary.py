#!/usr/bin/env python
"""Summarise an `ncu --set full` capture of the funnel iteration kernel into the small files the repo tracks:
  profiles/<tag>.csv              key metrics (one per line)
  profiles/<tag>_by_opcode.csv    instruction mix and stall attribution from the source page
  profiles/ncu_funnel_<math>.json what bench.py quotes (roofline.traffic, fp64 pipe)
Usage: python scripts/ncu_summary.py gpurun_out/x.ncu-rep <tag> [parity|fast|<name>] [workload description]"""
import collections, csv, io, json, os, re, subprocess, sys

rep, tag = sys.argv[1], sys.argv[2]
math = sys.argv[3] if len(sys.argv) > 3 else "parity"
WHAT = sys.argv[4] if len(sys.argv) > 4 else "151552 chains x 100 iterations"
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
m = {h: (vals[i], units[i]) for i, h in enumerate(hdr)}
keep = ["Kernel Name", "gpu__time_duration.sum", "launch__grid_size", "launch__block_size", "launch__registers_per_thread",
        "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "smsp__inst_executed.sum", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_fp64.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "lts__t_bytes.sum", "sass__inst_executed_local_loads",
        "sass__inst_executed_local_stores", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "smsp__warps_eligible.avg.per_cycle_active",
        "smsp__warps_active.avg.per_cycle_active"]
keep += sorted(h for h in hdr if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("per_issue_active.ratio"))
with open(os.path.join(ROOT, "profiles", tag + ".csv"), "w") as f:
    f.write("metric,value,unit\n")
    for k in keep:
        if k in m:
            f.write("%s,%s,%s\n" % (k, m[k][0], m[k][1]))


def num(k):
    v, u = m[k]
    v = float(v.replace(",", ""))
    return v * {"Gbyte": 1e9, "Mbyte": 1e6, "Kbyte": 1e3, "byte": 1.0}.get(u, 1.0)


out = {"dram_bytes": num("dram__bytes_read.sum") + num("dram__bytes_write.sum"), "dram_read_bytes": num("dram__bytes_read.sum"),
       "dram_write_bytes": num("dram__bytes_write.sum"), "kernel_ms_under_ncu": float(m["gpu__time_duration.sum"][0]),
       "fp64_pipe_active_pct": float(m["sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active"][0]),
       "issue_active_pct": float(m["smsp__issue_active.avg.pct_of_peak_sustained_active"][0]),
       "registers_per_thread": int(float(m["launch__registers_per_thread"][0])), "warp_instructions": float(m["smsp__inst_executed.sum"][0]),
       "source": "profiles/%s.csv (ncu --set full --clock-control none, one rn_k_iter launch, %s)" % (tag, WHAT)}
json.dump(out, open(os.path.join(ROOT, "profiles", "ncu_funnel_%s.json" % math), "w"), indent=1)
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr, data = rows[1], rows[2:]
ix = {h: i for i, h in enumerate(hdr)}
ops, samp = collections.Counter(), collections.Counter()
stalls = collections.Counter()
for r in data:
    op = re.sub(r"^@!?U?P\d+\s+", "", r[ix["Source"]]).split()[0].split(".")[0]
    ops[op] += int(r[ix["Instructions Executed"]] or 0)
    samp[op] += int(r[ix["# Samples"]] or 0)
    for h in hdr:
        if h.startswith("stall_") and "(Not Issued)" not in h:
            stalls[h] += int(r[ix[h]] or 0)
te, ts = sum(ops.values()), sum(samp.values())
with open(os.path.join(ROOT, "profiles", tag + "_by_opcode.csv"), "w") as f:
    f.write("opcode,pct_of_stall_samples,pct_of_warp_instructions\n")
    for op, _ in samp.most_common():
        f.write("%s,%.2f,%.2f\n" % (op, 100.0 * samp[op] / ts, 100.0 * ops[op] / te))
    f.write("# total warp instructions executed per launch,%d\n# stall samples,%d\n" % (te, ts))
    for h, v in stalls.most_common():
        f.write("# %s,%.2f%%\n" % (h, 100.0 * v / max(1, sum(stalls.values()))))
print(json.dumps(out))
