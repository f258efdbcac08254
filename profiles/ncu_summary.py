#!/usr/bin/env python
"""Summarise an .ncu-rep (read here, no GPU): key raw metrics, instruction share per kernel phase
(between barriers) and the hottest source lines.  usage: python profiles/ncu_summary.py gpurun_out/prof_rNN.ncu-rep"""
import collections
import csv
import subprocess
import sys

rep = sys.argv[1]
top_n = int(sys.argv[2]) if len(sys.argv) > 2 and sys.argv[2].isdigit() else 30


def ncu(*args):
    return subprocess.run(["ncu", "-i", rep, *args], capture_output=True, text=True).stdout


rows = list(csv.reader(ncu("--page", "raw", "--csv").splitlines()))
hdr, units = rows[0], rows[1]
KEYS = ["Kernel Name", "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "sm__warps_active.avg.pct_of_peak_sustained_active",
        "launch__registers_per_thread", "launch__occupancy_limit_registers", "launch__occupancy_limit_shared_mem",
        "launch__waves_per_multiprocessor", "smsp__inst_executed.sum", "smsp__thread_inst_executed_per_inst_executed.ratio",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "smsp__cycles_active.avg",
        "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "smsp__inst_executed_op_local_ld.sum",
        "smsp__inst_executed_op_local_st.sum", "lts__t_sectors_op_read.sum", "lts__t_sectors_op_write.sum"]
r = rows[2]
if "--traffic-json" in sys.argv:
    # profiles/k1_traffic.json: the measured DRAM bytes of this capture, tied to the kernel sources it was built from (bench.py
    # reports them as roofline.traffic only while the sources are unchanged):  ... --traffic-json ENVS STEPS
    import hashlib, json, os
    a = sys.argv.index("--traffic-json")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = b"".join(open(os.path.join(root, "minigrid-rl_b200", "csrc", n), "rb").read() for n in ("mgrl_core.cuh", "mgrl_kernels.cu"))
    def val(k):
        v, u = float(r[hdr.index(k)].replace(",", "")), units[hdr.index(k)]
        return v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}[u]
    rec = {"envs": int(sys.argv[a + 1]), "steps": int(sys.argv[a + 2]), "kernel": r[hdr.index("Kernel Name")],
           "dram_bytes_read": val("dram__bytes_read.sum"), "dram_bytes_write": val("dram__bytes_write.sum"),
           "source_sha16": hashlib.sha256(src).hexdigest()[:16], "report": os.path.basename(rep)}
    json.dump(rec, open(os.path.join(root, "profiles", "k1_traffic.json"), "w"), indent=1)
    print("wrote profiles/k1_traffic.json", rec)
print("== raw metrics (first captured launch) ==")
for k in KEYS:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k} = {r[i]} {units[i]}")
for i, h in enumerate(hdr):
    if h.startswith("smsp__average_warps_issue_stalled") and h.endswith("_per_issue_active.ratio"):
        try:
            if float(r[i]) > 0.3:
                print(f"{h} = {r[i]}")
        except ValueError:
            pass

rows = list(csv.reader(ncu("--page", "source", "--csv", "--print-source", "cuda,sass").splitlines()))
agg = collections.OrderedDict()
cur_file = kernel = first = None
col = {}
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Function Name":
        kernel = r[1]; first = first or kernel; continue
    if r[0] == "Line No":
        col = {n: i for i, n in enumerate(r)}; continue
    if kernel != first or r[0] == "":
        continue
    try:
        line = int(r[0])
    except ValueError:
        continue
    if line == 0:
        continue
    a = agg.setdefault((cur_file, line), [r[1], 0.0, 0.0, 0.0])
    try:
        vals = [float(r[col[k]] or 0) for k in ("Instructions Executed", "Thread Instructions Executed", "# Samples")]
    except (ValueError, IndexError):
        continue   # a source line whose text broke the CSV quoting
    a[1] += vals[0]; a[2] += vals[1]; a[3] += vals[2]
tot = sum(a[1] for a in agg.values()) or 1
tots = sum(a[3] for a in agg.values()) or 1
print(f"\n== hottest source lines of {first} (total warp-instructions {tot:.0f}, samples {tots:.0f}) ==")
for (f, l), (src, inst, tin, smp) in sorted(agg.items(), key=lambda kv: -kv[1][3])[:top_n]:
    print(f"{f}:{l:4d} samples={100*smp/tots:4.1f}% inst={100*inst/tot:4.1f}% lanes={tin/max(inst,1):5.1f} | {src.strip()[:80]}")
