"""Summarise an `ncu --set full` report of the hand-written kernels (tools/ncu_targets.py) into a markdown table and
profiles/ncu_traffic.json (DRAM bytes per launch, read by bench.py for `roofline.traffic`).
    python tools/ncu_summary.py gpurun_out/r1b_kernels.ncu-rep profiles/r1b_kernels_ncu.md"""
import csv, io, json, os, re, subprocess, sys
rep, out_md = sys.argv[1], sys.argv[2]
M = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "dram__throughput.avg.pct_of_peak_sustained_elapsed",
     "lts__throughput.avg.pct_of_peak_sustained_elapsed", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
     "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__warps_active.avg.pct_of_peak_sustained_active",
     "launch__registers_per_thread", "sm__cycles_active.avg", "sm__cycles_elapsed.max", "gpc__cycles_elapsed.avg.per_second",
     "l1tex__m_xbar2l1tex_read_bytes_mem_global_op_tma_ld.sum", "launch__grid_size", "launch__block_size",
     "sm__inst_executed_pipe_tensor_subpipe_hmma.avg.pct_of_peak_sustained_active"]
txt = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv", "--metrics", ",".join(M)], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(txt)))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
def val(r, m, scale=1.0):
    if m not in col or r[col[m]] in ("", "n/a"): return None
    v = float(r[col[m]].replace(",", ""))
    u = units[col[m]]
    mult = {"Mbyte": 1e6, "Kbyte": 1e3, "Gbyte": 1e9, "byte": 1, "ms": 1e3, "us": 1, "ns": 1e-3, "Ghz": 1e9, "Mhz": 1e6}.get(u, 1)
    return v * mult * scale
lines = ["| kernel | grid x block | regs | time us | DRAM rd MB | DRAM wr MB | DRAM % | L2 % | SM % | tensor-mem pipe % | TMA load MB | SM GHz | active / elapsed cycles |",
         "|---|---|---|---|---|---|---|---|---|---|---|---|---|"]
traffic = {}
for r in rows[2:]:
    name = r[col["Kernel Name"]]
    short = re.sub(r"\(.*", "", name).replace("void ", "").replace("hdmoe::", "").replace(" ", "")
    t = val(r, M[0]); rd = val(r, M[1]); wr = val(r, M[2])
    f = lambda x, d=1: "-" if x is None else f"{x:.{d}f}"
    tma = val(r, M[12])
    lines.append(f"| `{short}` | {r[col['Grid Size']]} x {r[col['Block Size']]} | {f(val(r, M[8]), 0)} | {f(t)} | {f(rd / 1e6 if rd is not None else None)} | "
                 f"{f(wr / 1e6 if wr is not None else None)} | {f(val(r, M[3]))} | {f(val(r, M[4]))} | {f(val(r, M[5]))} | {f(val(r, M[6]))} | "
                 f"{f(tma / 1e6 if tma is not None else None)} | {f(val(r, M[11]) / 1e9 if val(r, M[11]) else None, 2)} | {f(val(r, M[9]), 0)} / {f(val(r, M[10]), 0)} |")
    if rd is not None and wr is not None and short not in traffic:
        traffic[short] = int(rd + wr)
open(out_md, "w").write("\n".join(lines) + "\n")
tp = os.path.join(os.path.dirname(out_md), "ncu_traffic.json")
old = json.load(open(tp)) if os.path.exists(tp) else {}
old.update(traffic)
json.dump(old, open(tp, "w"), indent=1, sort_keys=True)
print("\n".join(lines))
