"""`ncu -i REPORT --page raw --csv` -> compact markdown table (+ optional JSON) of the metrics DESIGN.md / bench.py cite.
usage: python tools/ncu_summary.py gpurun_out/X.ncu-rep [--json out.json --labels a,b,c]"""
import csv, io, json, subprocess, sys

rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, data = rows[0], rows[1], rows[2:]
col = {h: i for i, h in enumerate(hdr)}
KEYS = [("time", "gpu__time_duration.sum"), ("dram_rd", "dram__bytes_read.sum"), ("dram_wr", "dram__bytes_write.sum"),
        ("tensor_pct", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_elapsed"),
        ("tc_smem_pipe_pct", "l1tex__data_pipe_tc_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed"),
        ("l2_to_sm_bytes", "l1tex__m_xbar2l1tex_read_bytes.sum"),
        ("dram_pct", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed"),
        ("issue_active_pct", "smsp__issue_active.avg.pct_of_peak_sustained_active"),
        ("regs", "launch__registers_per_thread"), ("grid", "launch__grid_size")]
print("| kernel | " + " | ".join(k for k, _ in KEYS) + " |")
print("|---|" + "---:|" * len(KEYS))
out = []
for r in data:
    name = r[col["Kernel Name"]].split("(")[0][:60]
    vals, ent = [], {"kernel": name}
    for k, m in KEYS:
        v, u = (r[col[m]], units[col[m]]) if m in col else ("n/a", "")
        vals.append("%s %s" % (v, u))
        try:
            ent[k] = float(v)
            ent[k + "_unit"] = u
        except ValueError:
            pass
    out.append(ent)
    print("| `%s` | %s |" % (name, " | ".join(vals)))
if "--json" in sys.argv:
    json.dump(out, open(sys.argv[sys.argv.index("--json") + 1], "w"), indent=1)
