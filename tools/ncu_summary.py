"""Summarise an .ncu-rep (read on the CPU box): headline counters, instruction mix, stall samples.
usage: python tools/ncu_summary.py gpurun_out/foo.ncu-rep [n_top_lines]"""
import collections
import csv
import io
import re
import subprocess
import sys

rep = sys.argv[1]
ntop = int(sys.argv[2]) if len(sys.argv) > 2 else 25
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, r = rows[0], rows[1], rows[2]
want = ["Kernel Name", "gpu__time_duration.sum", "sm__cycles_elapsed.max", "launch__registers_per_thread", "launch__grid_size",
        "launch__block_size", "smsp__issue_active.avg.pct_of_peak_sustained_active", "sm__inst_executed.avg.per_cycle_elapsed",
        "smsp__inst_executed.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "dram__bytes_read.sum", "dram__bytes_write.sum",
        "lts__t_bytes.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
        "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for k in want:
    if k in hdr:
        i = hdr.index(k)
        print(f"{k} = {r[i]} {units[i]}")
st = {h: r[i] for i, h in enumerate(hdr) if "pcsamp_warps_issue_stalled" in h and "not_issued" not in h}
tot = sum(float(v) for v in st.values()) or 1
print("stall samples:", ", ".join(f"{k.split('stalled_')[1]}={float(v) / tot:.1%}" for k, v in
                                  sorted(st.items(), key=lambda kv: -float(kv[1]))[:9]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(sass)))
h2 = rows[1]
if "# Samples" not in h2:
    sys.exit(0)
ia, ie, isamp = h2.index("Source"), h2.index("Instructions Executed"), h2.index("# Samples")
ops, samp = collections.Counter(), collections.Counter()
seen = set()
body = []
for row in rows[2:]:
    if len(row) <= isamp or not row[isamp].isdigit() or row[0] in seen:
        continue
    seen.add(row[0])
    body.append(row)
    m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", row[ia])
    op = m.group(2).split(".")[0] if m else "?"
    ops[op] += int(row[ie])
    samp[op] += int(row[isamp])
ti, ts = sum(ops.values()) or 1, sum(samp.values()) or 1
print(f"warp instructions (first kernel) = {ti}")
print("mix:", ", ".join(f"{o}={n / ti:.1%}/{samp[o] / ts:.1%}" for o, n in ops.most_common(16)), "(inst share / stall-sample share)")
for row in sorted(body, key=lambda x: -int(x[isamp]))[:ntop]:
    print(row[isamp].rjust(5), row[ie].rjust(8), row[ia][:110])
