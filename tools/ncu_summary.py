"""Summarise an .ncu-rep (raw + source pages) into a short text report."""
import csv, re, collections, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active",
        "sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active", "sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active",
        "smsp__issue_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "smsp__warps_active.avg.per_cycle_active",
        "sm__warps_active.avg.pct_of_peak_sustained_active", "lts__t_bytes.sum", "l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum",
        "sm__cycles_elapsed.max", "smsp__inst_executed.sum", "lts__t_sectors_srcunit_tex_op_read.sum", "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed",
        "sm__throughput.avg.pct_of_peak_sustained_elapsed"]
for r in rows[2:]:
    print("kernel:", r[hdr.index("Kernel Name")] if "Kernel Name" in hdr else "")
    for i, h in enumerate(hdr):
        if h in want or ("issue_stalled" in h and h.endswith("per_issue_active.ratio")):
            print(f"  {h} [{units[i]}] = {r[i]}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines()))
his = [i for i, r in enumerate(rows) if r and r[0] == "Address"]
if his:
    hi = his[0]; end = his[1] - 1 if len(his) > 1 else len(rows)
    hdr = rows[hi]; ci = {h: i for i, h in enumerate(hdr)}
    data = [r for r in rows[hi + 1:end] if len(r) == len(hdr) and r[0].startswith("0x")]
    tot_s = sum(int(r[ci["# Samples"]]) for r in data); tot_i = sum(int(r[ci["Instructions Executed"]]) for r in data)
    print(f"source page: {len(data)} SASS lines, {tot_i} warp-instructions, {tot_s} samples")
    op_i, op_s = collections.Counter(), collections.Counter()
    for r in data:
        m = re.match(r"\s*(@!?U?P\d+\s+)?([A-Z0-9_.]+)", r[ci["Source"]])
        op = m.group(2).split(".")[0] if m else "?"
        op_i[op] += int(r[ci["Instructions Executed"]]); op_s[op] += int(r[ci["# Samples"]])
    for k, v in op_i.most_common(22):
        print(f"   {k:10s} {v:11d} {100*v/tot_i:5.1f}% instr   {100*op_s[k]/max(tot_s,1):5.1f}% samples")
