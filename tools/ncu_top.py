"""Summarise an .ncu-rep: headline metrics + top stall instructions.  usage: ncu_top.py rep [n]"""
import csv, subprocess, sys, io
rep = sys.argv[1]; n = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units, vals = rows[0], rows[1], rows[2]
want = ["gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum", "sm__pipe_tensor_cycles_active_realtime.avg.pct",
        "sm__inst_executed_pipe_alu_realtime.avg.pct", "sm__cycles_active.avg", "launch__registers_per_thread", "lts__t_bytes.sum",
        "lts__t_sectors_srcunit_tex_op_read.sum", "l1tex__m_xbar2l1tex_read_bytes.sum", "sm__throughput.avg.pct", "lts__throughput.avg.pct",
        "smsp__issue_active.avg.pct", "smsp__inst_executed.sum", "sm__pipe_tensor_subpipe_imma_cycles_active", "lts__t_sector_hit_rate.pct",
        "smsp__cycles_active.avg", "sm__inst_executed_pipe_uniform", "dram__throughput.avg.pct", "lts__t_bytes.sum.per_second",
        "sm__warps_active.avg.pct", "l1tex__throughput.avg.pct"]
for h, u, v in zip(hdr, units, vals):
    if any(h.endswith(w) or (w in h and h.endswith("elapsed")) for w in want) and "TriageCompute" not in h:
        print(f"{h:90s} {u:12s} {v}")
src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(src)))
hdr = rows[1]; ix = {h: i for i, h in enumerate(hdr)}; data = rows[2:]
tot = sum(int(r[ix['# Samples']]) for r in data)
stall_cols = [h for h in hdr if h.startswith('stall_') and 'Not Issued' not in h]
agg = {h: sum(int(r[ix[h]]) for r in data) for h in stall_cols}
print("total samples", tot, {k: f"{100*v/tot:.1f}%" for k, v in sorted(agg.items(), key=lambda kv: -kv[1])[:8]})
for r in sorted(data, key=lambda r: -int(r[ix['# Samples']]))[:n]:
    s = int(r[ix['# Samples']])
    st = {h: int(r[ix[h]]) for h in stall_cols if int(r[ix[h]]) > 0}
    st = dict(sorted(st.items(), key=lambda kv: -kv[1])[:3])
    print(f"{s:7d} {100*s/tot:5.1f}% exec={r[ix['Instructions Executed']]:>10} {r[ix['Source']].strip()[:64]:64s} {st}")
