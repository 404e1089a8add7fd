"""Summarise an .ncu-rep (one kernel): duration, dram bytes, issue rate, stall reasons, top SASS lines."""
import csv, collections, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(f"ncu -i {rep} --page raw --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines())); hdr = rows[0]; vals = rows[-1]
m = dict(zip(hdr, vals))
keys = ['Kernel Name','gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','launch__grid_size','launch__block_size','launch__registers_per_thread',
        'launch__shared_mem_per_block_dynamic','sm__warps_active.avg.pct_of_peak_sustained_active','smsp__inst_executed.sum','sm__inst_issued.sum.per_cycle_active',
        'sm__icc_request_hit_rate.pct','idc__request_hit_rate.pct','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','sm__cycles_elapsed.max',
        'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','smsp__thread_inst_executed_per_inst_executed.ratio']
for k in keys:
    if k in m: print(f"{k} = {m[k]}")
st = {h.replace('smsp__pcsamp_warps_issue_stalled_', ''): int(float(v)) for h, v in zip(hdr, vals)
      if h.startswith('smsp__pcsamp_warps_issue_stalled_') and not h.endswith('_not_issued') and v not in ('', 'n/a')}
tot = sum(st.values()) or 1
print("stalls:", ", ".join(f"{k} {100*v/tot:.1f}%" for k, v in sorted(st.items(), key=lambda x: -x[1])[:8]))
src = subprocess.run(f"ncu -i {rep} --page source --csv", shell=True, capture_output=True, text=True).stdout
rows = list(csv.reader(src.splitlines())); hdr = rows[1]
ci, cs, ce, ca = hdr.index('Source'), hdr.index('# Samples'), hdr.index('Instructions Executed'), hdr.index('Address')
data = []
for r in rows[2:]:
    try: data.append((int(r[cs] or 0), int(r[ce] or 0), r[ci], r[ca]))
    except Exception: pass
te = sum(d[1] for d in data); ts = sum(d[0] for d in data)
print(f"sass instrs={len(data)} executed_distinct={sum(1 for d in data if d[1]>0)} hot(>=1/10 max)={sum(1 for d in data if d[1] >= max(x[1] for x in data)/10)} warp_inst={te} samples={ts}")
n = int(sys.argv[2]) if len(sys.argv) > 2 else 15
for s_, e_, src_, a_ in sorted(data, reverse=True)[:n]: print(f"  {s_:6d} {100*s_/max(ts,1):5.1f}% exec={e_:9d} {a_[-6:]} {src_[:80]}")
