"""Summarise ncu CSV exports: tools/ncu_summary.py raw.csv [source.csv]"""
import csv, sys
csv.field_size_limit(10**9)
WANT = ['gpu__time_duration.sum', 'sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'dram__bytes_read.sum', 'dram__bytes_write.sum', 'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__t_sectors_op_red.sum', 'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_shared_mem', 'smsp__inst_executed.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
rows = list(csv.reader(open(sys.argv[1])))
hdr, units = rows[0], rows[1]
idx = {h: i for i, h in enumerate(hdr)}
print("kernels:", [r[idx['Kernel Name']][:50] for r in rows[2:]])
for w in WANT:
    if w in idx:
        print(f"{w} [{units[idx[w]]}]:", [r[idx[w]][:12] for r in rows[2:]])
if len(sys.argv) > 2:
    rows = list(csv.reader(open(sys.argv[2])))
    secs, cur = [], None
    for r in rows:
        if r and r[0] == "Kernel Name":
            cur = {"name": r[1][:70], "rows": []}; secs.append(cur); continue
        if cur is not None: cur["rows"].append(r)
    seen = set()
    for sec in secs:
        if sec["name"] in seen: continue
        seen.add(sec["name"])
        hdr = sec["rows"][0]; idx = {h: i for i, h in enumerate(hdr)}
        data = []
        for r in sec["rows"][1:]:
            if len(r) < len(hdr): continue
            try:
                data.append((int(r[idx["# Samples"]]), r[idx["Source"]].strip(), int(r[idx["Instructions Executed"]]),
                             {k: int(r[idx[k]]) for k in hdr if k.startswith("stall_") and "Not Issued" not in k}))
            except ValueError: pass
        tot = sum(d[0] for d in data)
        print("=====", sec["name"], "samples", tot, "sass", len(data), "warp-instr", sum(d[2] for d in data))
        agg = {}
        for d in data:
            for k, v in d[3].items(): agg[k] = agg.get(k, 0) + v
        print("  ", sorted(agg.items(), key=lambda kv: -kv[1])[:6])
        top = sorted(enumerate(data), key=lambda kv: -kv[1][0])[:int(sys.argv[3]) if len(sys.argv) > 3 else 14]
        for i, d in sorted(top):
            st = sorted(d[3].items(), key=lambda kv: -kv[1])[:2]
            print("  ", i, d[0], d[2], d[1][:60], st)
