"""Summarise an .ncu-rep: headline metrics per launch, SASS opcode mix, stall reasons."""
import collections, csv, subprocess, sys
rep = sys.argv[1]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
want = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'launch__registers_per_thread', 'launch__grid_size',
        'launch__occupancy_limit_registers', 'launch__occupancy_limit_shared_mem', 'sm__warps_active.avg.per_cycle_active', 'smsp__inst_executed.sum',
        'sm__inst_executed.avg.per_cycle_elapsed', 'smsp__thread_inst_executed_per_inst_executed.ratio', 'l1tex__t_sector_hit_rate.pct',
        'lts__t_sector_hit_rate.pct', 'l1tex__data_pipe_lsu_wavefronts.sum.pct_of_peak_sustained_elapsed', 'lts__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__warps_eligible.avg.per_cycle_active', 'sm__cycles_elapsed.max',
        'l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum', 'l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum', 'lts__t_sectors.sum', 'lts__t_bytes.sum',
        'smsp__sass_inst_executed_op_shared_ld.sum', 'smsp__sass_inst_executed_op_shared_st.sum', 'l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum']
print(f"{'metric':78s}" + "".join(f"{('#%d' % i):>16s}" for i in range(len(rows) - 2)))
ki = hdr.index('Kernel Name')
print(f"{'kernel':78s}" + "".join(f"{r[ki][5:20]:>16s}" for r in rows[2:]))
for w in want:
    if w in hdr:
        i = hdr.index(w)
        print(f"{w[:70]:70s} {units[i][:7]:7s}" + "".join(f"{r[i][:15]:>16s}" for r in rows[2:]))
sass = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "sass"], capture_output=True, text=True).stdout
sections, cur = [], None
for r in csv.reader(sass.splitlines()):
    if r and r[0] == 'Kernel Name': cur = {'name': r[1], 'rows': []}; sections.append(cur); continue
    if r and r[0] == 'Address': cur['hdr'] = r; continue
    if cur is not None and r: cur['rows'].append(r)
for k, sec in enumerate(sections):
    h = sec['hdr']; iI = h.index('Instructions Executed'); iT = h.index('Thread Instructions Executed'); iS = h.index('# Samples')
    tot = sum(int(r[iI]) for r in sec['rows']); ts = max(1, sum(int(r[iS]) for r in sec['rows']))
    print(f"\n== launch {k}: {sec['name'][:70]}  sass {len(sec['rows'])}  warp-inst {tot}")
    hist = collections.Counter(); histt = collections.Counter(); hs = collections.Counter()
    for r in sec['rows']:
        toks = r[1].split(); op = toks[1] if toks[0].startswith('@') else toks[0]; op = op.split('.')[0]
        hist[op] += int(r[iI]); histt[op] += int(r[iT]); hs[op] += int(r[iS])
    print("  " + "  ".join(f"{op} {c / tot * 100:.1f}%/{histt[op] / max(c, 1):.0f}thr/{hs[op] / ts * 100:.0f}%smp" for op, c in hist.most_common(18)))
    st = {h[i]: sum(int(r[i]) for r in sec['rows']) for i, x in enumerate(h) if x.startswith('stall_') and 'Not Issued' not in x}
    print("  stalls:", {k_: round(v / ts * 100, 1) for k_, v in sorted(st.items(), key=lambda kv: -kv[1])[:8]})
