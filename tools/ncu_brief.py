"""Key metrics, stall breakdown and hottest instructions of the first kernel of an ncu report:
    python tools/ncu_brief.py report.ncu-rep [top-N instructions]"""
import collections, csv, io, subprocess, sys
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
def page(which):
    return list(csv.reader(io.StringIO(subprocess.run(["ncu", "-i", rep, "--page", which, "--csv"], capture_output=True, text=True).stdout)))
raw = page("raw"); hdr, units, vals = raw[0], raw[1], raw[2]
d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
keys = ['Kernel Name','gpu__time_duration.sum','launch__grid_size','launch__block_size','launch__registers_per_thread','launch__occupancy_limit_registers','launch__occupancy_limit_shared_mem','launch__occupancy_limit_warps','sm__warps_active.avg.pct_of_peak_sustained_active','launch__shared_mem_per_block_dynamic','dram__bytes_read.sum','dram__bytes_write.sum','l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed','l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum','l1tex__data_pipe_lsu_wavefronts_mem_shared.sum','smsp__issue_active.avg.pct_of_peak_sustained_active','sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active','sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active','lts__throughput.avg.pct_of_peak_sustained_elapsed','gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed','smsp__warps_eligible.avg.per_cycle_active','smsp__inst_executed.sum','l1tex__t_sector_hit_rate.pct','lts__t_sector_hit_rate.pct','smsp__thread_inst_executed_per_inst_executed.ratio','sm__inst_executed_pipe_alu.avg.pct_of_peak_sustained_active','sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active','sm__inst_executed_pipe_xu.avg.pct_of_peak_sustained_active']
for k in keys:
    if k in d: print(f"{k:75s} {d[k][0]} {d[k][1]}")
st = [(float(d[h][0]), h) for h in hdr if h.startswith('smsp__average_warps_issue_stalled') and h.endswith('_per_issue_active.ratio')]
print("stall cycles per issued instruction:")
for v, h in sorted(st, reverse=True)[:10]: print('  %.2f %s' % (v, h.replace('smsp__average_warps_issue_stalled_', '').replace('_per_issue_active.ratio', '')))
src = page("source")
for hi, r in enumerate(src):
    if r and r[0] == 'Address': break
h = src[hi]; ix = {n: i for i, n in enumerate(h)}
per = []; ops = collections.Counter()
for k, r in enumerate(src[hi + 1:]):
    if len(r) < len(h): continue
    try: s = float(r[ix['# Samples']] or 0); n = float(r[ix['Instructions Executed']] or 0)
    except ValueError: continue
    sr = r[ix['Source']].strip()
    op = sr.split()[1] if sr.startswith('@') else sr.split()[0]
    ops[op.split('.')[0]] += n
    dd = {c.replace('stall_', ''): float(r[ix[c]] or 0) for c in h if c.startswith('stall_') and 'Not Issued' not in c}
    per.append((s, k, sr[:70], {c: int(v) for c, v in dd.items() if v > 0.2 * s and v > 20}))
tot = sum(ops.values())
print("instruction mix (warp-level, % of executed):", ", ".join(f"{o} {100*n/tot:.1f}" for o, n in ops.most_common(14)))
print("hottest instructions (samples, index, SASS, dominant stalls):")
for s, k, sr, dd in sorted(per, reverse=True)[:top]: print(f"  {int(s):6d} {k:5d} {sr:70s} {dd}")
