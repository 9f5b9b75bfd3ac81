"""Turns gpurun_out/*.csv / *.ncu-rep into the text summaries committed under profiles/.
usage: python tools/ncu_summarize.py <tag> <launches.csv> <prof.ncu-rep> "<command line that was profiled>" """
import collections, csv, json, subprocess, sys

tag, launches, rep, cmd = sys.argv[1:5]
rows = list(csv.reader(open(launches)))
hi = next(i for i, r in enumerate(rows) if 'Kernel Name' in r)
hdr = rows[hi]; kn = hdr.index('Kernel Name'); mv = hdr.index('Metric Value')
agg = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= mv:
        continue
    name = r[kn].split('(')[0][:70]
    agg[name][0] += 1; agg[name][1] += float(r[mv].replace(',', ''))
out = [f"# {tag}: ncu launch list (gpu__time_duration.sum, --clock-control none)", f"#   {cmd}",
       "# per-launch times are cold-cache and serialised: compare SHARES, not absolutes",
       f"{'kernel':72s} {'launches':>8s} {'total_us':>12s} {'avg_us':>10s}"]
for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    out.append(f"{k:72s} {v[0]:8d} {v[1] / 1e3:12.1f} {v[1] / v[0] / 1e3:10.1f}")
open(f'profiles/{tag}_launches_summary.txt', 'w').write("\n".join(out) + "\n")
raw = subprocess.check_output(f"ncu -i {rep} --page raw --csv", shell=True).decode()
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
keep = ['gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors.sum', 'launch__grid_size',
        'launch__registers_per_thread', 'launch__occupancy_limit_shared_mem', 'launch__occupancy_limit_registers',
        'sm__warps_active.avg.pct_of_peak_sustained_active', 'smsp__inst_executed.sum',
        'smsp__issue_active.avg.pct_of_peak_sustained_active', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed_op_shared_atom.sum',
        'sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active', 'smsp__thread_inst_executed_per_inst_executed.ratio',
        'sm__inst_executed_pipe_alu.sum', 'sm__inst_executed_pipe_xu.sum', 'sm__inst_executed_pipe_fma.sum',
        'sm__inst_executed_pipe_lsu.sum', 'sm__pipe_alu_cycles_active.avg.pct_of_peak_sustained_active',
        'sm__pipe_xu_cycles_active.avg.pct_of_peak_sustained_active']
keep += [h for h in hdr if 'issue_stalled' in h and h.endswith('_per_issue_active.ratio')]
seen = set()
with open(f'profiles/{tag}_ncu_full.txt', 'w') as f:
    f.write(f"# {tag}: ncu --set full --clock-control none --import-source on\n#   {cmd}\n# first launch of each kernel shown\n")
    for r in rows[2:]:
        name = r[hdr.index('Kernel Name')].split('(')[0]
        if name in seen:
            continue
        seen.add(name)
        f.write("\n== " + name + "\n")
        for h in keep:
            if h in hdr:
                f.write(f"  {h:95s} {r[hdr.index(h)]:>18s} {units[hdr.index(h)]}\n")
print(open(f'profiles/{tag}_launches_summary.txt').read()[:2600])
