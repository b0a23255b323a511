"""Turn one tools/gpu_check.sh visit (gpurun_out/) into the committed summaries under profiles/.
usage: python tools/summarize_profiles.py r01_final"""
import csv, io, json, os, subprocess, sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OUT, PROF = os.path.join(ROOT, 'gpurun_out'), os.path.join(ROOT, 'profiles')
tag = sys.argv[1] if len(sys.argv) > 1 else 'r01_final'

# 1. bench line
line = [l for l in open(os.path.join(OUT, 'bench.log')) if l.startswith('{')][-1]
json.loads(line)
open(os.path.join(PROF, tag + '_bench.json'), 'w').write(line)

# 2. launch list -> shares
rows = [r for r in csv.reader(l for l in open(os.path.join(OUT, 'launches.csv')) if not l.startswith('=='))]
hdr = rows[0]
ki, vi = hdr.index('Kernel Name'), hdr.index('Metric Value')
agg = {}
for r in rows[1:]:
    if len(r) <= vi or r[hdr.index('Metric Name')] != 'gpu__time_duration.sum':
        continue
    us = float(r[vi].replace(',', '')) / (1e3 if r[hdr.index('Metric Unit')] in ('ns', 'nsecond') else 1.0)
    a = agg.setdefault(r[ki], [0, 0.0])
    a[0] += 1; a[1] += us
tot = sum(v[1] for v in agg.values())
with open(os.path.join(PROF, tag + '_launch_shares.txt'), 'w') as f:
    f.write('# ncu launch list of `bench.py --steps 2 --warmup 1 --pool 2 --no-cpu-baseline --no-e2e` (includes the GRU timing '
            'section); per-launch times are cold-cache/serialised: compare SHARES\n')
    for k, (n, us) in sorted(agg.items(), key=lambda x: -x[1][1]):
        f.write('%-82s n=%3d  avg %8.1f us  share %5.1f%%\n' % (k[:82], n, us / n, 100 * us / tot))
import shutil
shutil.copy(os.path.join(OUT, 'launches.csv'), os.path.join(PROF, tag + '_launches.csv'))

# 3. full capture of the gather kernel -> key metrics + DRAM traffic
raw = subprocess.run(['ncu', '-i', os.path.join(OUT, 'prof_gather.ncu-rep'), '--page', 'raw', '--csv'], capture_output=True, text=True).stdout
rows = list(csv.reader(io.StringIO(raw)))
hdr, units = rows[0], rows[1]
keys = ['Kernel Name', 'gpu__time_duration.sum', 'dram__bytes_read.sum', 'dram__bytes_write.sum', 'lts__t_sectors_srcunit_tex_op_read.sum',
        'lts__t_sector_hit_rate.pct', 'l1tex__t_sector_hit_rate.pct', 'sm__warps_active.avg.pct_of_peak_sustained_active',
        'launch__registers_per_thread', 'launch__occupancy_limit_registers', 'sm__throughput.avg.pct_of_peak_sustained_elapsed',
        'lts__throughput.avg.pct_of_peak_sustained_elapsed', 'l1tex__throughput.avg.pct_of_peak_sustained_elapsed',
        'l1tex__data_pipe_lsu_wavefronts.avg.pct_of_peak_sustained_elapsed', 'l1tex__data_pipe_lsu_wavefronts_mem_shared.sum.pct_of_peak_sustained_elapsed',
        'gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed', 'smsp__inst_executed.sum', 'smsp__issue_active.avg.pct_of_peak_sustained_active',
        'smsp__average_warp_latency_issue_stalled_long_scoreboard.ratio', 'smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_lg_throttle_per_issue_active.ratio',
        'smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 'smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio']
traffic = []
with open(os.path.join(PROF, tag + '_gather_ncu.txt'), 'w') as f:
    f.write('# ncu --set full --clock-control none --import-source on, rgcn_gather_d200_kernel (tile kernel, deterministic atomic-free '
            'hand-over, cp.async epilogue prefetch), ICEWS18-shaped batch, bench.py --steps 2\n')
    for r in rows[2:]:
        if len(r) != len(hdr):
            continue
        f.write('----\n')
        d = dict(zip(hdr, r)); u = dict(zip(hdr, units))
        for k in keys:
            if k in d:
                f.write('%-110s %s %s\n' % (k, d[k], u.get(k, '')))
        def tobytes(k):
            v = float(d[k].replace(',', '')); un = u[k].lower()
            return v * {'byte': 1, 'kbyte': 1e3, 'mbyte': 1e6, 'gbyte': 1e9}.get(un, 1)
        traffic.append(tobytes('dram__bytes_read.sum') + tobytes('dram__bytes_write.sum'))
json.dump({'rgcn_gather_bytes_per_launch': sum(traffic) / len(traffic),
           'source': 'profiles/%s_gather_ncu.txt (dram__bytes_read.sum + dram__bytes_write.sum, mean over %d captured launches, cold L2 under ncu cache control)' % (tag, len(traffic))},
          open(os.path.join(PROF, 'roofline_traffic.json'), 'w'), indent=1)
print('wrote', tag, 'traffic MB', sum(traffic) / len(traffic) / 1e6)
