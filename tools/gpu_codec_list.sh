#!/bin/bash
mkdir -p gpurun_out
FSB_CONV_NARROW=1 timeout -s KILL 500 ncu --metrics gpu__time_duration.sum,sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active,sm__warps_active.avg.pct_of_peak_sustained_active,launch__grid_size,smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio,smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio,smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio,smsp__issue_active.avg.pct_of_peak_sustained_active --clock-control none -k regex:"conv1d|pw_gemm|dwconv|fsq_decode" -c 140 --csv --log-file gpurun_out/codec_list.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline > gpurun_out/codec_list.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/codec_list.csv')) if len(r) > 8]
hdr = rows[0]; idi=hdr.index('ID'); ki = hdr.index('Kernel Name'); mi=hdr.index('Metric Name'); vi = hdr.index('Metric Value')
per = collections.OrderedDict()
for r in rows[1:]:
    d = per.setdefault(r[idi], {'k': r[ki][:40]})
    try: d[r[mi]] = float(r[vi].replace(',', ''))
    except: pass
tot = 0
for i, d in per.items():
    t = d.get('gpu__time_duration.sum', 0); tot += t
    print(i, d['k'], 'grid', int(d.get('launch__grid_size', 0)), 'us %.1f' % (t / 1000 if t > 5000 else t), 'fma%% %.0f' % d.get('sm__inst_executed_pipe_fma.avg.pct_of_peak_sustained_active', 0), 'issue%% %.0f' % d.get('smsp__issue_active.avg.pct_of_peak_sustained_active', 0), 'warps%% %.0f' % d.get('sm__warps_active.avg.pct_of_peak_sustained_active', 0), 'long %.1f short %.1f bar %.1f mio %.1f' % (d.get('smsp__average_warps_issue_stalled_long_scoreboard_per_issue_active.ratio', 0), d.get('smsp__average_warps_issue_stalled_short_scoreboard_per_issue_active.ratio', 0), d.get('smsp__average_warps_issue_stalled_barrier_per_issue_active.ratio', 0), d.get('smsp__average_warps_issue_stalled_mio_throttle_per_issue_active.ratio', 0)))
print('total', tot)
PY
