#!/bin/bash
set -x
mkdir -p gpurun_out
FSB_NO_TCGEN05=1 timeout -s KILL 200 python bench.py --steps 2 --warmup 1 --frames 8 --no-cpu-baseline > gpurun_out/bench_notc.json 2> gpurun_out/bench_notc.err; cat gpurun_out/bench_notc.json
timeout -s KILL 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 1500 --csv --log-file gpurun_out/launches_b.csv python bench.py --steps 1 --warmup 0 --frames 4 --no-cpu-baseline > gpurun_out/ncu_launches_b.log 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/launches_b.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
agg = collections.OrderedDict()
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except: continue
    if r[ui] == 'ns': v /= 1000
    elif r[ui] == 'ms': v *= 1000
    a = agg.setdefault(r[ki][:90], [0, 0.0]); a[0] += 1; a[1] += v
for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]): print(f"{t:12.1f} us  n={n:5d}  avg {t/n:9.1f}  {k}")
PY
