# round 2, call AR: ncu launch list of the default bench command at the final build; ncu --set full of the float64 walking variable node
CMD="python bench.py --steps 2 --warmup 3 --no-cpu-baseline"
$CMD > gpurun_out/r2ar_plain.json 2> gpurun_out/r2ar_plain.err && \
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/r2ar_launches_default_bench.csv $CMD > /dev/null 2>&1
python - <<'PY'
import csv, collections
rows = [r for r in csv.reader(open('gpurun_out/r2ar_launches_default_bench.csv')) if len(r) > 5]
hdr = rows[0]; ki = hdr.index('Kernel Name'); vi = hdr.index('Metric Value'); ui = hdr.index('Metric Unit')
tot = collections.Counter(); cnt = collections.Counter()
for r in rows[1:]:
    try: v = float(r[vi].replace(',', ''))
    except ValueError: continue
    v *= {'ns': 1e-6, 'us': 1e-3, 'ms': 1.0, 's': 1e3}.get(r[ui], 1e-6)
    name = r[ki].split('(')[0][:70]
    tot[name] += v; cnt[name] += 1
s = sum(tot.values())
for k, v in tot.most_common(14): print('%-72s %5d launches %10.2f ms %5.1f %%' % (k, cnt[k], v, 100 * v / s))
PY
CMD2="python bench.py --workload L100k_spa_q084 --frames 1024 --steps 1 --warmup 1 --no-cpu-baseline --no-e2e --no-secondary"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"vn_kernel" -s 20 -c 2 -o gpurun_out/prof_r02ar_l100k_spa64 $CMD2 > gpurun_out/r2ar_ncu.log 2>&1; tail -1 gpurun_out/r2ar_ncu.log
