for wl in I80_nmsa_q030 A79_nmsa_q020; do for t in 512 640 768; do
python bench.py --workload $wl --steps 2 --warmup 1 --no-cpu-baseline --no-e2e --path 2 --onchip-threads $t > gpurun_out/oc4_${wl}_$t.json 2>> gpurun_out/oc4.err
python - <<PY
import json
try:
    d=json.load(open('gpurun_out/oc4_${wl}_$t.json')); r=d['roofline']
    print('$wl threads $t: value %.3f Gbit/s ms/step %.1f ctas %d'%(d['value'], d['ms_per_step'], d['config']['pool_tiles']))
except Exception as e: print('$wl $t failed', e)
PY
done; done
tail -3 gpurun_out/oc4.err
