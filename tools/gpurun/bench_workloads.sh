# one bench line per secondary workload (profiles/r01_h_*.json)
for wl in A79_nmsa_q020 I80_nmsa_q015 A82_spa_q0162 A82_spalin_q0162; do
python bench.py --workload $wl --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/wl_$wl.json 2> gpurun_out/wl_$wl.err
python -c "
import json; d=json.load(open('gpurun_out/wl_$wl.json')); print('$wl value %.3f e2e %.3f path %s'%(d['value'], d['e2e']['value'], d['config']['decoder_path']))"
done
python bench.py --workload L100k_nmsa_q060 --frames 4096 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/wl_L100k_nmsa_q060.json 2> gpurun_out/wl_L100k.err
python -c "
import json; d=json.load(open('gpurun_out/wl_L100k_nmsa_q060.json')); print('L100k value %.3f e2e %.3f'%(d['value'], d['e2e']['value']))"
python bench.py --path 1 --steps 3 --warmup 3 --no-cpu-baseline > gpurun_out/wl_I80_nmsa_q030_streaming.json 2> gpurun_out/wl_I80s.err
python -c "
import json; d=json.load(open('gpurun_out/wl_I80_nmsa_q030_streaming.json')); print('I80 streaming value %.3f e2e %.3f'%(d['value'], d['e2e']['value']), d['roofline']['both_kernels'])"
