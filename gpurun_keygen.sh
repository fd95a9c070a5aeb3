timeout 1200 python -m pytest tests/test_gpu_keygen.py tests/test_sim_csv.py -x -q 2>&1 | tail -15
python - <<'PY'
import time, numpy as np, torch, sys
sys.path.insert(0,'tests'); import util
import qkd_ldpc_v_b200 as q
from qkd_ldpc_v_b200 import hostlib
arr=util.code_arrays('A79'); words=(arr['n']+31)//32
code=q.LdpcCode(arr['n'],arr['m'],arr['row_ptr'],arr['col_idx'],device=0)
F=65536
seeds=hostlib.trial_seeds(1,F)
da=torch.zeros((F,words),dtype=torch.int32,device='cuda'); db=torch.zeros_like(da)
for rep in range(2):
    t0=time.perf_counter(); code.generate_trial_inputs_device(seeds,0.02,da.data_ptr(),db.data_ptr()); torch.cuda.synchronize(); t=time.perf_counter()-t0
print('device keygen: %.1f ms for %d frames of n=%d -> %.2f M frames/s'%(t*1e3,F,arr['n'],F/t/1e6))
t0=time.perf_counter(); hostlib.gen_keys(seeds[:8192],arr['n'],0.02); t=time.perf_counter()-t0
print('host keygen (all cores): %.2f M frames/s'%(8192/t/1e6))
cfg=q.DecoderConfig(decoding_algorithm=2)
t0=time.perf_counter(); r=code.run_trials(seeds,0.02,(0.71,0),cfg,want_bits=False); t=time.perf_counter()-t0
print('run_trials: %.1f ms -> %.2f Gbit/s end to end incl. input generation, FER %.4f'%(t*1e3, F*arr['n']/t/1e9, 1-(r.flags==3).mean()))
PY
