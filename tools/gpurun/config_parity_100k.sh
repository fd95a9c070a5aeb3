# config 100k.json (BASELINE.json configs[4]) on the golden n = 102400 code: SPA as shipped and the NMSA variant, 100 trials
for c in config100k config100k_nmsa; do
python tools/config_parity.py --config $c --ref-trials 100 --full --out gpurun_out/${c}_parity.json > /dev/null 2> gpurun_out/$c.err; tail -2 gpurun_out/$c.err
python - <<PY
import json
d=json.load(open("gpurun_out/${c}_parity.json")); r=d["runs"]
print("$c ref cpu %.1fs (%d trials, %d threads)"%(r["reference_cpu"]["seconds"], d["ref_trials"], r["reference_cpu"]["threads"]))
for k in ("qkdldpc_sim_fp32","qkdldpc_sim_fp64"):
    print(" ",k,"%.1fs"%r[k]["seconds"],"csv identical",r[k]["csv_identical"], [(x["fer_ref"],x["fer_gpu"],x["fer_inside_ci"],x["iter_mean_ref"],x["iter_mean_gpu"]) for x in r[k]["rows"]])
PY
done
