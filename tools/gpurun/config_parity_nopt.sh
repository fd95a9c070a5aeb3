# NOPT_R=0,82_SPA / SPA_LIN_APPROX (BASELINE.json configs[2]): reference executable at 6000 trials vs qkdldpc_sim, then the
# config's own 30 000 trials through qkdldpc_sim
for c in nopt_spa nopt_spalin; do
python tools/config_parity.py --config $c --ref-trials 6000 --full --out gpurun_out/${c}_parity.json > /dev/null 2> gpurun_out/$c.err; tail -2 gpurun_out/$c.err
python - <<PY
import json
d=json.load(open("gpurun_out/${c}_parity.json")); r=d["runs"]
print("$c ref cpu %.1fs (%d trials, %d threads)"%(r["reference_cpu"]["seconds"], d["ref_trials"], r["reference_cpu"]["threads"]))
for k in ("qkdldpc_sim_fp32","qkdldpc_sim_fp64"):
    print(" ",k,"%.1fs"%r[k]["seconds"],"csv identical",r[k]["csv_identical"], [(x["fer_ref"],x["fer_gpu"],x["fer_inside_ci"],x["iter_mean_ref"],x["iter_mean_gpu"]) for x in r[k]["rows"]])
f=r.get("qkdldpc_sim_full")
if f: print("  full: %.1fs for %d trials; reference extrapolated %.0fs; rows"%(f["seconds"],f["trials"],f["reference_cpu_seconds_extrapolated"]), [(x["fer"],x["iter_mean"]) for x in f["rows"]], [round(x["decoded_gbit_s"],2) for x in f["sidecar"]])
PY
done
