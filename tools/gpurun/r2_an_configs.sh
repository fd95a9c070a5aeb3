# round 2, call AN: BASELINE.json's configs end to end at HEAD -- reference executable vs qkdldpc_sim, CSV against CSV
run() {  # name, ref trials
  timeout 600 python tools/config_parity.py --config $1 --ref-trials $2 --full --out gpurun_out/r2an_$1.json > /dev/null 2> gpurun_out/r2an_$1.err
  python - <<PY
import json
try:
    d = json.load(open("gpurun_out/r2an_$1.json")); r = d["runs"]
    for k, v in r.items():
        rows = v.get("rows") or []
        print("$1", k, {kk: (round(vv, 2) if isinstance(vv, float) else vv) for kk, vv in v.items() if kk in ("seconds", "trials", "csv_identical", "threads", "combinations")},
              ("rows identical %d of %d" % (sum(1 for x in rows if x.get("row_identical")), len(rows))) if rows and "row_identical" in rows[0] else "")
except Exception as e:
    print("$1 failed", e); print(open("gpurun_out/r2an_$1.err").read()[-800:])
PY
}
run config1k 20000
run config10k 5000
run nopt_spa 6000
run nopt_spalin 6000
run adaptiveR 50
run config100k_nmsa 100
