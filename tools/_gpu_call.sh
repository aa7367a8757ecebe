set -x
mkdir -p gpurun_out/r2h
timeout 1200 python -m pytest tests -m gpu -x -q 2>&1 | tail -5 > gpurun_out/r2h/tests.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2h/bench.json 2> gpurun_out/r2h/bench.err
QILQR_BENCH_MAX_ITERS=18 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2h/bench_mi18.json 2> gpurun_out/r2h/bench_mi18.err
QILQR_BENCH_MAX_ITERS=30 timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2h/bench_mi30.json 2> gpurun_out/r2h/bench_mi30.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 40 --csv --log-file gpurun_out/r2h/launch_variant.csv python tools/variant_once.py 0 65536 > gpurun_out/r2h/variant.log 2>&1
cat gpurun_out/r2h/tests.txt
python - <<'PY'
import json
for f in ("bench","bench_mi18","bench_mi30"):
    try:
        d=json.loads(open(f"gpurun_out/r2h/{f}.json").read().strip().splitlines()[-1])
        r=d["roofline"]
        print(f, d["value"], d["ms_per_step"], d["serial_ms_per_step"], r["bulk_ms_per_step"], r["all_launches_ms_per_step"], r["rollout_kernel"], d.get("iterations_per_solve"))
    except Exception as e: print(f, "ERR", e)
PY
