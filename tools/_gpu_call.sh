mkdir -p gpurun_out/r2e
timeout 600 python -m pytest tests/test_gpu_api_contract.py -x -q 2>&1 | tail -8
python bench.py --no-cpu-baseline > gpurun_out/r2e/bench_1gpu.json 2> gpurun_out/r2e/bench_1gpu.err; tail -3 gpurun_out/r2e/bench_1gpu.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2e/bench_1gpu.json').read().strip().splitlines()[-1])
print({k:b[k] for k in ('value','ms_per_step','n_gpus')}, b['e2e']['value'], b['e2e_from_controls'], b['host_dma'])
PY
