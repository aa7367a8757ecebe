mkdir -p gpurun_out/r2e
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/r2e/bench_8gpu_b.json 2> gpurun_out/r2e/bench_8gpu_b.err
tail -3 gpurun_out/r2e/bench_8gpu_b.err
python - <<'PY'
import json
b=json.loads(open('gpurun_out/r2e/bench_8gpu_b.json').read().strip().splitlines()[-1])
print({k:b[k] for k in ('value','ms_per_step','n_gpus')}, 'e2e', b['e2e']['value'], b['e2e']['ms_per_step'], 'ctl', b['e2e_from_controls']['value'], b['e2e_from_controls']['ms_per_step'], b['host_dma']['gbs_per_direction_all_gpus'], b['host_dma']['gbs_per_direction_per_gpu'])
PY
