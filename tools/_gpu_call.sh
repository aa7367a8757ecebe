set -x
mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_gpu_dense_q.py tests/test_gpu_full_size.py -x -q 2>&1 | tail -5 > gpurun_out/r2g/tests.txt
timeout 300 python bench.py --no-cpu-baseline > gpurun_out/r2g/bench.json 2> gpurun_out/r2g/bench.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/r2g/launch_variant.csv python tools/variant_once.py 0 65536 > gpurun_out/r2g/variant.log 2>&1
cat gpurun_out/r2g/tests.txt; cat gpurun_out/r2g/bench.json
