timeout 900 python -m pytest tests/test_gpu_user_model.py tests/test_cpp_mirror.py tests/test_gpu_model_variants.py -x -q 2>&1 | tail -30
