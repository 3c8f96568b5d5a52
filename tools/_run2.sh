timeout 1200 python -m pytest tests/test_em_gpu.py tests/test_gemm_gpu.py tests/test_trainer_gpu.py -m gpu -q --timeout 900 2>&1 | tail -25
