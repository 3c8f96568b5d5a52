set -x
timeout 600 python -m pytest "tests/test_em_gpu.py::test_run_em_bf16_b8_variant" -m gpu -q --timeout 600 2>&1 | grep -v Warning | grep "AssertionError\|passed\|failed" | cut -c1-600
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -5
timeout 300 python tools/step_profile.py 4096 bf16 2>&1 | grep -v -i Warn | head -18 | tee gpurun_out/r02_step_profile_bf16_v5.txt
timeout 200 python tools/infer_profile.py 16384 3 10 2>&1 | grep -v -i warn | head -10 | tee gpurun_out/r02_infer_profile_v3.txt
