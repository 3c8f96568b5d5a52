timeout 1500 python -m pytest tests/test_em_gpu.py tests/test_trainer_gpu.py -m gpu -q -x --timeout 900 2>&1 | tail -12
timeout 300 python tools/step_profile.py 4096 bf16 2>&1 | grep -v -i Warn | head -24 | tee gpurun_out/r02_step_profile_bf16_v1.txt
