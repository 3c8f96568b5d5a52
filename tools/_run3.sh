timeout 1200 python -m pytest tests/test_em_gpu.py -m gpu -q --timeout 900 -k "bf16 or cfg3 or full_size or kink" 2>&1 | tail -15
timeout 300 python tools/step_profile.py 4096 bf16 2>&1 | grep -v Warn | tee gpurun_out/r02_step_profile_bf16_before.txt
timeout 300 python tools/step_profile.py 4096 bf16_b8 2>&1 | grep -v Warn | tee gpurun_out/r02_step_profile_bf16_b8_before.txt
