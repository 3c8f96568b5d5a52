set -x
timeout 900 python -m pytest tests/test_em_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -4
timeout 300 python tools/step_profile.py 4096 bf16 2>&1 | grep -v -i Warn | head -18 | tee gpurun_out/r02_step_profile_bf16_v6.txt
