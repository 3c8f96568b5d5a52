set -x
timeout 900 python -m pytest tests/test_gemm_gpu.py tests/test_em_gpu.py -m gpu -q -x --timeout 600 2>&1 | tail -6
timeout 200 python tools/estep_only.py 4096 all 2>&1 | grep estep | tee gpurun_out/r02_estep_only_v2.txt
timeout 300 python tools/step_profile.py 4096 bf16 2>&1 | grep -v -i Warn | head -24 | tee gpurun_out/r02_step_profile_bf16_v3.txt
