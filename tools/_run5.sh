timeout 1500 python -m pytest tests/test_em_gpu.py -m gpu -q -x --timeout 900 2>&1 | tail -6
timeout 300 python tools/step_profile.py 4096 bf16 2>&1 | grep -v -i Warn | head -22 | tee gpurun_out/r02_step_profile_bf16_v2.txt
python tools/gemm_breakdown.py 1024 bf16 > gpurun_out/plain.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:'enc1_backward_fused|lnw4k_backward|estep_dpred|lnw_forward|ln1k_forward|lnw_backward_rows|gru_' -s 220 -c 60 -o gpurun_out/prof_stream python tools/gemm_breakdown.py 1024 bf16 > gpurun_out/ncu_stream.log 2>&1
tail -3 gpurun_out/ncu_stream.log
