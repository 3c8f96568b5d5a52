set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -3 | tee gpurun_out/r02_gputests.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b4096_bf16.json 2> gpurun_out/r02_bench_b4096_bf16.err
tail -c 200 gpurun_out/r02_bench_b4096_bf16.json
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16_b8 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_b4096_bf16_b8.json 2>/dev/null
timeout 300 python tools/step_profile.py 4096 bf16 2>&1 | grep -v -i Warn | head -30 > gpurun_out/r02_step_profile_b4096_bf16.txt
timeout 200 python tools/infer_profile.py 16384 3 10 2>&1 | grep -v -i warn | head -12 > gpurun_out/r02_infer_profile_b16384.txt
timeout 300 python tools/gemm_breakdown.py 4096 bf16 2>&1 | grep -v -i warn > gpurun_out/r02_gemm_breakdown_b4096_bf16.txt
timeout 200 python tools/estep_only.py 4096 all 2>&1 | grep estep > gpurun_out/r02_estep_only_final.txt; cat gpurun_out/r02_estep_only_final.txt
timeout 900 python bench.py --workload infer_sweep --objects 1000000 > gpurun_out/r02_infer_sweep_1gpu_1M.json 2> gpurun_out/r02_infer_sweep_1gpu.err
du -sh gpurun_out
