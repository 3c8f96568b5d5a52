set -x
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b4096_bf16.json 2> gpurun_out/r02_bench_b4096_bf16.err
tail -c 200 gpurun_out/r02_bench_b4096_bf16.json
timeout 300 python tools/step_profile.py 4096 bf16 2>&1 | grep -v -i Warn | head -30 > gpurun_out/r02_step_profile_b4096_bf16.txt
timeout 300 python tools/gemm_breakdown.py 4096 bf16 2>&1 | grep -v -i warn > gpurun_out/r02_gemm_breakdown_b4096_bf16.txt
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --value-only > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/gemm_traffic.csv python tools/gemm_breakdown.py 4096 bf16 > gpurun_out/ncu_traffic.log 2>&1
du -sh gpurun_out
