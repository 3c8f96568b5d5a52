set -x
timeout 200 python tools/infer_profile.py 16384 3 10 2>&1 | grep -v -i warn | tee gpurun_out/r02_infer_profile.txt
timeout 300 python bench.py --steps 2 --warmup 3 --value-only > gpurun_out/r02_bench_value_only.json 2> gpurun_out/r02_bench_value_only.err || exit 1
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --value-only > gpurun_out/r02_ncu_launches.log 2>&1
tail -2 gpurun_out/r02_ncu_launches.log | cut -c1-300
timeout 300 python tools/gemm_breakdown.py 4096 bf16 > gpurun_out/plain_4096.log 2>&1 || exit 1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/gemm_traffic.csv python tools/gemm_breakdown.py 4096 bf16 > gpurun_out/ncu_traffic.log 2>&1
tail -2 gpurun_out/ncu_traffic.log | cut -c1-300
