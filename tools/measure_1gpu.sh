set -x
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_b4096_bf16.json 2> gpurun_out/r02_bench_b4096_bf16.err
tail -c 300 gpurun_out/r02_bench_b4096_bf16.json
timeout 300 python bench.py --steps 10 --warmup 3 --precision bf16_b8 --no-extras --no-cpu-baseline > gpurun_out/r02_bench_b4096_bf16_b8.json 2>/dev/null
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_cpu.json 2>/dev/null
timeout 900 python bench.py --workload infer_sweep --objects 1000000 > gpurun_out/r02_infer_sweep_1gpu_1M.json 2> gpurun_out/r02_infer_sweep_1gpu.err
tail -c 300 gpurun_out/r02_infer_sweep_1gpu_1M.json
du -sh gpurun_out
