set -x
nvidia-smi -L | wc -l
PORT=29611
for sc in weak strong; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 8 --steps 10 --warmup 3 --scaling $sc 2> gpurun_out/r02_bench_8gpu_$sc.err | grep "^{" > gpurun_out/r02_bench_8gpu_$sc.json
cut -c1-330 gpurun_out/r02_bench_8gpu_$sc.json; tail -2 gpurun_out/r02_bench_8gpu_$sc.err
PORT=$((PORT+1))
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29620 bench.py --gpus 8 --workload infer_sweep --objects 1000000 2> gpurun_out/r02_infer_sweep_8gpu.err | grep "^{" > gpurun_out/r02_infer_sweep_8gpu_1M.json
cut -c1-300 gpurun_out/r02_infer_sweep_8gpu_1M.json; tail -2 gpurun_out/r02_infer_sweep_8gpu.err
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29630 bench.py --gpus 4 --steps 10 --warmup 3 --scaling strong 2>/dev/null | grep "^{" > gpurun_out/r02_bench_4gpu_strong.json
cut -c1-200 gpurun_out/r02_bench_4gpu_strong.json
