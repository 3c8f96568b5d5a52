set -x
nvidia-smi -L
timeout 600 python -m pytest tests/test_dist_gpu.py tests/test_trainer_gpu.py -m gpu -q -x --timeout 580 2>&1 | tail -8
PORT=29511
for sc in weak strong; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port $PORT bench.py --gpus 2 --steps 5 --warmup 3 --scaling $sc > gpurun_out/r02_bench_2gpu_$sc.json 2> gpurun_out/r02_bench_2gpu_$sc.err
tail -c 600 gpurun_out/r02_bench_2gpu_$sc.json; tail -3 gpurun_out/r02_bench_2gpu_$sc.err
PORT=$((PORT+1))
done
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29520 bench.py --gpus 2 --workload infer_sweep --objects 200000 > gpurun_out/r02_infer_sweep_2gpu_200k.json 2> gpurun_out/r02_infer_sweep_2gpu.err
tail -c 1500 gpurun_out/r02_infer_sweep_2gpu_200k.json; tail -3 gpurun_out/r02_infer_sweep_2gpu.err
