for sc in strong weak; do
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29731 bench.py --gpus 4 --steps 10 --warmup 3 --scaling $sc 2>/dev/null | grep "^{" > gpurun_out/r02_bench_4gpu_${sc}_final.json
python -c "import json; d=json.load(open('gpurun_out/r02_bench_4gpu_${sc}_final.json')); print('$sc', round(d['value']), round(d['ms_per_step'],2), round(d['e2e']['value']), d['clocks']['sm_mhz'])"
done
