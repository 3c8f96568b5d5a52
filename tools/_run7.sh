set -x
timeout 300 python -m pytest tests/test_trainer_gpu.py -m gpu -q -x --timeout 280 2>&1 | grep -v Warning | tail -40 > gpurun_out/r02_traj_fail.txt
timeout 1200 python -m pytest tests -m gpu -q --timeout 900 --deselect tests/test_trainer_gpu.py::test_trainer_trajectory_matches_oracle_adam 2>&1 | tail -15 > gpurun_out/r02_gputests.txt
tail -5 gpurun_out/r02_gputests.txt
timeout 900 python bench.py --steps 10 --warmup 3 > gpurun_out/r02_bench_1gpu.json 2> gpurun_out/r02_bench_1gpu.err
tail -c 3000 gpurun_out/r02_bench_1gpu.json; tail -3 gpurun_out/r02_bench_1gpu.err
