set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -4 | tee gpurun_out/r02_gputests.txt
timeout 600 python tools/cfg3_bench.py 256 bf16 2>&1 | grep -v -i warn | tail -3 | tee gpurun_out/r02_cfg3_bench.json
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
