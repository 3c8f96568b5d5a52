timeout 400 python -m pytest tests/test_em_gpu.py::test_full_size_properties_configs1 -m gpu -q --timeout 380 2>&1 | grep -E "^E  |assert|passed|failed" | head -12
for i in 1 2; do
timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('tail  ', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
VMM_TAIL_SPLIT=0 timeout 600 python bench.py --steps 20 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | python -c "import sys,json; d=json.loads(sys.stdin.read()); print('notail', round(d['value']), round(d['ms_per_step'],2), d['clocks']['sm_mhz'])"
done
