set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -2
timeout 300 python tools/gemm_breakdown.py 4096 bf16 2>&1 | grep -v -i warn | head -24 | tee gpurun_out/r02_gemm_breakdown_tail.txt
VMM_TAIL_SPLIT=0 timeout 300 python tools/gemm_breakdown.py 4096 bf16 2>&1 | grep -v -i warn | head -24 | tee gpurun_out/r02_gemm_breakdown_notail.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | cut -c1-200
VMM_TAIL_SPLIT=0 timeout 600 python bench.py --steps 10 --warmup 3 --no-extras --no-cpu-baseline 2>/dev/null | cut -c1-200
