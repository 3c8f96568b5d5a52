set -x
timeout 1500 python -m pytest tests -m gpu -q --timeout 900 2>&1 | tail -6 | tee gpurun_out/r02_gputests.txt
timeout 300 python tools/gemm_breakdown.py 4096 bf16 2>&1 | grep -v -i warn | tee gpurun_out/r02_gemm_breakdown_bf16.txt
timeout 600 python bench.py --steps 10 --warmup 3 --no-extras > gpurun_out/r02_bench_1gpu_v2.json 2> gpurun_out/r02_bench_1gpu_v2.err
python - <<'PY'
import json
for line in open("gpurun_out/r02_bench_1gpu_v2.json"):
    if line.startswith("{"):
        d=json.loads(line); r=d["roofline"]
        print(round(d["value"]), d["ms_per_step"], round(d["e2e"]["value"]), d["clocks"], {k:r[k] for k in ("frac","kernel_ms_per_step","kernel_share_of_step","executed_frac","step_frac")})
PY
