set -x
timeout 900 python -m pytest tests/test_em_gpu.py tests/test_gemm_gpu.py -m gpu -q -x --timeout 900 2>&1 | tail -8
for B in 8 32; do timeout 400 python tools/relu_flips.py $B "bf16:enc2:a,dec2:a,dec3:s" "bf16:enc2:w,dec2:w,dec3:s" "bf16:enc2:s,dec2:s,dec3:s" "bf16:xw1:s,g1:s,enc2:s,gi:s,gh:s,dec1:s,dec2:s,dec3:s" "bf16:g1:a,enc2:a,gi:a,gh:a,dec1:a,dec2:a,dec3:s" "bf16:g1:w,enc2:w,gi:w,gh:w,dec1:w,dec2:w,dec3:s" 2>&1 | grep -v Warning; done
