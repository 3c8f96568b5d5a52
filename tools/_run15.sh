set -x
timeout 300 python -m pytest tests/test_gemm_gpu.py -m gpu -q -x --timeout 280 -k estep 2>&1 | tail -3
timeout 200 python tools/estep_only.py 4096 all 2>&1 | grep estep | tee gpurun_out/r02_estep_only_v3.txt
VMM_B200_LIB=$PWD/multi_view_sort_b200/build/libvmm_b200_bufs3.so timeout 200 python tools/estep_only.py 4096 all 2>&1 | grep estep | tee gpurun_out/r02_estep_only_v3_bufs3.txt
timeout 600 python tools/estep_only.py 1024 > gpurun_out/plain_estep.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 2 -o gpurun_out/r02_prof_estep python tools/estep_only.py 1024 > gpurun_out/ncu_estep.log 2>&1
python tools/ncu_table.py gpurun_out/r02_prof_estep.ncu-rep | tee gpurun_out/r02_estep_ncu.txt | cut -c1-250
