set -x
# ncu: launch list of the bench command, contraction traffic, full sets of the E-step launch and the streaming stages
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 3 --value-only > gpurun_out/r02_ncu_launches.log 2>&1
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:gemm_tcgen05 --csv --log-file gpurun_out/gemm_traffic.csv python tools/gemm_breakdown.py 4096 bf16 > gpurun_out/ncu_traffic.log 2>&1
timeout 600 python tools/estep_only.py 1024 > gpurun_out/plain_estep.log 2>&1 && timeout 900 ncu --set full --clock-control none --import-source on -k regex:gemm_tcgen05 -s 3 -c 2 -o gpurun_out/r02_prof_estep python tools/estep_only.py 1024 > gpurun_out/ncu_estep.log 2>&1
timeout 600 python tools/gemm_breakdown.py 1024 bf16 > gpurun_out/plain.log 2>&1 && timeout 1500 ncu --set full --clock-control none -k regex:'enc1_backward_fused|lnw4k_backward|estep_dpred|lnw_forward|ln1k_forward|lnw_backward_rows|gru_|ln_backward_cols4' -s 220 -c 60 -o /tmp/r02_prof_stream python tools/gemm_breakdown.py 1024 bf16 > gpurun_out/ncu_stream.log 2>&1
tail -2 gpurun_out/ncu_stream.log
python tools/ncu_table.py gpurun_out/r02_prof_estep.ncu-rep /tmp/r02_prof_stream.ncu-rep > gpurun_out/r02_estep_and_stream_ncu.txt
head -5 gpurun_out/r02_estep_and_stream_ncu.txt | cut -c1-200
du -sh gpurun_out
