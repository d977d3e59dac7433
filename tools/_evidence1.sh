set -x
python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-c5 > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_mesh_region|k_heights|k_count_region|k_region_fused' -s 12 -c 6 -o gpurun_out/r2_c4_full python bench.py --steps 5 --warmup 3 --no-cpu --no-e2e --no-c5 > gpurun_out/ncu_c4.log 2>&1
python tools/bench_kernels.py --scale 4 --reps 3 > gpurun_out/plain_k.log 2>&1 && \
ncu --set full --clock-control none -k regex:'k_fill|k_pack|k_count_border|k_emit_ids|k_rle_encode|k_rle_decode' -c 14 -o gpurun_out/r2_c23_full python tools/bench_kernels.py --scale 4 --reps 3 > gpurun_out/ncu_k.log 2>&1
ls -la gpurun_out/*.ncu-rep | tail -3
