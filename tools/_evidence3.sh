set -x
timeout 900 python -m pytest tests -q -m gpu 2>&1 | tail -3
python __graft_entry__.py --smoke 2>&1 | tail -2
python tools/bench_world.py --no-cpu > gpurun_out/r2_bench_world.json 2> gpurun_out/r2_world.err; tail -1 gpurun_out/r2_world.err
python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-c5 --no-c23 --pipeline staged > gpurun_out/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches_bench_c4.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-e2e --no-c5 --no-c23 --pipeline staged > gpurun_out/ncu_l.log 2>&1
python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-c5 --no-c23 > gpurun_out/plain_b.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_mesh_region|k_heights|k_count_region|k_region_fused' -s 12 -c 6 -f -o gpurun_out/r2_c4_full python bench.py --steps 3 --warmup 3 --no-cpu --no-e2e --no-c5 --no-c23 > gpurun_out/ncu_c4.log 2>&1
python tools/bench_kernels.py --scale 4 --reps 3 > gpurun_out/plain_k.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:'k_fill|k_pack|k_count_border|k_rle_encode|k_rle_decode' -c 8 -f -o gpurun_out/r2_c23_full python tools/bench_kernels.py --scale 4 --reps 3 > gpurun_out/ncu_k.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:'k_emit_ids' -c 1 -f -o gpurun_out/r2_c3emit_full python tools/bench_kernels.py --scale 4 --reps 3 > gpurun_out/ncu_k2.log 2>&1
bash tools/_c3small.sh > /dev/null 2>&1
python tools/bench_kernels.py --scale 1 > gpurun_out/r2_bench_kernels.json 2>/dev/null; python tools/bench_kernels.py --scale 4 > gpurun_out/r2_bench_kernels_x4.json 2>/dev/null
ls -la gpurun_out/r2_c4_full.ncu-rep gpurun_out/r2_c23_full.ncu-rep gpurun_out/r2_c3emit_full.ncu-rep gpurun_out/r2_launches_bench_c4.csv
