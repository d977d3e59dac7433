N=$1
nvidia-smi -L | wc -l; nproc; free -g | head -2 | tail -1
tests/cpp/_build/test_multi --columns 64; echo rc=$?
timeout 400 python tools/bench_multi.py --reps 2 > gpurun_out/r2_multi_n$N.json 2> gpurun_out/r2_multi_n$N.err; cat gpurun_out/r2_multi_n$N.json | cut -c1-700
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 10 --warmup 3 2> gpurun_out/r2_bench_n$N.err | grep '^{' > gpurun_out/r2_bench_n$N.json; tail -2 gpurun_out/r2_bench_n$N.err | cut -c1-300
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n$N.json"))
print("c4 value", d["value"], "e2e", d["e2e"]["value"], "host_frac", d["e2e"]["host_frac"], "packed_sink", d["e2e_packed_sink"]["value"], d["e2e_packed_sink"]["link_frac"])
c=d["c5"]; print("c5 kernel", c["kernel_only"]["value"], "e2e", c["e2e"]["value"], c["e2e"].get("host_frac"), "packed", c["e2e_packed_sink"]["value"], c["e2e_packed_sink"].get("link_frac"), "parity", c["parity"]["ok_on_all_ranks"], c["parity"]["chunks_compared"])
print(d["probes"])
PY
