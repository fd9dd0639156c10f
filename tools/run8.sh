#!/bin/bash
# One 8-GPU box: topology, the GPU tests (with the multi-device ones), config 5 strong-sharded, the weak-scaling line with e2e.
N=${N:-8}
nvidia-smi topo -m 2>&1 | head -14 > gpurun_out/r2c_topo_${N}gpu.txt
echo "nodes online: $(cat /sys/devices/system/node/online)  nproc: $(nproc)" >> gpurun_out/r2c_topo_${N}gpu.txt
grep -i "allowed_list" /proc/self/status >> gpurun_out/r2c_topo_${N}gpu.txt
python -m pytest tests -m gpu -x -q > gpurun_out/r2c_gputest_${N}gpu.log 2>&1; tail -2 gpurun_out/r2c_gputest_${N}gpu.log
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --total-columns 16777216 --steps 10 --warmup 3 --e2e-steps 3 > gpurun_out/r2c_bench_strong_${N}gpu.json 2> gpurun_out/r2c_bench_strong_${N}gpu.err
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 3 --e2e-steps 3 > gpurun_out/r2c_bench_weak_${N}gpu.json 2> gpurun_out/r2c_bench_weak_${N}gpu.err
for f in gpurun_out/r2c_bench_strong_${N}gpu.json gpurun_out/r2c_bench_weak_${N}gpu.json; do python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    print(sys.argv[1], "value %.3e ms %.2f e2e %.3e (%.1f ms) numa %s" % (d["value"], d["ms_per_step"], d["e2e"]["value"], d["e2e"]["ms_per_step"], d["e2e"].get("numa")))
except Exception as e:
    print(sys.argv[1], "unreadable", e)
PY
done
cat gpurun_out/r2c_topo_${N}gpu.txt
