#!/bin/bash
# Round-2 multi-GPU measurements on one N-GPU box (run under gpurun --gpus N): writes JSON lines to gpurun_out/.
N=${1:-8}
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29561"
mkdir -p gpurun_out
timeout 600 $TR bench.py --gpus $N --steps 20 --warmup 5 2>gpurun_out/r2_bench_${N}gpu.err | tail -1 > gpurun_out/r2_bench_${N}gpu_peer.json; echo "bench rc=$?"
timeout 400 $TR tools/cfg34_run.py 2>gpurun_out/r2_cfg34_${N}gpu.err | tail -1 > gpurun_out/r2_cfg34_${N}gpu.json; echo "cfg34 rc=$?"
timeout 400 $TR tools/cfg5_run.py 200 2>gpurun_out/r2_cfg5_${N}gpu.err | tail -1 > gpurun_out/r2_cfg5_${N}gpu_sharded.json; echo "cfg5 rc=$?"
for f in gpurun_out/r2_bench_${N}gpu_peer.json gpurun_out/r2_cfg34_${N}gpu.json gpurun_out/r2_cfg5_${N}gpu_sharded.json; do echo "== $f"; head -c 1500 $f; echo; done
