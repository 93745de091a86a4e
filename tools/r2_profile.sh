#!/bin/bash
# Round-2 single-GPU captures (run under gpurun): bench line, ncu launch list of the same command, ncu --set full of
# the step kernel, cfg1 pipeline, smoke.  Everything lands in gpurun_out/.
mkdir -p gpurun_out
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
timeout 600 python bench.py --steps 20 --warmup 5 > gpurun_out/r2_bench_1gpu.json 2> gpurun_out/r2_bench_1gpu.err; echo "bench rc=$?"
timeout 300 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r2_bench_reference.json 2>/dev/null; echo "ref rc=$?"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_train_launches.csv python bench.py --steps 20 --warmup 5 --skip-cpu --skip-extras > gpurun_out/ncu_launch.log 2>&1; echo "launch list rc=$?"
timeout 600 ncu --set full --import-source on --clock-control none -k regex:chunk_kernel -s 3 -c 1 -o gpurun_out/r2_chunk_replay -f python tools/train_probe.py replay 512 3 > gpurun_out/ncu_chunk.log 2>&1; echo "ncu full rc=$?"
timeout 600 python tools/cfg1_pipeline.py 7000000 /tmp/cfg1 > gpurun_out/r2_cfg1_pipeline.json 2> gpurun_out/r2_cfg1.err; echo "cfg1 rc=$?"; tail -c 1500 gpurun_out/r2_cfg1_pipeline.json
