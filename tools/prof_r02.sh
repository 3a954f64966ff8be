#!/bin/bash
# Round-2 evidence on one B200: the bench line, the ncu launch list of the bench's own workload, and --set full captures of
# K3 (the two launches of a step: 256-hypothesis CTAs, then the shortest pairs as 128-hypothesis CTAs) and of the matcher's three kernels on shorter sequences (ncu saves and
# restores device memory per replay pass).  Summaries: tools/ncu_extract.py.
set -x
mkdir -p gpurun_out/r02
python bench.py --steps 20 --warmup 3 > gpurun_out/r02/bench_full.json 2> gpurun_out/r02/bench_full.err
python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02/bench_reference_arm.json 2>> gpurun_out/r02/bench_full.err
python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02/plain_l.log 2>&1 && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r02/launches4541.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02/ncu_l.log 2>&1
python bench.py --frames 1025 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02/plain_k3.log 2>&1 && \
timeout 400 ncu --set full --clock-control none --import-source on -k regex:twophase -s 2 -c 2 -f -o gpurun_out/r02/k3_twophase_full python bench.py --frames 1025 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/r02/ncu_k3.log 2>&1
python tools/match_parts.py 1025 > gpurun_out/r02/plain_m.log 2>&1 && \
timeout 300 ncu --set full --clock-control none --import-source on -k regex:"match_tc_kernel|lead_kernel|compact_candidates" -s 9 -c 3 -f -o gpurun_out/r02/match_full python tools/match_parts.py 1025 > gpurun_out/r02/ncu_m.log 2>&1
python tools/match_sweep.py > gpurun_out/r02/match_sweep.jsonl 2>&1
python tools/match_parts.py > gpurun_out/r02/match_parts.json 2>&1
python tools/stress_bench.py --frames 223 > gpurun_out/r02/stress_bench_1gpu.json 2>&1
for n in 1000 330; do python tools/pnp_bench.py --n $n; done > gpurun_out/r02/pnp_bench_configs2.jsonl 2>&1
ls -la gpurun_out/r02
