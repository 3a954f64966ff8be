#!/bin/bash
# round-2 GPU call 1: parity (incl. the never-run checks), int8 peak, K3 two-phase vs streaming, bench line
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,clocks.sm --format=csv > gpurun_out/c1_smi.txt 2>&1
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/c1_pytest.log 2>&1; echo "pytest exit $?" >> gpurun_out/c1_pytest.log
tail -5 gpurun_out/c1_pytest.log
timeout 120 python tools/int8_peak.py --out gpurun_out/int8_peak.json > gpurun_out/c1_int8.log 2>&1; tail -2 gpurun_out/c1_int8.log
timeout 300 python tools/k3_ab.py MV_PNP_STREAM=1 MV_PNP_GPW=1 > gpurun_out/c1_k3ab.jsonl 2>&1; cat gpurun_out/c1_k3ab.jsonl
timeout 600 python bench.py --steps 5 --warmup 3 > gpurun_out/c1_bench.json 2> gpurun_out/c1_bench.err; tail -c 3000 gpurun_out/c1_bench.json; tail -3 gpurun_out/c1_bench.err
