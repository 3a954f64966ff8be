set -x
mkdir -p gpurun_out/p
python tools/ffma_peak.py > gpurun_out/p/ffma_peak.txt 2>&1
python tools/match_sweep.py > gpurun_out/p/match_sweep.jsonl 2> gpurun_out/p/match_sweep.err
python bench.py --frames 513 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/p/bench513.json 2> gpurun_out/p/bench513.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p/launches513.csv python bench.py --frames 513 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/p/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pnp_gn -s 1 -c 1 -o gpurun_out/p/k3_full python bench.py --frames 513 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/p/ncu_k3.log 2>&1
ls -la gpurun_out/p
