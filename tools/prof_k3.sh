# ncu launch list + --set full capture of the PnP kernel (K3) inside a short bench run; summaries via tools/ncu_extract.py
set -x
mkdir -p gpurun_out/p
python bench.py --frames 1025 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/p/bench1025.json 2> gpurun_out/p/bench1025.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p/launches1025.csv python bench.py --frames 1025 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/p/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pnp_gn -s 1 -c 1 -f -o gpurun_out/p/k3_full python bench.py --frames 1025 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/p/ncu_k3.log 2>&1
