# Launch list of the bench's own workload (4 541 frames, the <2> instance of the PnP kernel) and a
# --set full capture of that instance on a shorter sequence (ncu saves and restores device memory per
# replay pass: 10.8 GB of frames would take minutes); summaries via tools/ncu_extract.py
set -x
mkdir -p gpurun_out/pf
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/pf/launches4541.csv python bench.py --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/pf/ncu_l.log 2>&1
MV_PNP_GPW=2 timeout 150 ncu --set full --clock-control none --import-source on -k regex:pnp_gn -s 1 -c 1 -f -o gpurun_out/pf/k3_gpw2_full python bench.py --frames 2049 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/pf/ncu_k3.log 2>&1
ls -la gpurun_out/pf
