# Round-1 profile set (run on a B200 via gpurun; summaries are condensed into profiles/r01 by tools/ncu_extract.py)
set -x
mkdir -p gpurun_out/p
python tools/ffma_peak.py > gpurun_out/p/ffma_peak.txt 2>&1
python - > gpurun_out/p/pcie_h2d.txt 2>&1 <<'PY'
import torch
h = torch.empty(1 << 30, dtype=torch.uint8, pin_memory=True); d = torch.empty(1 << 30, dtype=torch.uint8, device="cuda")
for _ in range(2): d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(5): d.copy_(h, non_blocking=True)
e1.record(); torch.cuda.synchronize()
print("pinned H2D cudaMemcpyAsync 1 GiB x5: %.1f GB/s" % (5 * (1 << 30) / (e0.elapsed_time(e1) * 1e-3) / 1e9))
PY
python tools/match_sweep.py > gpurun_out/p/match_sweep.jsonl 2> gpurun_out/p/match_sweep.err
python bench.py --frames 1025 --steps 2 --warmup 1 --no-e2e --no-cpu > gpurun_out/p/bench1025.json 2> gpurun_out/p/bench1025.err || exit 1
ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/p/launches1025.csv python bench.py --frames 1025 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/p/ncu_l.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:pnp_gn -s 1 -c 1 -f -o gpurun_out/p/k3_full python bench.py --frames 1025 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/p/ncu_k3.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:match_tc -s 1 -c 1 -f -o gpurun_out/p/tc_full python bench.py --frames 1025 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/p/ncu_tc.log 2>&1
ncu --set full --clock-control none --import-source on -k regex:softmax_cells -s 1 -c 1 -f -o gpurun_out/p/k0a_full python bench.py --frames 1025 --steps 1 --warmup 1 --no-e2e --no-cpu > gpurun_out/p/ncu_k0a.log 2>&1
ls -la gpurun_out/p
