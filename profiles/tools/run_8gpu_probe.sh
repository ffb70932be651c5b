set -x
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
python profiles/tools/pcie_probe.py --json gpurun_out/pcie_ceiling.json
for n in 2 4 8; do $TR --nproc-per-node $n --master-port $((29520+n)) profiles/tools/pcie_probe.py --json gpurun_out/pcie_ceiling.json 2>/dev/null | tail -1; done
cp gpurun_out/pcie_ceiling.json profiles/pcie_ceiling.json
$TR --nproc-per-node 8 --master-port 29540 bench.py --gpus 8 --steps 3 --warmup 3 > gpurun_out/bench_cfg2_8gpu_r02.json 2> gpurun_out/bench_cfg2_8gpu_r02.err
tail -2 gpurun_out/bench_cfg2_8gpu_r02.err
python - <<'PY'
import json
d=json.load(open("gpurun_out/bench_cfg2_8gpu_r02.json"))
print("cfg2 x8", round(d["value"]), "e2e", round(d["e2e"]["value"]), "ceiling", d["e2e"].get("ceiling_fps"))
for k,v in (d.get("also") or {}).items(): print(k, round(v["value"]), v["intra_gpu_shards"], round(v["frac"],3), round(v["roofline"]["frac_sustained"],3))
PY
