python -m pytest tests/test_gpu_multi_rank.py -m gpu -x -q 2>&1 | tail -1
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 50 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_r2f_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 3 --knets 6 --batch 2097152 2>/dev/null | tail -1 > gpurun_out/bench_r2f_cfg4_n8.json
python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 4 --steps 50 --warmup 5 2>/dev/null | tail -1 > gpurun_out/bench_r2f_n4.json
python - <<'P'
import json
for f in ["bench_r2f_n8", "bench_r2f_cfg4_n8", "bench_r2f_n4"]:
    d = json.loads(open("gpurun_out/" + f + ".json").read())
    print(f, d["value"] / 1e6, d["ms_per_step"], d["config"]["exchange"], d["parity"]["ok"], d["e2e"]["value"] / 1e6, d.get("gather_look_ahead", {}).get("ok"))
P
