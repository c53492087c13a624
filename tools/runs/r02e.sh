set +e
out=gpurun_out/r02_e; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 20 --warmup 5 > $out/bench_n2.json 2> $out/bench_n2.err; echo "bench rc=$?"; tail -c 800 $out/bench_n2.err
python - <<PY
import json
d = json.loads(open("$out/bench_n2.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","n_gpus","e2e","roofline","fwd_bwd","prob"):
    print(k, json.dumps(d.get(k))[:700])
PY
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --impl reference --gpus 2 --steps 5 --warmup 2 2>/dev/null | tail -1 | cut -c1-300
