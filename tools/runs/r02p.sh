set +e
out=gpurun_out/r02_p; mkdir -p $out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29521 bench.py --gpus 4 --steps 20 --warmup 5 > $out/bench_n4.json 2> $out/bench_n4.err; echo "bench rc=$?"; tail -c 400 $out/bench_n4.err
python - <<PY
import json
d = json.loads(open("$out/bench_n4.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","n_gpus","fwd_bwd","prob"):
    print(k, json.dumps(d.get(k))[:400])
print("e2e", d["e2e"]["ms_per_step"], d["e2e"]["h2d_floor_ms"])
PY
