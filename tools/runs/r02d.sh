set +e
out=gpurun_out/r02_d; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -5 $out/pytest.log
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_full.json 2> $out/bench_full.err; echo "bench rc=$?"; tail -c 600 $out/bench_full.err
python - <<PY
import json
d = json.loads(open("$out/bench_full.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","launch","e2e","e2e_on_grid","roofline","roofline_cfg3","fwd_bwd","prob","cpu_baseline","splat_bwd_ms","ref_cuda_op","clocks"):
    print(k, json.dumps(d.get(k))[:900])
PY
timeout 300 python bench.py --impl reference --steps 20 --warmup 5 | cut -c1-400
