set +e
out=gpurun_out/r02_i; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -4 $out/pytest.log
for cfg in gs25600_solid prob_gs6400 gs144000; do for b in 1 4; do
  timeout 120 python tools/time_bwd.py $cfg $b 2>&1 | tail -1
done; done
timeout 600 python bench.py --steps 20 --warmup 5 > $out/bench_full.json 2> $out/bench_full.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$out/bench_full.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","e2e","roofline","roofline_cfg3","fwd_bwd","prob","splat_bwd_ms"):
    print(k, json.dumps(d.get(k))[:600])
PY
