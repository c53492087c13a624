set +e
out=gpurun_out/r02_ze; mkdir -p $out
t0=$(date +%s)
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -3 $out/pytest.log
t1=$(date +%s); echo "pytest seconds $((t1-t0))"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -2
t2=$(date +%s)
timeout 900 python bench.py > $out/bench_default.json 2> $out/bench_default.err; echo "bench rc=$?"
t3=$(date +%s); echo "default bench seconds $((t3-t2))"
timeout 600 python bench.py --impl reference > $out/bench_reference.json 2> $out/bench_reference.err; echo "reference rc=$?"
t4=$(date +%s); echo "reference arm seconds $((t4-t3))"
python - <<PY
import json
d = json.loads(open("$out/bench_default.json").read().strip().splitlines()[-1])
for k in ("value","ms_per_step","steps","clocks","e2e","roofline","roofline_cfg3","fwd_bwd","prob","cpu_baseline"):
    print(k, json.dumps(d.get(k))[:500])
r = json.loads(open("$out/bench_reference.json").read().strip().splitlines()[-1])
print("reference", json.dumps(r)[:600])
print("same config", r.get("config") == d.get("config"), "steps", r.get("steps"), d.get("steps"))
PY
