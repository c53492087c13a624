set +e
out=gpurun_out/r02_zl; mkdir -p $out
timeout 900 python -m pytest tests -m gpu -q -x 2>&1 | tail -2
timeout 200 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -1
timeout 300 python bench.py --no-extras > $out/bench.json 2> $out/bench.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$out/bench.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "value", d["value"], "render", d["roofline"]["kernel_ms"], "frac", d["roofline"]["frac"], "e2e", d["e2e"]["ms_per_step"], "floor", d["e2e"]["h2d_floor_ms"])
PY
