set +e
out=gpurun_out/r02_w; mkdir -p $out

for rep in 1 2; do
timeout 900 python bench.py > $out/bench_$rep.json 2> $out/bench_$rep.err; echo "bench rc=$?"
python - <<PY
import json
d = json.loads(open("$out/bench_$rep.json").read().strip().splitlines()[-1])
print("ms/step", d["ms_per_step"], "render", d["roofline"]["kernel_ms"])
e = d["e2e"]; print("e2e", e["ms_per_step"], e["regions_ms_per_step"], "wall", e["wall_ms_per_step"], "floor", e["h2d_floor_ms"], "| on grid", d["e2e_on_grid"]["ms_per_step"])
print("fwd_bwd", d["fwd_bwd"]["ms_per_step"], d["fwd_bwd"]["fwd_ms"], "| prob", d["prob"]["fwd_ms"], d["prob"]["fwd_bwd_ms"], d["prob"]["roofs_bwd"]["fp32_frac"], "| cfg3", d["roofline_cfg3"]["kernel_ms"], d["roofline_cfg3"]["op_call_ms"])
PY
done
