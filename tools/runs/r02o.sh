set +e
out=gpurun_out/r02_o; mkdir -p $out
for v in stubacc; do
export GF_B200_LIB=$PWD/gaussianformer_b200/csrc/variants/libgf_b200_$v.so
timeout 300 python bench.py --steps 100 --warmup 10 --no-extras > $out/bench_$v.json 2> $out/bench_$v.err; python - <<PY
import json
d = json.loads(open("$out/bench_$v.json").read().strip().splitlines()[-1])
print("$v", "ms/step", round(d["ms_per_step"], 5), "render_ms", round(d["roofline"]["kernel_ms"], 5))
PY
done
