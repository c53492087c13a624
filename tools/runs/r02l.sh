set +e
out=gpurun_out/r02_l; mkdir -p $out
timeout 900 python -m pytest tests/test_splat_gpu.py tests/test_cabi_gpu.py tests/test_batch_fused_gpu.py tests/test_daf_gpu.py -m gpu -q -x 2>&1 | tail -2
GF_B200_LIB=$PWD/gaussianformer_b200/csrc/variants/libgf_b200_timing.so timeout 300 python tools/phase_timing.py gs25600_solid 1 2>&1 | tail -12
for rep in 1 2; do timeout 300 python bench.py --steps 200 --warmup 20 --no-extras > $out/bench_$rep.json 2> $out/bench_$rep.err; python - <<PY
import json
d = json.loads(open("$out/bench_$rep.json").read().strip().splitlines()[-1])
print("default", $rep, "ms/step", round(d["ms_per_step"], 5), "render_ms", round(d["roofline"]["kernel_ms"], 5), "e2e_ms", round(d["e2e"]["ms_per_step"], 4))
PY
done
timeout 600 python tools/daf_experiments.py > $out/daf.json 2> $out/daf.err; python -c "
import json; d=json.load(open('$out/daf.json'))
for k,v in d.items(): print(k, v)"
