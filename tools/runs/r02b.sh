set +e
out=gpurun_out/r02_b; mkdir -p $out
timeout 1500 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -25 $out/pytest.log
for b in 1 4; do GF_B200_LIB=$PWD/gaussianformer_b200/csrc/variants/libgf_b200_timing.so timeout 300 python tools/phase_timing.py gs25600_solid $b 2>&1 | tail -7; done
GF_B200_LIB=$PWD/gaussianformer_b200/csrc/variants/libgf_b200_timing.so timeout 300 python tools/phase_timing.py gs144000 1 2>&1 | tail -7
for rep in 1 2; do timeout 300 python bench.py --steps 200 --warmup 20 --no-extras > $out/bench_$rep.json 2> $out/bench_$rep.err; python - <<PY
import json
try:
    d = json.loads(open("$out/bench_$rep.json").read().strip().splitlines()[-1])
    print("default", $rep, "ms/step", round(d["ms_per_step"], 5), "render_ms", round(d["roofline"]["kernel_ms"], 5), "e2e_ms", round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("failed", e); print(open("$out/bench_$rep.err").read()[-2000:])
PY
done
