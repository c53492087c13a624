#!/bin/bash
# One GPU visit: parity tests, then A/B bench lines of the render-kernel variants, then the side measurements.
# Usage (through gpurun): bash tools/gpu_round.sh <tag>
set +e
tag=${1:-x}
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1
echo "pytest rc=$?" >> $out/pytest.log
tail -5 $out/pytest.log
for v in default walk0 b64; do
  if [ $v = default ]; then unset GF_B200_LIB; else export GF_B200_LIB=$PWD/gaussianformer_b200/csrc/variants/libgf_b200_$v.so; fi
  for rep in 1 2; do
    timeout 300 python bench.py --steps 200 --warmup 20 --no-extras > $out/bench_${v}_$rep.json 2> $out/bench_${v}_$rep.err
    python - <<PY
import json
try:
    d = json.loads(open("$out/bench_${v}_$rep.json").read().strip().splitlines()[-1])
    print("$v", $rep, "ms/step", round(d["ms_per_step"], 5), "render_ms", round(d["roofline"]["kernel_ms"], 5), "e2e_ms", round(d["e2e"]["ms_per_step"], 4))
except Exception as e:
    print("$v", $rep, "failed", e)
PY
  done
done
unset GF_B200_LIB
timeout 600 python bench.py --steps 200 --warmup 20 > $out/bench_full.json 2> $out/bench_full.err
tail -c 3000 $out/bench_full.json
