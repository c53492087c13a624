#!/bin/bash
# One GPU visit: parity tests, then A/B bench lines of the render-kernel variants (libraries built by
# tools/build_variant.py), optionally ncu captures.  Usage (through gpurun):
#   bash tools/gpu_round.sh <tag> "<variants>" [tests] [ncu] [full]
# A variant is `default`, the name of a library built by tools/build_variant.py (e.g. `ring6` after
# `python tools/build_variant.py ring6 -DGF_TILE_RING=6`), `st:<n>` for GF_B200_ST=<n>, and `<selector>@<library>` combines
# an environment selector with a variant library (e.g. `st:8@ring6`).
# Every non-default variant first runs the splat parity tests, then two bench lines.
set +e
tag=${1:-x}; variants=${2:-default}; shift 2
out=gpurun_out/$tag
mkdir -p $out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $out/smi.txt 2>&1
for what in "$@"; do
  if [ $what = tests ]; then
    timeout 1200 python -m pytest tests -m gpu -x -q > $out/pytest.log 2>&1
    echo "pytest rc=$?" >> $out/pytest.log
    tail -5 $out/pytest.log
  fi
done
for v in $variants; do
  unset GF_B200_LIB GF_B200_ST
  # a variant is <selector>[@<library>]: selector = default | render:<x> | st:<n> | <library>
  sel=${v%%@*}; lib=""
  if [ "$sel" != "$v" ]; then lib=${v#*@}; fi
  if [ ${sel:0:3} = st: ]; then export GF_B200_ST=${sel:3};            # supertile edge, e.g. st:8
  elif [ $sel != default ]; then lib=$sel; fi
  if [ -n "$lib" ]; then export GF_B200_LIB=$PWD/gaussianformer_b200/csrc/variants/libgf_b200_$lib.so; fi
  if [ $v != default ]; then
    timeout 600 python -m pytest tests/test_splat_gpu.py tests/test_cabi_gpu.py tests/test_batch_fused_gpu.py -m gpu -x -q 2>&1 | tail -1 | sed "s/^/$v parity: /"
  fi
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
unset GF_B200_LIB GF_B200_ST
for what in "$@"; do
  if [ $what = ncu ]; then
    timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $out/launches.csv \
        python bench.py --steps 6 --warmup 3 --no-extras > $out/ncu_launch.log 2>&1
    timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_tile -s 4 -c 2 -o $out/render_full -f \
        python bench.py --steps 6 --warmup 3 --no-extras > $out/ncu_full.log 2>&1
    ls -la $out | tail -5
  fi
  if [ $what = full ]; then
    timeout 600 python bench.py --steps 200 --warmup 20 > $out/bench_full.json 2> $out/bench_full.err
    tail -c 1500 $out/bench_full.json
  fi
done
