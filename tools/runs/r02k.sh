set +e
out=gpurun_out/r02_k; mkdir -p $out
timeout 600 python tools/daf_experiments.py > $out/daf.json 2> $out/daf.err; echo rc=$?; tail -c 1500 $out/daf.err; python -c "
import json; d=json.load(open('$out/daf.json'))
for k,v in d.items(): print(k, v)"
