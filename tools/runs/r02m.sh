set +e
out=gpurun_out/r02_m; mkdir -p $out
# launch list of the bench step (cold-cache, serialised) and of the config legs
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $out/launches_fwd.csv python bench.py --steps 6 --warmup 3 --no-extras --no-graph > $out/ncu_l1.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_bwd_b4.csv python tools/time_bwd.py gs25600_solid 4 > $out/ncu_l2.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_bwd_prob.csv python tools/time_bwd.py prob_gs6400 1 > $out/ncu_l3.log 2>&1
# full captures
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_tile -s 6 -c 1 -o $out/render_full -f python bench.py --steps 6 --warmup 3 --no-extras --no-graph > $out/ncu_f1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:backward_bin_kernel -s 2 -c 1 -o $out/bwd_bin_full -f python tools/time_bwd.py gs25600_solid 1 > $out/ncu_f2.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:backward_bin_kernel -s 2 -c 1 -o $out/bwd_bin_prob_full -f python tools/time_bwd.py prob_gs6400 1 > $out/ncu_f3.log 2>&1
ls -la $out
timeout 600 python tools/daf_experiments.py > $out/daf.json 2> $out/daf.err; python -c "
import json; d=json.load(open('$out/daf.json'))
for k,v in d.items(): print(k, v)"
