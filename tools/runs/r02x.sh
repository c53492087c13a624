set +e
out=gpurun_out/r02_zf; mkdir -p $out
# final kernels: launch list of the bench step (cold-cache, serialised; direct launches so that ncu sees the kernels) and
# one --set full capture of the render kernel at 1 and 4 samples per launch
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 120 --csv --log-file $out/launches_fwd.csv python bench.py --steps 6 --warmup 3 --no-extras --no-graph > $out/ncu_l1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_tile -s 6 -c 1 -o $out/render_full -f python bench.py --steps 6 --warmup 3 --no-extras --no-graph > $out/ncu_f1.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:render_tile -s 3 -c 1 -o $out/render_b4_full -f python tools/time_render.py gs25600_solid:4 > $out/ncu_f2.log 2>&1
ls -la $out
