set +e
out=gpurun_out/r02_g; mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file $out/launches_bwd.csv python tools/time_bwd.py gs25600_solid 1 > $out/l.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:backward_bin_kernel -s 2 -c 1 -o $out/bwd_bin_full -f python tools/time_bwd.py gs25600_solid 1 > $out/n.log 2>&1
ls -la $out
