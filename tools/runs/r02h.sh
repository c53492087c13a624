set +e
out=gpurun_out/r02_h; mkdir -p $out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 80 --csv --log-file $out/launches_bwd.csv python tools/time_bwd.py gs25600_solid 1 > $out/l.log 2>&1
for i in 1 2 3; do timeout 120 python tools/time_bwd.py gs25600_solid 1 2>&1 | tail -1; done
GF_B200_PDL=0 timeout 120 python tools/time_bwd.py gs25600_solid 1 2>&1 | tail -1
