set +e
out=gpurun_out/r02_f; mkdir -p $out
timeout 900 python -m pytest tests/test_splat_gpu.py tests/test_cabi_gpu.py tests/test_batch_fused_gpu.py tests/test_parity_full_gpu.py -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -30 $out/pytest.log
for cfg in gs25600_solid prob_gs6400 gs144000; do for b in 1 4; do
  timeout 120 python tools/time_bwd.py $cfg $b 2>&1 | tail -1
  GF_B200_BWD=gauss timeout 120 python tools/time_bwd.py $cfg $b 2>&1 | tail -1
done; done
