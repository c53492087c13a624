set +e
out=gpurun_out/r02_c; mkdir -p $out
timeout 2400 python -m pytest tests -m gpu -q > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -60 $out/pytest.log
