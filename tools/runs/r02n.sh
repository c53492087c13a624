set +e
out=gpurun_out/r02_n; mkdir -p $out
timeout 600 python tests/golden/make_golden_ref.py gpurun_out/golden 2>&1 | tail -3
cp gpurun_out/golden/ref_splat_base_mid.npz tests/golden/ 2>/dev/null
timeout 1500 python -m pytest tests -m gpu -q -x > $out/pytest.log 2>&1; echo "pytest rc=$?" >> $out/pytest.log; tail -4 $out/pytest.log
for cfg in gs25600_solid prob_gs6400 gs144000; do for b in 1 4; do timeout 120 python tools/time_bwd.py $cfg $b 2>&1 | tail -1; done; done
timeout 600 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
