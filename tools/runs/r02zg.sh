set +e
timeout 900 python -m pytest tests/test_splat_gpu.py tests/test_cabi_gpu.py tests/test_batch_fused_gpu.py tests/test_parity_full_gpu.py -m gpu -q -x 2>&1 | tail -2
for rep in 1 2; do
for v in default prev; do
  unset GF_B200_LIB
  if [ $v != default ]; then export GF_B200_LIB=$PWD/gaussianformer_b200/csrc/variants/libgf_b200_$v.so; fi
  echo "== $v"
  for c in gs25600_solid:1 gs25600_solid:4 prob_gs6400:1 gs144000:1; do timeout 120 python tools/time_bwd.py ${c%%:*} ${c##*:} 2>&1 | tail -1; done
done; done
