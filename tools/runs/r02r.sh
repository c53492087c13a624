set +e
out=gpurun_out/r02_r; mkdir -p $out
timeout 900 python -m pytest tests/test_splat_gpu.py tests/test_cabi_gpu.py tests/test_batch_fused_gpu.py tests/test_parity_full_gpu.py -m gpu -q -x 2>&1 | tail -2
for rep in 1 2; do
  echo "== window (default lib)"; timeout 300 python tools/time_render.py
  echo "== prev"; GF_B200_LIB=$PWD/gaussianformer_b200/csrc/variants/libgf_b200_prev.so timeout 300 python tools/time_render.py
done
