set +e
for rep in 1 2; do
for v in default va vb vc; do
  unset GF_B200_LIB
  if [ $v != default ]; then export GF_B200_LIB=$PWD/gaussianformer_b200/csrc/variants/libgf_b200_$v.so; fi
  echo "== $v"; timeout 200 python tools/time_render.py gs25600_solid:1 gs144000:1
done; done
