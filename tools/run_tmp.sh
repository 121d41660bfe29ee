python -m pytest tests/test_gpu_parity.py tests/test_fuzz.py tests/test_gpu_archive.py -m gpu -x -q 2>&1 | tail -3
for v in stock trip8; do
  if [ $v = stock ]; then unset DSRC_B200_LIB; else export DSRC_B200_LIB=$PWD/build_variants/libdsrc_$v.so; fi
  echo "== $v"; python tools/phase_prof.py 50000000 0 8192 2>&1 | grep -E "call|model_|rc_"
  DSRCGPU_SLOTS=1 python tools/phase_prof.py 6000000 0 8192 2>&1 | grep -E "model_|rc_"
done
unset DSRC_B200_LIB
