for v in stock rcring walk48; do
  if [ $v = stock ]; then unset DSRC_B200_LIB; else export DSRC_B200_LIB=$PWD/build_variants/libdsrc_$v.so; fi
  for c in 1 0; do echo "== $v carve=$c"; DSRCGPU_RC_CARVEOUT=$c python tools/phase_prof.py 50000000 0 8192 2>&1 | grep -E "call|model_quality|rc_"; done
done
unset DSRC_B200_LIB
