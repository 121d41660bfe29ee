# developer aid (run under gpurun): isolated kernel times (one slot) of library variants on a 2.2 GB sample
#   bash tools/ab.sh <profile> <lib> [<lib> ...]     ("stock" = dsrc_b200/libdsrc_b200.so)
prof=$1; shift
for lib in "$@"; do
  if [ "$lib" = stock ]; then unset DSRC_B200_LIB; else export DSRC_B200_LIB=$PWD/build_variants/libdsrc_$lib.so; fi
  echo "== $lib (profile $prof)"
  DSRCGPU_SLOTS=1 python tools/phase_prof.py 6000000 $prof 8192 2>&1 | grep -E "call|preprocess|tags|model_quality|model_dna|rc_encode|parse|count|phase q.tab"
done
unset DSRC_B200_LIB
