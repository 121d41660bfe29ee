python -m pytest tests/test_gpu_parity.py tests/test_gpu_archive.py tests/test_gpu_shim.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2l_parity.log; cat gpurun_out/r2l_parity.log
run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --no-decode --no-serial --no-extras $EXTRA > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err; python -c "import json;d=json.load(open('gpurun_out/sw_$name.json'));print('$name',round(d['value']),{k:round(v) for k,v in d['roofline']['kernel_ms_per_step'].items()})"; }
run g1 DSRCGPU_RC_GROUP=1
run g2 DSRCGPU_RC_GROUP=2
EXTRA="--inflight 6144" run g3s4 DSRCGPU_RC_GROUP=3 DSRCGPU_SLOTS=4
EXTRA="--inflight 6144" run g2s4 DSRCGPU_RC_GROUP=2 DSRCGPU_SLOTS=4
EXTRA="--inflight 4096" run g4s6 DSRCGPU_RC_GROUP=4 DSRCGPU_SLOTS=6
EXTRA="--inflight 4096" run g3s5 DSRCGPU_RC_GROUP=3 DSRCGPU_SLOTS=5
DSRCGPU_TIMELINE=gpurun_out/tl_g2.csv python tools/phase_prof.py 50000000 0 8192 2>&1 | grep call
