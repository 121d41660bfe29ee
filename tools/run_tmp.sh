python -m pytest tests/test_gpu_parity.py tests/test_gpu_archive.py tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2m_parity.log; cat gpurun_out/r2m_parity.log
bash tools/ab.sh 0 stock 2>&1 | grep -v phase | tee gpurun_out/r2m_ab.log
run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --no-decode --no-serial --no-extras $EXTRA > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err; python -c "import json;d=json.load(open('gpurun_out/sw_$name.json'));print('$name',round(d['value']),{k:round(v) for k,v in d['roofline']['kernel_ms_per_step'].items()})"; }
run dw_g2 DSRCGPU_RC_GROUP=2
run dw_g2p0 DSRCGPU_RC_GROUP=2 DSRCGPU_PSERIAL=0
