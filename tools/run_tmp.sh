python -m pytest tests/test_gpu_parity.py tests/test_gpu_archive.py tests/test_gpu_decode.py -m gpu -x -q 2>&1 | tail -3 > gpurun_out/r2k_parity.log; cat gpurun_out/r2k_parity.log
bash tools/ab.sh 0 stock 2>&1 | tee gpurun_out/r2k_ab.log
run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --no-decode --no-serial --no-extras > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err; python -c "import json;d=json.load(open('gpurun_out/sw_$name.json'));print('$name',round(d['value']),{k:round(v) for k,v in d['roofline']['kernel_ms_per_step'].items()})"; }
run walk_p2r0 DSRCGPU_PSERIAL=2 DSRCGPU_RSTREAM=0
run walk_p0r0 DSRCGPU_PSERIAL=0 DSRCGPU_RSTREAM=0
run walk_p2r1 DSRCGPU_PSERIAL=2 DSRCGPU_RSTREAM=1
