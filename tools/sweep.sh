# developer aid: scheduler parameter sweep on the bench workload (run under gpurun)
run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --no-decode $EXTRA > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err; python -c "import json;d=json.load(open('gpurun_out/sw_$name.json'));print('$name',round(d['value']),{k:round(v) for k,v in d['roofline']['kernel_ms_per_step'].items()})"; }
python -m pytest tests/test_gpu_parity.py -m gpu -x -q 2>&1 | tail -1
run s1_c740 DSRCGPU_SLOTS=1 DSRCGPU_MODEL_CTAS=740
run s3_c740 DSRCGPU_SLOTS=3 DSRCGPU_MODEL_CTAS=740
run s4_c740 DSRCGPU_SLOTS=4 DSRCGPU_MODEL_CTAS=740
run s3_c592 DSRCGPU_SLOTS=3
