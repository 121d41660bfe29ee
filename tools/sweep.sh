# developer aid: scheduler parameter sweep on the bench workload (run under gpurun): bash tools/sweep.sh
run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --no-decode --no-serial $EXTRA > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err; python -c "import json;d=json.load(open('gpurun_out/sw_$name.json'));print('$name',round(d['value']),{k:round(v) for k,v in d['roofline']['kernel_ms_per_step'].items()})"; }
run s1 DSRCGPU_SLOTS=1
run s3 DSRCGPU_SLOTS=3
run s4 DSRCGPU_SLOTS=4
EXTRA="--inflight 8192" run s3_i8192 DSRCGPU_SLOTS=3
