run() { name=$1; shift; env "$@" python bench.py --steps 2 --warmup 2 --no-cpu --no-decode --no-serial --no-extras $EXTRA > gpurun_out/sw_$name.json 2> gpurun_out/sw_$name.err; python -c "import json;d=json.load(open('gpurun_out/sw_$name.json'));print('$name',round(d['value']),round(d['e2e']['value']),{k:round(v) for k,v in d['roofline']['kernel_ms_per_step'].items()})" || tail -3 gpurun_out/sw_$name.err; }
run s3g2
run s4g3 DSRCGPU_SLOTS=4 DSRCGPU_RC_GROUP=3
