# developer aid (run under gpurun): parity tests, isolated kernel times on a 2.2 GB sample, one short bench line
python -m pytest tests/test_gpu_parity.py tests/test_gpu_archive.py -m gpu -x -q 2>&1 | tail -1
DSRCGPU_SLOTS=1 python tools/phase_prof.py 6000000 0 4096 | grep -E "rc_encode|model_quality|model_dna|preprocess|tags"
python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --no-decode --no-serial > gpurun_out/bench_u.json 2> gpurun_out/bench_u.err; python -c "import json;d=json.load(open('gpurun_out/bench_u.json'));print(d['value'],d['roofline']['kernel_ms_per_step'])"
