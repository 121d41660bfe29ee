python -m pytest tests/test_gpu_parity.py tests/test_gpu_archive.py tests/test_gpu_shim.py -m gpu -x -q 2>&1 | tail -1
python bench.py --steps 2 --warmup 2 --no-cpu --no-decode --no-serial --no-extras > gpurun_out/sw_h.json 2> gpurun_out/sw_h.err; python -c "import json;d=json.load(open('gpurun_out/sw_h.json'));print(round(d['value']),round(d['e2e']['value']),d['e2e']['frac_of_copy_ceiling'])"
