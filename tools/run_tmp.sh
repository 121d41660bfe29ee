python -m pytest tests/test_gpu_parity.py tests/test_fuzz.py -m gpu -x -q 2>&1 | tail -1
bash tools/ab.sh 0 stock 2>&1 | grep -E "==|call|model_dna"
python bench.py --steps 2 --warmup 2 --no-cpu --no-e2e --no-decode --no-serial --no-extras > gpurun_out/sw_d64.json 2> gpurun_out/sw_d64.err; python -c "import json;d=json.load(open('gpurun_out/sw_d64.json'));print(round(d['value']),{k:round(v) for k,v in d['roofline']['kernel_ms_per_step'].items()})"
