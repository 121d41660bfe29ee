python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu_final.log; cat gpurun_out/r02_pytest_gpu_final.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -2
python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 200 gpurun_out/bench_r2_final.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_final.json') if l.startswith('{')][-1])
print(d['value'], d['e2e']['value'], d['decode']['value'], d['decode']['e2e']['value'], d['roofline']['kernel'], d['roofline']['frac'])
"
