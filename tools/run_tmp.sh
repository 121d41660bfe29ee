python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 --no-cpu --no-extras > gpurun_out/bench_r2_n2.json 2> gpurun_out/bench_r2_n2.err; tail -c 300 gpurun_out/bench_r2_n2.err; python -c "
import json
d=json.loads([l for l in open('gpurun_out/bench_r2_n2.json') if l.startswith('{')][-1])
print(d['n_gpus'], d['value'], d['e2e']['value'], d['decode']['value'], d['decode']['e2e']['value'], d.get('archive',{}).get('archives_identical'))
"
