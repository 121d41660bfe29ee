python -m pytest tests -m gpu -q 2>&1 | tail -3 > gpurun_out/r02_pytest_gpu_final.log; cat gpurun_out/r02_pytest_gpu_final.log
python bench.py > gpurun_out/bench_r2_final.json 2> gpurun_out/bench_r2_final.err; tail -c 200 gpurun_out/bench_r2_final.err
ncu --metrics gpu__time_duration.sum --clock-control none -c 8000 --csv --log-file gpurun_out/r02_launches.csv python bench.py --steps 2 --warmup 1 --no-cpu --no-extras --no-decode --no-e2e --no-serial > gpurun_out/r02_bench_under_ncu.json 2> gpurun_out/r02_bench_under_ncu.err
DSRCGPU_SLOTS=1 timeout 900 ncu --set full --import-source on --clock-control none -k regex:"k_model_walk|k_dna_walk|k_rc_encode|k_preprocess_flat|k_tags" --launch-skip 20 -c 5 -f -o gpurun_out/r02_full python tools/phase_prof.py 6000000 0 8192 2>&1 | tail -1
