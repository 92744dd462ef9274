set -x
python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1f.log 2>&1; echo "pytest rc=$?" 
python bench.py > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err; echo "bench rc=$?"
python profiles/write_bw_probe.py > gpurun_out/write_bw.log 2>&1
python bench.py --steps 5 --warmup 3 --no-extras --e2e-steps 1 > gpurun_out/plain_f.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1f.csv python bench.py --steps 5 --warmup 3 --no-extras --e2e-steps 1 > gpurun_out/ncu_f1.log 2>&1
python bench.py --steps 3 --warmup 3 --no-extras --e2e-steps 1 > gpurun_out/plain_f2.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cc_kernel -s 4 -c 1 -o gpurun_out/prof_r1f_fp32 -f python bench.py --steps 3 --warmup 3 --no-extras --e2e-steps 1 > gpurun_out/ncu_f2.log 2>&1
tail -3 gpurun_out/pytest_r1f.log; cat gpurun_out/write_bw.log
