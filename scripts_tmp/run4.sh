python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1r.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/pytest_r1r.log
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r1r.json 2> gpurun_out/bench_r1r.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r1r.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1r.json'))
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
for k,v in d['extras'].items(): print(k, round(v['ms_per_step'],4), round(v['algorithmic_GBps']))
PY
python bench.py --steps 3 --warmup 3 --no-extras --e2e-steps 1 > gpurun_out/plain_r.log 2>&1 && ncu --set full --clock-control none --import-source on -k regex:cc_step_tpe -s 4 -c 1 -o gpurun_out/prof_r1r_fp32 -f python bench.py --steps 3 --warmup 3 --no-extras --e2e-steps 1 > gpurun_out/ncu_r1.log 2>&1
