python -m pytest tests -m gpu -x -q > gpurun_out/pytest_r1g.log 2>&1; echo "pytest rc=$?"; tail -15 gpurun_out/pytest_r1g.log
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r1g.json 2> gpurun_out/bench_r1g.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r1g.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1g.json'))
print(d['ms_per_step'], d['roofline']['frac'], d['e2e']['value'])
for k,v in d['extras'].items(): print(k, round(v['ms_per_step'],4), round(v['algorithmic_GBps']))
PY
