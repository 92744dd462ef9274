python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py > gpurun_out/bench_r1w.json 2> gpurun_out/bench_r1w.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r1w.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1w.json'))
print(d['value']/1e9, d['ms_per_step'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['traffic'], d['e2e']['value']/1e6, d['cpu_baseline']['value']/1e6, d['clocks'])
for k,v in d['extras'].items(): print(k, round(v['ms_per_step'],4), round(v['algorithmic_GBps']), v['kernel'], round(v['agent_steps_per_sec']/1e9,2))
PY
python bench.py --impl reference --steps 3 --warmup 1 | cut -c1-400
