python bench.py > gpurun_out/bench_r1s.json 2> gpurun_out/bench_r1s.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r1s.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1s.json'))
print(d['ms_per_step'], d['roofline'], d['e2e']['value'], d['clocks'])
for k,v in d['extras'].items(): print(k, round(v['ms_per_step'],4), round(v['algorithmic_GBps']), v['kernel'], round(v['agent_steps_per_sec']/1e9,2))
PY
python bench.py --steps 5 --warmup 3 --no-extras --e2e-steps 1 > gpurun_out/plain_s.log 2>&1 && ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/launches_r1s.csv python bench.py --steps 5 --warmup 3 --no-extras --e2e-steps 1 > gpurun_out/ncu_s1.log 2>&1
