
python bench.py --steps 100 --warmup 10 > gpurun_out/bench_r1u.json 2> gpurun_out/bench_r1u.err; echo "bench rc=$?"; tail -3 gpurun_out/bench_r1u.err
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1u.json'))
print(d['ms_per_step'], d['roofline']['frac'])
for k,v in d['extras'].items(): print(k, round(v['ms_per_step'],4), round(v['algorithmic_GBps']), v['kernel'], round(v['agent_steps_per_sec']/1e9,2))
PY
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_r1u.json'))
print(d['value']/1e9, d['gpu_launches'], d['roofline'], d['clocks'], d['config']['launch'])
PY
python bench.py --steps 7 --warmup 3 --no-extras --e2e-steps 1 --steps-per-launch 4 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['gpu_launches'], d['ms_per_step'], d['episode_stats']['env_steps'])"
