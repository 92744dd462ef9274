python -m pytest tests -m gpu -x -q 2>&1 | tail -3
python bench.py --steps 112 --warmup 10 --no-extras --e2e-steps 2 --cpu-seconds 0.5 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('fused16', d['ms_per_step'], d['roofline']['frac'])"
python bench.py --steps 100 --warmup 10 --no-extras --e2e-steps 2 --cpu-seconds 0.5 --steps-per-launch 1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('single', d['ms_per_step'], d['roofline']['frac'])"
