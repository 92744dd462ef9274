for lib in lib_nopolicy; do
  for obs in float32 none; do
    CCB200_LIB=$PWD/scripts_tmp/$lib.so python bench.py --steps 100 --warmup 10 --no-extras --e2e-steps 1 --cpu-seconds 0.2 --obs-dtype $obs 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('$lib $obs', round(d['ms_per_step'],4), round(d['roofline']['frac'],3))"
  done
done
