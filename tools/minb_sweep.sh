for B in 3 4 5; do echo "== MC_SEED_MINB=$B"; MC_SEED_MINB=$B python bench.py --pairs 2000000 --steps 3 --warmup 2 --resident-only --no-cpu 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read().replace('Infinity','1e30'))
print(d['value'], d['ms_per_step'], {k:round(v,3) for k,v in d['stages_ms_per_step'].items()}, d['roofline']['frac'])"; done
