python bench.py --preset cfg2 --no-cpu-baseline --no-exhaustive 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('ms_per_step','fit_s','acq_s','fit_optimizer_host_s')}, d['e2e']['ms_per_step'])"
python -m cProfile -o /tmp/b.prof bench.py --preset cfg2 --no-cpu-baseline --no-exhaustive --steps 6 --warmup 3 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print({k:d.get(k) for k in ('ms_per_step','fit_s','acq_s','fit_optimizer_host_s')}, d['e2e']['ms_per_step'])"
python -c "
import pstats; pstats.Stats('/tmp/b.prof').sort_stats('tottime').print_stats(22)" | tail -40
