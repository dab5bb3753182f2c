python tools/dev/mem_probe.py 2>&1 | tail -5
python bench.py --preset cfg2 --no-cpu-baseline --no-exhaustive 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg2', {k:d.get(k) for k in ('ms_per_step','fit_s','acq_s','fit_optimizer_host_s')}, d['e2e']['ms_per_step'])"
python bench.py --preset cfg5 --no-cpu-baseline --no-exhaustive 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg5', {k:d.get(k) for k in ('ms_per_step','fit_s','acq_s','fit_optimizer_host_s')}, d['e2e']['ms_per_step'])"
