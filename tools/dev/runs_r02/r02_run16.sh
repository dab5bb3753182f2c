python -m pytest tests/test_gpu_parity.py tests/test_gpu_fullsize.py -m gpu -q -k "knn or cfg" 2>&1 | tail -3
python bench.py --preset cfg5 --steps 1 --warmup 1 --no-cpu-baseline --no-exhaustive 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg5', {k:d.get(k) for k in ('ms_per_step','fit_s','acq_s','selected')})"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_detail_cfg5_n1.json'))
for k in ('knn','knn_exact','knn_prep','cross','predict_vnorm'): print(k, d['kernel_classes'][k])
PY
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-exhaustive 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read()); print('cfg3', {k:d.get(k) for k in ('ms_per_step','fit_s','acq_s','selected')})"
python - <<'PY'
import json
d=json.load(open('gpurun_out/bench_detail_cfg3_n1.json'))
for k in ('knn','knn_exact','knn_prep','cross','predict_vnorm'): print(k, d['kernel_classes'][k])
PY
