mkdir -p gpurun_out
python -m pytest tests/test_gpu_prune.py tests/test_gpu_fullsize.py tests/test_gpu_parity.py -m gpu -q -k "prune or cfg or select or hygiene or mspe" > gpurun_out/r02j_pytest.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02j_pytest.log
: > gpurun_out/r02j_scan_probe.jsonl
python tools/dev/scan_probe.py 32768 1048576 >> gpurun_out/r02j_scan_probe.jsonl 2> gpurun_out/r02j_probe.err
CGP_SCORE_CHUNK_MIN=0 python tools/dev/scan_probe.py 32768 1048576 >> gpurun_out/r02j_scan_probe.jsonl 2>> gpurun_out/r02j_probe.err
cat gpurun_out/r02j_scan_probe.jsonl; tail -3 gpurun_out/r02j_probe.err
python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-exhaustive > gpurun_out/r02j_bench_cfg3.json 2> gpurun_out/r02j_bench_cfg3.err
python -c "
import json
d=json.loads(open('gpurun_out/r02j_bench_cfg3.json').read())
print({k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','selected')})"
