mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_final2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r02_final2_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | tail -1
python bench.py > gpurun_out/r02_final2_bench_cfg3.json 2> gpurun_out/r02_final2_bench_cfg3.err; echo "cfg3 rc=$?"
python -c "
import json
s=open('gpurun_out/r02_final2_bench_cfg3.json').read(); d=json.loads(s)
print(len(s), {k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','parity','selected')}); r=d['roofline']; print({k:r.get(k) for k in ('kernel','achieved','peak','frac')}, r['step'], r['classes']); print(d['e2e'], d['clocks'], d['acq_exhaustive_scan']); print(d['cpu_baseline']['value'], d['cpu_comparison_fully_measured'])"
