mkdir -p gpurun_out
python -m pytest tests -m gpu -q > gpurun_out/r02_final_pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r02_final_pytest_gpu.log; tail -4 gpurun_out/r02_final_pytest_gpu.log
python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r02_final_smoke.log 2>&1; echo "smoke rc=$?"; tail -1 gpurun_out/r02_final_smoke.log
python bench.py > gpurun_out/r02_final_bench_cfg3.json 2> gpurun_out/r02_final_bench_cfg3.err; echo "cfg3 rc=$?"
python bench.py --impl reference > gpurun_out/r02_final_reference_arm.json 2> gpurun_out/r02_final_reference_arm.err; echo "ref rc=$?"
python bench.py --preset cfg5 > gpurun_out/r02_final_bench_cfg5.json 2> gpurun_out/r02_final_bench_cfg5.err; echo "cfg5 rc=$?"
python bench.py --preset cfg2 > gpurun_out/r02_final_bench_cfg2.json 2> gpurun_out/r02_final_bench_cfg2.err; echo "cfg2 rc=$?"
python bench.py --preset cfg4 --steps 1 --warmup 1 > gpurun_out/r02_final_bench_cfg4.json 2> gpurun_out/r02_final_bench_cfg4.err; echo "cfg4 rc=$?"; tail -2 gpurun_out/r02_final_bench_cfg4.err
python bench.py --impl reference --preset cfg2 --measured --steps 1 --warmup 0 > gpurun_out/r02_final_reference_cfg2_measured.json 2> gpurun_out/r02_final_reference_cfg2_measured.err; echo "ref cfg2 measured rc=$?"
python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-exhaustive > gpurun_out/r02_final_plain_for_ncu.json 2> gpurun_out/r02_final_plain_for_ncu.err && \
ncu --metrics gpu__time_duration.sum --clock-control none -c 40000 --csv --log-file gpurun_out/r02_launches_bench.csv python bench.py --steps 1 --warmup 0 --no-cpu-baseline --no-exhaustive > gpurun_out/r02_final_ncu.log 2>&1
echo "ncu rc=$?"; gzip -f gpurun_out/r02_launches_bench.csv; ls -la gpurun_out/r02_launches_bench.csv.gz
for f in gpurun_out/r02_final_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read() or '{}')
print({k:d.get(k) for k in ('impl','value','ms_per_step','fit_s','acq_s','parity','lml_evals_per_fit_rank0','lml_evals_committed')})
r=d.get('roofline') or {}
print({k:r.get(k) for k in ('kernel','achieved','peak','frac')}, r.get('step'))
print(d.get('acq_exhaustive_scan'), d.get('e2e'), d.get('clocks')); c=d.get('cpu_baseline') or {}; print({k:c.get(k) for k in ('value','cores','kind','extrapolated','s_per_lml_grad_eval','iteration_s_extrapolated','iteration_s','fit_s','selected_idx')})
"; done
