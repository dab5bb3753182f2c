mkdir -p gpurun_out
runN() { n=$1; shift; python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $n "$@"; }
runN 8 --steps 3 --warmup 2 > gpurun_out/r02e_bench_cfg3_n8_async.json 2> gpurun_out/r02e_bench_cfg3_n8_async.err; echo "n8 async rc=$?"; tail -2 gpurun_out/r02e_bench_cfg3_n8_async.err
CGP_SHARD_MODE=balanced runN 8 --steps 2 --warmup 1 --no-exhaustive > gpurun_out/r02e_bench_cfg3_n8_balanced.json 2> gpurun_out/r02e_bench_cfg3_n8_balanced.err; echo "n8 bal rc=$?"
CGP_LANES=1 runN 8 --steps 2 --warmup 1 --no-exhaustive > gpurun_out/r02e_bench_cfg3_n8_lanes1.json 2> gpurun_out/r02e_bench_cfg3_n8_lanes1.err; echo "n8 lanes1 rc=$?"
for f in gpurun_out/r02e_bench_*.json; do echo $f; python -c "
import json,sys
d=json.loads(open('$f').read() or '{}')
print({k:d.get(k) for k in ('value','ms_per_step','fit_s','acq_s','lml_evals_per_fit_rank0','lockstep_per_fit','shard_mode','fit_async_rank0')})
print(d.get('acq_exhaustive_scan'), (d.get('e2e') or {}).get('ms_per_step'), d.get('selected'))
"; done
