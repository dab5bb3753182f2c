mkdir -p gpurun_out
: > gpurun_out/r02g_share_probe.jsonl
python tools/dev/share_probe.py 1 >> gpurun_out/r02g_share_probe.jsonl 2> gpurun_out/r02g_probe.err
CGP_TILE_MODE=big python tools/dev/share_probe.py 1 >> gpurun_out/r02g_share_probe.jsonl 2>> gpurun_out/r02g_probe.err
CGP_LANES=1 python tools/dev/share_probe.py 1 >> gpurun_out/r02g_share_probe.jsonl 2>> gpurun_out/r02g_probe.err
CGP_LANES=3 python tools/dev/share_probe.py 1 >> gpurun_out/r02g_share_probe.jsonl 2>> gpurun_out/r02g_probe.err
CGP_LANES=4 python tools/dev/share_probe.py 1 >> gpurun_out/r02g_share_probe.jsonl 2>> gpurun_out/r02g_probe.err
CGP_FORK_MAX=0 python tools/dev/share_probe.py 1 >> gpurun_out/r02g_share_probe.jsonl 2>> gpurun_out/r02g_probe.err
CGP_FORK_MAX=4 python tools/dev/share_probe.py 1 >> gpurun_out/r02g_share_probe.jsonl 2>> gpurun_out/r02g_probe.err
CGP_LANES=3 CGP_TILE_MODE=big python tools/dev/share_probe.py 1 >> gpurun_out/r02g_share_probe.jsonl 2>> gpurun_out/r02g_probe.err
CGP_LANES=1 CGP_SHARD_MODE=static python tools/dev/share_probe.py 1 >> gpurun_out/r02g_share_probe.jsonl 2>> gpurun_out/r02g_probe.err
cat gpurun_out/r02g_share_probe.jsonl; tail -3 gpurun_out/r02g_probe.err
