mkdir -p gpurun_out
python tools/ncu_target_r02.py > gpurun_out/r02m_target_plain.log 2>&1 && \
ncu --set full --clock-control none --import-source on -k regex:"k_knn_tc" -c 1 -o gpurun_out/r02m_full_knn_tc python tools/ncu_target_r02.py > gpurun_out/r02m_target_ncu.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/r02m_target_ncu.log
