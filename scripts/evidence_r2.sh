# round-2 evidence: launch lists (time + DRAM bytes per launch) of one Stage-II step and one dense-regime step
set -x
cd $GRAFT_REPO_ROOT
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_step.csv python scripts/profile_step.py 1 128 native > gpurun_out/r2_ncu_step.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_launches_step.csv gpurun_out/r2_gemm_traffic.json > gpurun_out/r2_launches_step_ncu.txt 2>&1
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/r2_launches_dense.csv python scripts/profile_step.py 1 16 synthetic 8192 512 > gpurun_out/r2_ncu_dense.log 2>&1
python scripts/summarize_launches.py gpurun_out/r2_launches_dense.csv > gpurun_out/r2_launches_dense_ncu.txt 2>&1
rm -f gpurun_out/r2_launches_step.csv gpurun_out/r2_launches_dense.csv
head -45 gpurun_out/r2_launches_step_ncu.txt
