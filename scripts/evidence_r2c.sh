# round-2 closing evidence on ONE B200: bench line + GEMM table, ncu launch lists (time + DRAM bytes per launch) of one
# Stage-II / dense / Stage-I step, ncu --set full of the student fc1 GEMM on 8 vs 4 epilogue warps
set -x
cd $GRAFT_REPO_ROOT
ACT_BENCH_GEMM_TABLE=gpurun_out/r2_gemm_table_c.json timeout 600 python bench.py > gpurun_out/r2_bench_1gpu_d.json 2> gpurun_out/r2_bench_1gpu_d.err
M=gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/l_step.csv python scripts/profile_step.py 1 128 native > gpurun_out/ncu_step.log 2>&1
python scripts/summarize_launches.py gpurun_out/l_step.csv gpurun_out/r2_gemm_traffic.json > gpurun_out/r2_launches_step_ncu.txt 2>&1
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/l_dense.csv python scripts/profile_step.py 1 16 synthetic 8192 512 > gpurun_out/ncu_dense.log 2>&1
python scripts/summarize_launches.py gpurun_out/l_dense.csv > gpurun_out/r2_launches_dense_ncu.txt 2>&1
timeout 400 ncu --profile-from-start off --metrics $M --clock-control none --csv --log-file gpurun_out/l_dvae.csv python scripts/profile_dvae.py 64 > gpurun_out/ncu_dvae.log 2>&1
python scripts/summarize_launches.py gpurun_out/l_dvae.csv > gpurun_out/r2_launches_dvae_ncu.txt 2>&1
rm -f gpurun_out/l_step.csv gpurun_out/l_dense.csv gpurun_out/l_dvae.csv
for ew in 1 0; do
ACT_B200_EW8=$ew timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -f -o gpurun_out/r2_fc1_ew$ew python scripts/prof_one_gemm.py enc_fc1 > gpurun_out/ncu_fc1_$ew.log 2>&1
( echo "## enc_fc1 (3456x1536x384 + bias + GELU + pre-activation output), ACT_B200_EW8=$ew"; python scripts/ncu_summary.py gpurun_out/r2_fc1_ew$ew.ncu-rep ) >> gpurun_out/r2_ncu_fc1_ew8_full.txt 2>&1
done
rm -f gpurun_out/*.ncu-rep
head -12 gpurun_out/r2_launches_step_ncu.txt
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_1gpu_d.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['student_only']['ms_per_step'], d['sustained']['ms_per_step'], d['configs']['dvae']['ms_per_step'], d['configs']['dense']['ms_per_step'], d['roofline']['frac'])"
