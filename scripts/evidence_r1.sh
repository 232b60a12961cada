set -x
cd $GRAFT_REPO_ROOT
ACT_BENCH_GEMM_TABLE=gpurun_out/gemm_table_r1f.json timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/bench_r1f.json 2> gpurun_out/bench_r1f.err
timeout 200 python scripts/bench_dvae.py 64 20 > gpurun_out/dvae_bench.json 2> gpurun_out/dvae_bench.err
timeout 300 ncu --profile-from-start off --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none --csv --log-file gpurun_out/launches_dvae.csv python scripts/profile_dvae.py 64 > gpurun_out/ncu_dvae.log 2>&1
python scripts/summarize_launches.py gpurun_out/launches_dvae.csv > gpurun_out/launches_dvae_summary.txt 2>&1
timeout 300 ncu --profile-from-start off --set full --clock-control none --import-source on -k regex:"dgcnn_edge_train|gn_rows_train" -c 14 -o gpurun_out/dgcnn_train python scripts/profile_dvae.py 64 > gpurun_out/ncu_dgcnn.log 2>&1
python scripts/ncu_summary.py gpurun_out/dgcnn_train.ncu-rep > gpurun_out/dgcnn_train_summary.txt 2>&1
rm -f gpurun_out/launches_dvae.csv
ls -la gpurun_out | tail -12
