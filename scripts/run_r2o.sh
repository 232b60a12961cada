cd $GRAFT_REPO_ROOT
timeout 300 python scripts/ab_pair_small.py > gpurun_out/r2o_pair_small.txt 2> gpurun_out/r2o_pair_small.err
cat gpurun_out/r2o_pair_small.txt; tail -3 gpurun_out/r2o_pair_small.err
