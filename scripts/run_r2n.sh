set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests/test_gpu_gemm.py tests/test_gpu_layers.py tests/test_gpu_model.py -m gpu -x -q > gpurun_out/r2n_pytest.log 2>&1; tail -3 gpurun_out/r2n_pytest.log
ACT_B200_EW8=1 timeout 200 python scripts/ab_ew8.py > gpurun_out/r2n_ew8_1.json 2>gpurun_out/r2n_ew8_1.err
ACT_B200_EW8=0 timeout 200 python scripts/ab_ew8.py > gpurun_out/r2n_ew8_0.json 2>gpurun_out/r2n_ew8_0.err
cat gpurun_out/r2n_ew8_1.json gpurun_out/r2n_ew8_0.json
Q="--config stage2 --steps 60 --sustain-seconds 0 --no-cpu-baseline"
for i in 1 2; do
ACT_B200_EW8=1 ACT_BENCH_QUICK=1 timeout 300 python bench.py $Q > gpurun_out/r2n_q_new$i.json 2> gpurun_out/r2n_q_new$i.err
ACT_B200_EW8=0 ACT_BENCH_QUICK=1 timeout 300 python bench.py $Q > gpurun_out/r2n_q_old$i.json 2> gpurun_out/r2n_q_old$i.err
done
python - <<'PY'
import json
for f in ("r2n_q_new1","r2n_q_old1","r2n_q_new2","r2n_q_old2"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d.get("e2e",{}).get("value"), d.get("student_only",{}).get("ms_per_step"))
    except Exception as e: print(f, "ERR", e)
PY
