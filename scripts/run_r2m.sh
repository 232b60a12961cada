set -x
cd $GRAFT_REPO_ROOT
timeout 600 python -m pytest tests -m gpu -x -q > gpurun_out/r2m_pytest.log 2>&1; tail -3 gpurun_out/r2m_pytest.log
OLD=$GRAFT_REPO_ROOT/act_b200/csrc/build/old/libact_b200_old.so
timeout 120 python scripts/ab_edge.py > gpurun_out/r2m_edge_new.json 2>gpurun_out/r2m_edge_new.err
ACT_B200_LIB=$OLD timeout 120 python scripts/ab_edge.py > gpurun_out/r2m_edge_old.json 2>gpurun_out/r2m_edge_old.err
cat gpurun_out/r2m_edge_new.json gpurun_out/r2m_edge_old.json
Q="--config stage2 --steps 60 --sustain-seconds 0 --no-cpu-baseline"
for i in 1 2; do
ACT_BENCH_QUICK=1 timeout 300 python bench.py $Q > gpurun_out/r2m_q_new$i.json 2> gpurun_out/r2m_q_new$i.err
ACT_B200_LIB=$OLD ACT_BENCH_QUICK=1 timeout 300 python bench.py $Q > gpurun_out/r2m_q_old$i.json 2> gpurun_out/r2m_q_old$i.err
done
( time timeout 600 python bench.py > gpurun_out/r2m_bench.json 2> gpurun_out/r2m_bench.err ) 2>&1 | tail -4
python - <<'PY'
import json
for f in ("r2m_q_new1","r2m_q_old1","r2m_q_new2","r2m_q_old2","r2m_bench"):
    try:
        d=json.loads(open(f"gpurun_out/{f}.json").read().strip().splitlines()[-1]); print(f, d["ms_per_step"], d["value"], d.get("e2e",{}).get("value"))
    except Exception as e: print(f, "ERR", e)
PY
