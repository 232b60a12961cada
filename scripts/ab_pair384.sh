cd $GRAFT_REPO_ROOT
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_gemm.py -x -q -m gpu 2>&1 | tail -5
for v in 2 1; do
  echo "ACT_B200_PAIR=$v"
  ACT_B200_PAIR=$v timeout 200 python scripts/kbench_vit.py gpurun_out/kbench_vit_pair$v.json 2>&1 | grep -E "proj|fc2|fc1|qkv_tok|kv_prm"
done
export ACT_BENCH_QUICK=1
for v in 2 1 2 1; do
  ACT_B200_PAIR=$v timeout 200 python bench.py --config stage2 --no-cpu-baseline --sustain-seconds 0 --steps 30 --warmup 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pair=$v', 'ms', d['ms_per_step'], 'student', d['student_only']['ms_per_step'])"
done
