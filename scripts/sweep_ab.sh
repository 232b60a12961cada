cd $GRAFT_REPO_ROOT
export ACT_BENCH_QUICK=1
for v in 1 0 1 0; do
  ACT_B200_FUSED_BN_STATS=$v timeout 200 python bench.py --config stage2 --no-cpu-baseline --sustain-seconds 0 --steps 30 --warmup 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('fused_stats=$v', 'ms', d['ms_per_step'], 'student', d['student_only']['ms_per_step'])"
done
