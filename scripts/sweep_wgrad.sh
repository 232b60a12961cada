cd $GRAFT_REPO_ROOT
export ACT_BENCH_QUICK=1
for v in 148 111 74 148 50; do
  ACT_B200_WGRAD_CTAS=$v timeout 200 python bench.py --config stage2 --no-cpu-baseline --sustain-seconds 0 --steps 30 --warmup 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('wgrad_ctas=$v', 'ms', d['ms_per_step'], 'student', d['student_only']['ms_per_step'])"
done
