cd $GRAFT_REPO_ROOT
export ACT_BENCH_QUICK=1
for cfg in "0 0 0" "0 0 120" "1 1 0" "1 1 120" "1 1 100" "1 1 84" "1 0 0"; do
  set -- $cfg
  ACT_B200_PIPELINE=$1 ACT_BENCH_LOOKAHEAD=$2 ACT_B200_TEACHER_SM_CAP=$3 timeout 200 python bench.py --config stage2 --no-cpu-baseline --sustain-seconds 0 --steps 30 --warmup 6 2>/dev/null | python -c "
import json,sys
d=json.loads(sys.stdin.read())
print('pipeline=$1 lookahead=$2 cap=$3', 'ms', d['ms_per_step'], 'student', d['student_only']['ms_per_step'] if d['student_only'] else None)"
done
