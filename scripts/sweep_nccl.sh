# N=8 A/B of NCCL settings for the overlapped bf16 gradient all-reduce (Stage-II config only)
cd $GRAFT_REPO_ROOT
run() {
  tag=$1; shift
  env "$@" timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29519 bench.py --gpus 8 --config stage2 --no-cpu-baseline --sustain-seconds 0 2> gpurun_out/nccl_$tag.err | python -c "import json,sys; d=json.loads(sys.stdin.read()); print('$tag', d['value'], d['ms_per_step'], d['student_only']['ms_per_step'])"
}
run default NCCL_DEBUG=INFO NCCL_DEBUG_SUBSYS=INIT
grep -E "NVLS|Channel|channels|Using network|algo" gpurun_out/nccl_default.err | grep -v "Channel [0-9]*/[0-9]* :" | sort | uniq -c | sort -rn | head -12
run ctas4 NCCL_MAX_CTAS=4
run ctas8 NCCL_MAX_CTAS=8
run ctas16 NCCL_MAX_CTAS=16
run ring NCCL_ALGO=Ring
run nvls NCCL_ALGO=NVLS
run fp32 ACT_B200_GRAD_COMM=fp32
