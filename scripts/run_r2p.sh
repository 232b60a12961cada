cd $GRAFT_REPO_ROOT
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/r2_bench_2gpu_c.json 2> gpurun_out/r2_bench_2gpu_c.err
tail -2 gpurun_out/r2_bench_2gpu_c.err
python -c "
import json; d=json.loads(open('gpurun_out/r2_bench_2gpu_c.json').read().strip().splitlines()[-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('per_rank_ms_per_step'), d['configs']['dvae']['ms_per_step'], d['configs']['dense']['ms_per_step'])"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29518 bench.py --impl reference --gpus 2 --steps 3 --warmup 1 > gpurun_out/r2_bench_2gpu_ref.json 2> gpurun_out/r2_bench_2gpu_ref.err
tail -1 gpurun_out/r2_bench_2gpu_ref.json | cut -c1-400
