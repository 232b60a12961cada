"""Data-parallel plumbing of the step: the batch shards by cloud (every cloud is independent through Group,
Encoder, Blocks, decoder and loss; BatchNorm statistics stay per-rank as in the reference unless --sync_bn),
so the ONLY collective per step is the all-reduce of the flat fp32 gradient buffer
(reference: DistributedDataParallel's bucketed all-reduce, /root/reference/tools/runner_pretrain.py:84-90).
One process per GPU, torch.distributed (NCCL over NVLink on the GPU box, gloo in the CPU tests)."""
import torch
import torch.distributed as dist


def world_size():
    return dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1


def broadcast_params(fp, src=0):
    """Make every rank start from rank `src`'s parameters (DDP does this at construction); one broadcast of
    the flat buffer, then refresh the bf16 shadow."""
    if world_size() > 1:
        dist.broadcast(fp.flat, src)
    fp.refresh_shadow()


def sync_gradients(fp):
    """Sum the flat gradient over ranks in ONE all-reduce; returns the scale (1/world) that the fused AdamW
    applies while reading the gradient, so no separate averaging pass touches HBM."""
    w = world_size()
    if w > 1:
        dist.all_reduce(fp.grad, op=dist.ReduceOp.SUM)
    return 1.0 / w


def shard_seed(base, rank, step):
    """Seed of the synthetic batch a rank draws at a step: disjoint across ranks and steps."""
    return base + 100003 * rank + step
