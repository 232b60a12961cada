"""Data-parallel plumbing of the step: the batch shards by cloud (every cloud is independent through Group,
Encoder, Blocks, decoder and loss; BatchNorm statistics stay per-rank as in the reference unless --sync_bn),
so the ONLY collective per step is the all-reduce of the flat gradient buffer -- its bf16 copy in the speed mode
(half the bytes), the fp32 buffer itself in the parity mode or with ACT_B200_GRAD_COMM=fp32
(reference: DistributedDataParallel's bucketed fp32 all-reduce, /root/reference/tools/runner_pretrain.py:84-90).
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
    g16 = getattr(fp, "grad16", None)
    if g16 is not None:
        # bf16 communication (FlatParams.enable_bf16_comm): one cast pass, all-reduce of half the bytes; AdamW reads g16.
        # Done for any world size once enabled, because the captured AdamW reads this buffer.
        if fp.grad.is_cuda:
            from . import ops
            ops.cast_flat_(fp.grad, g16)
        else:
            g16.copy_(fp.grad)
        if w > 1:
            dist.all_reduce(g16, op=dist.ReduceOp.SUM)
    elif w > 1:
        dist.all_reduce(fp.grad, op=dist.ReduceOp.SUM)
    return 1.0 / w


def shard_seed(base, rank, step):
    """Seed of the synthetic batch a rank draws at a step: disjoint across ranks and steps."""
    return base + 100003 * rank + step
