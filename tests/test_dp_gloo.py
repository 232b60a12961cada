"""CPU, world_size 2, gloo: the N>1 host logic -- flat parameter/gradient buffers, the single gradient all-reduce
and its 1/world scale, parameter broadcast, disjoint per-rank batches."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from act_b200 import dp, layers
    torch.manual_seed(100 + rank)                      # deliberately different initial weights per rank
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.LayerNorm(32), torch.nn.Linear(32, 8))
    fp = layers.FlatParams(model)
    dp.broadcast_params(fp)
    w0 = fp.flat.clone()
    assert fp.shadow.dtype == torch.bfloat16 and torch.equal(fp.shadow.float(), fp.flat.bfloat16().float())
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(8, 16, generator=g), torch.randn(8, 8, generator=g)
    xs, ys = x_all[rank * 4:(rank + 1) * 4], y_all[rank * 4:(rank + 1) * 4]      # shard by sample
    fp.zero_grad()
    ((model(xs) - ys) ** 2).sum().backward()
    for p in fp.params:                                 # autograd accumulated INTO the flat views
        assert p.grad.data_ptr() == p._act_grad.data_ptr()
    scale = dp.sync_gradients(fp)
    g32 = fp.grad.clone() * scale
    # bf16 communication (the speed mode's default for N>1): the same step again, all-reduced as a bf16 copy
    fp.zero_grad()
    ((model(xs) - ys) ** 2).sum().backward()
    assert fp.enable_bf16_comm() and fp.grad16.dtype == torch.bfloat16
    local = fp.grad.clone()
    dp.sync_gradients(fp)
    assert torch.equal(fp.grad, local)                  # the fp32 buffer keeps the local gradient; AdamW reads grad16
    q.put((rank, w0, g32, dp.shard_seed(5, rank, 0), fp.n_decay, fp.names, fp.grad16.float() * scale))
    dist.barrier()
    dist.destroy_process_group()


def test_flat_gradient_allreduce_world2():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in range(2)], key=lambda t: t[0])
    for p in procs:
        p.join(60)
        assert p.exitcode == 0
    (_, w_a, g_a, s_a, nd, names, g16_a), (_, w_b, g_b, s_b, _, _, g16_b) = res
    assert torch.equal(g16_a, g16_b)                    # every rank applies the same bf16 sum
    assert (g16_a - g_a).abs().max() <= 2 ** -7 * g_a.abs().max()     # bf16 rounding of each addend and of the sum
    assert torch.equal(w_a, w_b)                        # broadcast made the replicas identical
    assert torch.allclose(g_a, g_b) and s_a != s_b
    # reference: full-batch gradient of the mean-over-ranks loss on one process
    torch.manual_seed(100)
    model = torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.LayerNorm(32), torch.nn.Linear(32, 8))
    from act_b200 import layers
    fp = layers.FlatParams(model)
    g = torch.Generator().manual_seed(7)
    x_all, y_all = torch.randn(8, 16, generator=g), torch.randn(8, 8, generator=g)
    fp.zero_grad()
    (((model(x_all) - y_all) ** 2).sum() / 2).backward()
    assert torch.allclose(fp.grad, g_a, rtol=1e-5, atol=1e-6)
    # decay group first (2-D weights), then biases / norm parameters (tools/builder.py:38-51)
    n_dec = sum(1 for n in names if n.endswith("weight") and "1." not in n)
    assert all(n in ("0.weight", "2.weight") for n in names[:n_dec]) and nd == 16 * 32 + 32 * 8
