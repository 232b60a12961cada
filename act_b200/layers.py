"""Autograd glue between the reference's nn.Module surface and the act_b200 kernels.

Each autograd.Function here is one FUSED region of the ACT step, forward and hand-written backward, made
only of libact_b200 launches (ops.py) plus O(channels) host-side tensor arithmetic:

  TransformerStack  -- N pre-LN Blocks, `x = block(x + pos)`           (models/act.py:72-112, 140-142)
  LayerNormFn       -- the trailing nn.LayerNorm of encoder / decoder   (act.py:301, 144)
  LinearFn          -- nn.Linear (+GELU) on the tcgen05 GEMM           (pos_embed[2], proj_head, reduce_dim)
  PointNetEncoderFn -- the per-group mini-PointNet `Encoder`            (models/dvae.py:185-215)
  CosineLossFn      -- the distillation loss                            (act.py:1243-1254)

Parameters stay fp32 nn.Parameters (state_dict compatible with the reference).  When a module has been
adopted by `FlatParams`, every parameter is a view into one flat fp32 buffer, carries `_act_shadow` (its
bf16 copy inside one flat bf16 buffer, refreshed by the fused AdamW kernel) and `_act_grad` (its slice of one
flat fp32 gradient buffer): weight-gradient GEMMs then accumulate straight into that slice and the Functions
return None for those inputs.  Without FlatParams the Functions cast weights on the fly and return gradients
the normal autograd way, so the modules also work inside a stock PyTorch training loop / DDP.
"""
import math
import os

import torch

from . import ops


class _SideStream:
    """Weight-gradient work (wgrad GEMMs, bias column sums) is off the critical path of a backward pass: only
    the dgrad chain feeds the next layer.  It is issued on a second CUDA stream, forked after the dY it needs
    and joined once at the end of the Function's backward, so inside the captured graph the wgrads run
    CONCURRENTLY with the latency-bound dgrad / LayerNorm / attention kernels of the following layers."""
    _streams = {}
    enabled = os.environ.get("ACT_B200_SIDE_STREAM", "1") != "0"

    def __init__(self, device, tag="wgrad"):
        self.main = torch.cuda.current_stream(device)
        self.side = None
        if _SideStream.enabled:
            key = (device.index, self.main.cuda_stream, tag)
            if key not in _SideStream._streams:
                _SideStream._streams[key] = torch.cuda.Stream(device=device)
            self.side = _SideStream._streams[key]
        self.used = False

    def run(self, fn, *tensors):
        """fn() launches kernels reading `tensors` (already produced on the main stream)."""
        if self.side is None:
            fn()
            return
        self.side.wait_stream(self.main)
        for t in tensors:
            if t is not None:
                t.record_stream(self.side)
        with torch.cuda.stream(self.side):
            fn()
        self.used = True

    def join(self):
        if self.side is not None and self.used:
            self.main.wait_stream(self.side)


def _flat_attr(p, name):
    """The flat-buffer slice (`_act_shadow` / `_act_grad`, set by FlatParams) of parameter p -- or, when p is a pure reshape
    of a parameter (conv.weight.squeeze(-1): same elements, same order, same address), the base parameter's slice viewed in
    p's shape, so that 1x1-conv weights need no per-step cast / zero-fill / add passes either."""
    v = getattr(p, name, None)
    if v is not None:
        return v
    b = getattr(p, "_base", None)
    if (b is not None and b.numel() == p.numel() and p.is_contiguous() and b.is_contiguous()
            and p.data_ptr() == b.data_ptr()):
        v = getattr(b, name, None)
        if v is not None:
            return v.view(p.shape)
    return None


def shadow(p):
    """The GEMM operand form of a weight: its slice of the flat bf16 shadow (or an on-the-fly bf16 cast); in the
    fp32x3 parity mode the f32 master weight itself (ops.gemm splits it into bf16 pieces)."""
    if ops.act_dtype() == torch.float32:
        return p.detach()
    s = _flat_attr(p, "_act_shadow")
    return s if s is not None else p.detach().to(torch.bfloat16)


class _GradSink:
    """Hands out accumulation targets for parameter gradients: the flat-buffer slice when there is one
    (result: None returned to autograd), else a fresh zero tensor (returned to autograd)."""

    def __init__(self):
        self.ret = {}

    def get(self, p, key):
        g = _flat_attr(p, "_act_grad")
        if g is not None:
            self.ret[key] = None
            return g
        g = torch.zeros_like(p, dtype=torch.float32)
        self.ret[key] = g
        return g

    def result(self, key):
        return self.ret.get(key)


class FlatParams:
    """Re-homes a module's trainable parameters into flat buffers: fp32 master `flat`, fp32 `grad`, bf16
    `shadow`, AdamW moments.  Decay group first, no-decay group after (tools/builder.py:38-51: no decay for
    1-D tensors, `.bias`, and names containing `token`).  Each tensor is padded to a multiple of 8 elements so
    every bf16 shadow starts 16-byte aligned (TMA requirement)."""

    def __init__(self, module, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.05, exclude=()):
        """exclude: name prefixes of parameters that never receive a gradient on this path (the student's unused
        `lm_head` / `cls_head`).  torch.optim.AdamW skips parameters whose .grad is None -- no
        moment update and NO weight decay -- so they are left out of the flat buffers and stay untouched."""
        named = [(n, p) for n, p in module.named_parameters()
                 if p.requires_grad and not any(n.startswith(e) for e in exclude)]
        if not named:
            raise ValueError("no trainable parameters")
        dev = named[0][1].device

        def no_decay(n, p):
            return p.dim() <= 1 or n.endswith(".bias") or "token" in n

        order = [(n, p) for n, p in named if not no_decay(n, p)] + [(n, p) for n, p in named if no_decay(n, p)]
        self.n_decay_params = sum(1 for n, p in named if not no_decay(n, p))
        offs, total = [], 0
        for i, (n, p) in enumerate(order):
            if i == self.n_decay_params:
                self.n_decay = total
            offs.append(total)
            total += (p.numel() + 7) // 8 * 8
        if self.n_decay_params == len(order):
            self.n_decay = total
        self.total = total
        self.flat = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad = torch.zeros(total, dtype=torch.float32, device=dev)
        self.grad16 = None                     # bf16 communication copy (enable_bf16_comm)
        self.shadow = torch.zeros(total, dtype=torch.bfloat16, device=dev)
        self.exp_avg = torch.zeros(total, dtype=torch.float32, device=dev)
        self.exp_avg_sq = torch.zeros(total, dtype=torch.float32, device=dev)
        self.params, self.names = [], []
        for (n, p), o in zip(order, offs):
            k = p.numel()
            self.flat[o:o + k].copy_(p.detach().reshape(-1))
            p.data = self.flat[o:o + k].view(p.shape)
            p.grad = self.grad[o:o + k].view(p.shape)
            p._act_grad = p.grad
            p._act_shadow = self.shadow[o:o + k].view(p.shape)
            self.params.append(p)
            self.names.append(n)
        self.shadow.copy_(self.flat)
        self.betas, self.eps = betas, eps
        self.step_count = 0
        self.hyper = torch.zeros(8, dtype=torch.float32, device=dev)
        self._offs = dict(zip(self.names, offs))
        # torch.optim.AdamW's view of the same parameters, in the reference's order (tools/builder.py:38-55: group 0 =
        # no-decay, group 1 = decay, each in named_parameters() order, every requires_grad parameter listed -- the ones
        # excluded above included: torch keeps them in the groups and simply never creates state for them)
        all_named = [(n, p) for n, p in module.named_parameters() if p.requires_grad]
        g_nd = [n for n, p in all_named if no_decay(n, p)]
        g_d = [n for n, p in all_named if not no_decay(n, p)]
        self._opt_index = {n: i for i, n in enumerate(g_nd + g_d)}
        common = dict(lr=lr, betas=tuple(betas), eps=eps, amsgrad=False, maximize=False, foreach=None, capturable=False,
                      differentiable=False, fused=None)
        self.param_groups = [dict(common, weight_decay=0.0, params=[self._opt_index[n] for n in g_nd]),
                             dict(common, weight_decay=weight_decay, params=[self._opt_index[n] for n in g_d])]

    # lr / weight_decay live in param_groups, like torch.optim: an LR scheduler that writes group['lr'] (timm's
    # CosineLRScheduler, torch.optim.lr_scheduler.*: the reference's build_opti_sche) drives this object unchanged.  The
    # fused kernel takes ONE learning rate: both groups must carry the same one (they do in the reference).
    @property
    def lr(self):
        return self.param_groups[1]["lr"]

    @lr.setter
    def lr(self, v):
        for g in self.param_groups:
            g["lr"] = v

    @property
    def weight_decay(self):
        return self.param_groups[1]["weight_decay"]

    @weight_decay.setter
    def weight_decay(self, v):
        self.param_groups[1]["weight_decay"] = v

    def state_dict(self):
        """torch.optim.AdamW.state_dict() layout ({'state': {index: {'step', 'exp_avg', 'exp_avg_sq'}}, 'param_groups':
        [...]}) with the reference's parameter indexing, so `ckpt['optimizer']` written by tools/builder.py:132-144
        round-trips: a reference run resumes from it and vice versa.  Call the engine's flush() first when training in
        pipelined mode (engine.PretrainStep.save_checkpoint does)."""
        state = {}
        if self.step_count > 0:
            for n, p in zip(self.names, self.params):
                o, k = self._offs[n], p.numel()
                state[self._opt_index[n]] = {"step": torch.tensor(float(self.step_count)),
                                             "exp_avg": self.exp_avg[o:o + k].view(p.shape).clone(),
                                             "exp_avg_sq": self.exp_avg_sq[o:o + k].view(p.shape).clone()}
        return {"state": state, "param_groups": [dict(g, params=list(g["params"])) for g in self.param_groups]}

    def load_state_dict(self, sd):
        """Inverse of state_dict(); also accepts a torch.optim.AdamW state_dict of the reference's optimizer."""
        by_index = {i: n for n, i in self._opt_index.items()}
        steps = []
        self.exp_avg.zero_()
        self.exp_avg_sq.zero_()
        for idx, st in sd["state"].items():
            n = by_index.get(int(idx))
            if n is None or n not in self._offs:
                continue                      # state of a parameter that is excluded here (never had a gradient)
            o = self._offs[n]
            k = st["exp_avg"].numel()
            self.exp_avg[o:o + k].copy_(st["exp_avg"].reshape(-1))
            self.exp_avg_sq[o:o + k].copy_(st["exp_avg_sq"].reshape(-1))
            steps.append(int(float(st["step"])))
        if len(set(steps)) > 1:
            raise ValueError("FlatParams keeps one step count; the loaded per-parameter steps differ")
        self.step_count = steps[0] if steps else 0
        for g, src in zip(self.param_groups, sd["param_groups"]):
            for k in ("lr", "betas", "eps", "weight_decay"):
                if k in src:
                    g[k] = src[k]
        self.betas, self.eps = tuple(self.param_groups[1]["betas"]), self.param_groups[1]["eps"]
        self.refresh_shadow()

    def refresh_shadow(self):
        """After loading a state_dict (which writes through the fp32 views)."""
        self.shadow.copy_(self.flat)

    def zero_grad(self):
        if self.grad.is_cuda:
            ops.zero_(self.grad)          # a memset node in the captured step
        else:
            self.grad.zero_()

    def set_hyper(self, grad_scale=1.0):
        """Stage this step's AdamW scalars (lr from the scheduler, bias corrections) into device memory."""
        self.step_count += 1
        t = self.step_count
        if self.param_groups[0]["lr"] != self.param_groups[1]["lr"]:
            raise ValueError("FlatParams: the fused AdamW kernel applies one learning rate to both parameter groups")
        vals = [self.lr, self.betas[0], self.betas[1], self.eps, self.weight_decay, 1.0 - self.betas[0] ** t,
                1.0 - self.betas[1] ** t, grad_scale]
        host = torch.tensor(vals, dtype=torch.float32)
        if self.hyper.is_cuda:       # fresh pinned staging per step: the host may run many steps ahead of the GPU
            host = host.pin_memory()
        self.hyper.copy_(host, non_blocking=True)

    def enable_bf16_comm(self):
        """N>1: all-reduce a bf16 copy of the flat gradient (half the NVLink bytes and half NCCL's HBM traffic beside the
        teacher's GEMMs) and apply AdamW from it.  Must be called before the step is captured; ACT_B200_GRAD_COMM=fp32
        keeps DDP's fp32 all-reduce (the parity mode always does)."""
        import os
        if os.environ.get("ACT_B200_GRAD_COMM", "bf16") == "fp32" or ops.act_dtype() == torch.float32:
            return False
        if self.grad16 is None:
            self.grad16 = torch.zeros(self.grad.numel(), dtype=torch.bfloat16, device=self.grad.device)
        return True

    def step(self):
        """One fused AdamW launch over the flat buffers (hyper-parameters read from device memory)."""
        g = self.grad16 if self.grad16 is not None else self.grad
        ops.adamw(self.flat, g, self.exp_avg, self.exp_avg_sq, self.shadow, self.n_decay, self.hyper)


# ------------------------------------------------------------------------------------- Transformer stack
BLOCK_KEYS = ("norm1.weight", "norm1.bias", "attn.qkv.weight", "attn.proj.weight", "attn.proj.bias", "norm2.weight",
              "norm2.bias", "mlp.fc1.weight", "mlp.fc1.bias", "mlp.fc2.weight", "mlp.fc2.bias")
NP = len(BLOCK_KEYS)


class TransformerStack(torch.autograd.Function):
    """y = blocks[-1](... blocks[0](x + pos) ... + pos): `depth` pre-LN Blocks with `pos` re-added before every
    Block (models/act.py:109-112).  gates: None or f32 [2*depth, B] DropPath gates (0 or 1/keep) for the
    attention / MLP branch of each Block (timm DropPath, act.py:88-89)."""

    @staticmethod
    def forward(ctx, x, pos, gates, num_heads, eps, *params):
        B, T, C = x.shape
        M = B * T
        depth = len(params) // NP
        scale = (C // num_heads) ** -0.5
        need = any(ctx.needs_input_grad)
        cur = x.reshape(M, C).contiguous().float()
        pos2 = pos.reshape(M, C).contiguous().float() if pos is not None else None
        saved = []
        for l in range(depth):
            n1w, n1b, wqkv, wproj, bproj, n2w, n2b, w1, b1, w2, b2 = params[l * NP:(l + 1) * NP]
            g1 = gates[2 * l] if gates is not None else None
            g2 = gates[2 * l + 1] if gates is not None else None
            h1, xs, mean1, rstd1 = ops.layernorm_fwd(cur, n1w, n1b, eps, pos=pos2)
            if xs is None:
                xs = cur
            qkv = ops.gemm(h1, shadow(wqkv))
            o, lse = ops.attention_fwd(qkv, B, T, num_heads, scale)
            xmid = ops.gemm(o, shadow(wproj), bias=bproj, resid=xs, row_scale=g1, rows_per_scale=T,
                            out_dtype=torch.float32)
            h2, _, mean2, rstd2 = ops.layernorm_fwd(xmid, n2w, n2b, eps)
            u = torch.empty(M, w1.shape[0], dtype=ops.act_dtype(), device=x.device) if need else None
            a = ops.gemm(h2, shadow(w1), bias=b1, act=ops.ACT_GELU, preact_out=u)
            cur = ops.gemm(a, shadow(w2), bias=b2, resid=xmid, row_scale=g2, rows_per_scale=T,
                           out_dtype=torch.float32)
            if need:
                saved += [xs, h1, mean1, rstd1, qkv, o, lse, xmid, h2, mean2, rstd2, u, a]
        if need:
            ctx.save_for_backward(*params, *saved)
            ctx.gates = gates
            ctx.meta = (B, T, C, depth, num_heads, scale, pos is not None)
        return cur.view(B, T, C)

    @staticmethod
    def backward(ctx, dy):
        B, T, C, depth, H, scale, has_pos = ctx.meta
        M = B * T
        tensors = ctx.saved_tensors
        params, saved = tensors[:depth * NP], tensors[depth * NP:]
        gates = ctx.gates
        sink = _GradSink()
        side = _SideStream(dy.device)
        dx = dy.reshape(M, C).contiguous().float()
        dpos = ops.zero_(torch.empty(M, C, dtype=torch.float32, device=dy.device)) if has_pos else None
        g = None
        NS = 13
        for l in reversed(range(depth)):
            n1w, n1b, wqkv, wproj, bproj, n2w, n2b, w1, b1, w2, b2 = params[l * NP:(l + 1) * NP]
            xs, h1, mean1, rstd1, qkv, o, lse, xmid, h2, mean2, rstd2, u, a = saved[l * NS:(l + 1) * NS]
            gate1 = gates[2 * l] if gates is not None else None
            gate2 = gates[2 * l + 1] if gates is not None else None
            G = lambda p, k: sink.get(p, (l, k))  # noqa: E731
            if g is None:       # top of the stack: the incoming gradient is plain fp32
                g = ops.cast_rows(dx, gate2, T, dbias=G(b2, 10))
            # MLP branch (weight / bias gradients go to the side stream)
            gw2, gw1, gb1, gwp, gwq = G(w2, 9), G(w1, 7), G(b1, 8), G(wproj, 3), G(wqkv, 2)
            side.run(lambda: ops.wgrad(g, a, gw2), g)
            du = ops.gemm(g, shadow(w2), b_mn=True, mul_in=u, mul_mode=ops.MUL_GELU_GRAD)
            side.run(lambda: (ops.wgrad(du, h2, gw1), ops.colsum(du, gb1)), du)
            dh2 = ops.gemm(du, shadow(w1), b_mn=True)
            dxm, g2 = ops.layernorm_bwd(dh2, xmid, mean2, rstd2, n2w, G(n2w, 5), G(n2b, 6), dres=dx, want_bf16=True,
                                        row_scale=gate1, rows_per_scale=T, dbias=G(bproj, 4))
            # attention branch
            side.run(lambda: ops.wgrad(g2, o, gwp), g2)
            do = ops.gemm(g2, shadow(wproj), b_mn=True)
            dqkv = ops.attention_bwd(qkv, o, do, lse, B, T, H, scale)
            side.run(lambda: ops.wgrad(dqkv, h1, gwq), dqkv)
            dh1 = ops.gemm(dqkv, shadow(wqkv), b_mn=True)
            if l > 0:           # this LayerNorm backward also emits the gated bf16 dY (+ bias grad) of Block l-1's fc2
                pb2 = params[(l - 1) * NP + 10]
                pg2 = gates[2 * (l - 1) + 1] if gates is not None else None
                dx, g = ops.layernorm_bwd(dh1, xs, mean1, rstd1, n1w, G(n1w, 0), G(n1b, 1), dres=dxm, dacc=dpos,
                                          want_bf16=True, row_scale=pg2, rows_per_scale=T,
                                          dbias=sink.get(pb2, (l - 1, 10)))
            else:
                dx, _ = ops.layernorm_bwd(dh1, xs, mean1, rstd1, n1w, G(n1w, 0), G(n1b, 1), dres=dxm, dacc=dpos)
        side.join()
        pgrads = [sink.result((l, k)) for l in range(depth) for k in range(NP)]
        return (dx.view(B, T, C), dpos.view(B, T, C) if has_pos else None, None, None, None, *pgrads)


def transformer_stack(x, pos, blocks, num_heads, eps=1e-5, gates=None):
    params = []
    for blk in blocks:
        sd = dict(blk.named_parameters())
        params += [sd[k] for k in BLOCK_KEYS]
    return TransformerStack.apply(x, pos, gates, num_heads, eps, *params)


_KEEP_CACHE = {}
_DP = {"seed": None, "calls": 0}


class drop_path_seed:
    """Context manager: inside it, DropPath gates are drawn by the act_b200 kernel from `seed` (a device int64 [1] the
    caller refreshes before every step -- engine.PretrainStep stages it from the host next to the mask and the AdamW
    scalars), so a captured step contains no library RNG kernel.  Outside it the seed comes from torch.randint on the
    device (one library kernel; graph-safe too)."""

    def __init__(self, seed):
        self.seed = seed

    def __enter__(self):
        self.prev = _DP["seed"]
        _DP["seed"] = self.seed

    def __exit__(self, *a):
        _DP["seed"] = self.prev


def drop_path_gates(rates, batch, device, training):
    """timm 0.5.4 DropPath: per-sample gate floor(keep + U[0,1)) / keep, drawn independently for the
    attention and the MLP branch of every Block.  None when no Block drops (eval, or all rates 0)."""
    if not training or all(r == 0.0 for r in rates):
        return None
    key = (tuple(rates), str(device))
    keep = _KEEP_CACHE.get(key)
    if keep is None:       # built once, outside any CUDA-graph capture (the engine warms up eagerly first)
        keep = torch.tensor([1.0 - r for r in rates for _ in (0, 1)], dtype=torch.float32).to(device)
        _KEEP_CACHE[key] = keep
    seed = _DP["seed"]
    if seed is None:
        seed = torch.randint(0, 2 ** 62, (1,), dtype=torch.int64, device=device)
    _DP["calls"] += 1      # distinct stream per call site of a step (baked into a captured graph; the seed changes per step)
    return ops.drop_path_gates(seed, keep, batch, draw_id=_DP["calls"] & 0x7fffffff)


# ------------------------------------------------------------------------------------------- LayerNorm
class LayerNormFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, x, weight, bias, eps):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1]).contiguous().float()
        y, _, mean, rstd = ops.layernorm_fwd(x2, weight, bias, eps, out_dtype=torch.float32)
        ctx.save_for_backward(x2, mean, rstd, weight, bias)
        return y.view(shp)

    @staticmethod
    def backward(ctx, dy):
        x2, mean, rstd, weight, bias = ctx.saved_tensors
        sink = _GradSink()
        dx, _ = ops.layernorm_bwd(dy.reshape(x2.shape).contiguous().float(), x2, mean, rstd, weight,
                                  sink.get(weight, "w"), sink.get(bias, "b"))
        return dx.view(dy.shape), sink.result("w"), sink.result("b"), None


def layer_norm(x, weight, bias, eps=1e-5):
    return LayerNormFn.apply(x, weight, bias, eps)


# ---------------------------------------------------------------------------------------------- Linear
class LinearFn(torch.autograd.Function):
    """y = act(x W^T + b) on the tensor-core GEMM (K, N multiples of 8).  x may be f32 (cast on the way in) or already in the
    activation dtype (bf16 in the speed mode: no cast pass); out_act=True returns y in the activation dtype instead of f32, so
    that a chain Linear -> BatchNorm/ReLU -> Linear never widens to f32 between kernels; the input gradient comes back in the
    dtype the input had.  relu_mask (optional, activation dtype, the OUTPUT of a following ReLU): the incoming gradient is
    first masked by (relu_mask > 0) -- not used by the layers themselves, see BnReluFn."""

    @staticmethod
    def forward(ctx, x, weight, bias, gelu, out_act):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        adt = ops.act_dtype()
        xb = ops.to_act(x2)                                      # our cast kernel (no library launch in the step)
        odt = adt if out_act else torch.float32
        u = None
        if gelu:
            u = torch.empty(xb.shape[0], weight.shape[0], dtype=adt, device=x.device)
            y = ops.gemm(xb, shadow(weight), bias=bias, act=ops.ACT_GELU, preact_out=u, out_dtype=odt)
        else:
            y = ops.gemm(xb, shadow(weight), bias=bias, out_dtype=odt)
        ctx.save_for_backward(xb, weight, bias, u)
        ctx.x_needs = x.requires_grad
        ctx.x_dtype = x.dtype
        return y.view(*shp[:-1], weight.shape[0])

    @staticmethod
    def backward(ctx, dy):
        xb, weight, bias, u = ctx.saved_tensors
        sink = _GradSink()
        N = weight.shape[0]
        adt = ops.act_dtype()
        d2 = dy.reshape(-1, N).contiguous()
        if u is None and d2.dtype == adt and adt != torch.float32:
            g = d2                                               # already the GEMM operand: no cast pass
            if bias is not None:
                ops.colsum(g, sink.get(bias, "b"))
        else:
            d2 = d2.float()
            if u is not None:      # through the GELU first (needs its own pass: bias grad is of the pre-activation)
                uu = u.float()
                cdf = 0.5 * (1 + torch.erf(uu * 0.7071067811865476))
                d2 = d2 * (cdf + uu * torch.exp(-0.5 * uu * uu) * 0.3989422804014327)
            if N % 128 == 0 and N <= 1024:
                g = ops.cast_rows(d2, dbias=sink.get(bias, "b") if bias is not None else None)
            else:
                g = ops.to_act(d2)
                if bias is not None:
                    ops.colsum(d2, sink.get(bias, "b"))
        ops.wgrad(g, xb, sink.get(weight, "w"))
        dx = None
        if ctx.x_needs:
            dx = ops.gemm(g, shadow(weight), b_mn=True, out_dtype=ctx.x_dtype if ctx.x_dtype == adt else torch.float32)
            dx = dx.view(*dy.shape[:-1], weight.shape[1])
        return dx, sink.result("w"), sink.result("b") if bias is not None else None, None, None


def linear(x, weight, bias=None, gelu=False, out_act=False):
    return LinearFn.apply(x, weight, bias, gelu, out_act)


# ----------------------------------------------------------------------------- BatchNorm1d (train) + ReLU
class BnReluFn(torch.autograd.Function):
    """relu(BatchNorm1d(x)) over the rows of x [M, C] (channels last), train mode, on the kernels the mini-PointNet uses
    (csrc/pointnet.cu: per-channel sum / sum of squares, finalize + running-statistics update, normalise + ReLU; backward:
    ReLU mask folded into the two-pass BatchNorm backward).  x and the result are in the activation dtype (a following
    LinearFn consumes the result without a cast).  The FoldingNet decoder's final_conv (models/dvae.py:234-240)."""

    @staticmethod
    def forward(ctx, x, gamma, beta, running_mean, running_var, nbt, momentum, eps):
        M, C = x.shape
        adt = ops.act_dtype()
        xa = ops.to_act(x.contiguous())
        sm, sq = ops.bn_stats(xa)
        sc, sh, mean, rstd = ops.bn_finalize(sm, sq, M, gamma, beta, eps, momentum, running_mean, running_var, nbt)
        a = ops.bn_apply(xa, sc, sh, relu=True)
        ctx.save_for_backward(xa, a, mean, rstd, gamma, beta)
        ctx.x_dtype = x.dtype
        return a

    @staticmethod
    def backward(ctx, da):
        xa, a, mean, rstd, gamma, beta = ctx.saved_tensors
        sink = _GradSink()
        adt = ops.act_dtype()
        d = ops.to_act(da.contiguous())
        dz = ops.relu_mask_(d, a)                                # dz = da * (a > 0), in place on our own temporary
        dx, dbeta, dgamma = ops.bn_bwd(dz, xa, mean, rstd, gamma)
        ops.accumulate_(sink.get(gamma, "g"), dgamma)
        ops.accumulate_(sink.get(beta, "b"), dbeta)
        if ctx.x_dtype != adt:
            dx = dx.to(ctx.x_dtype)
        return dx, sink.result("g"), sink.result("b"), None, None, None, None, None


def bn_relu(x, bn, training=True):
    """x [M, C] -> relu(bn(x)) with bn an nn.BatchNorm1d (train mode: batch statistics + running-stat update)."""
    if not training:
        raise NotImplementedError("act_b200 BnReluFn: train-mode BatchNorm only (the Stage-I training step)")
    return BnReluFn.apply(x, bn.weight, bn.bias, bn.running_mean, bn.running_var, bn.num_batches_tracked, bn.momentum, bn.eps)


# ------------------------------------------------------------ Stage-I soft gumbel-softmax + KL(mean softmax || uniform)
class GumbelSoftmaxFn(torch.autograd.Function):
    """(y, qbar) = (softmax((logits + gumbel) / tau), mean over each cloud's G rows of softmax(logits)) for logits f32
    [B*G, V] -- F.gumbel_softmax(logits, tau, hard=False) of DiscreteVAE.forward (models/dvae.py:346) and the mean softmax
    get_loss needs (dvae.py:326), from ONE read of the logits per kernel (csrc/gumbel.cu).  y is in the activation dtype (the
    codebook GEMM's A operand, no cast pass).  The backward takes both upstream gradients (dy from the reconstruction path,
    dqbar from the KL term) and applies both softmax Jacobians in one pass.  tau: float or 1-element device tensor (no
    gradient).  noise: f32 [B*G, V] gumbel noise (parity runs) or None = drawn in-kernel from `seed` (device int64 [1])."""

    @staticmethod
    def forward(ctx, logits, tau, noise, seed, draw_id, B, G):
        logits = logits.contiguous()
        y, lse = ops.gumbel_softmax_fwd(logits, tau, noise, seed, draw_id)
        qbar = ops.softmax_colmean(logits, lse, B, G)
        ctx.save_for_backward(logits, lse, y)
        ctx.tau, ctx.G = tau, G
        ctx.set_materialize_grads(False)
        return y, qbar

    @staticmethod
    def backward(ctx, dy, dqbar):
        logits, lse, y = ctx.saved_tensors
        if dy is None and dqbar is None:
            return (None,) * 7
        dl = ops.gumbel_softmax_bwd(logits, lse, y, dy, ctx.tau, None if dqbar is None else dqbar.contiguous().float(), ctx.G)
        return dl, None, None, None, None, None, None


class KlUniformFn(torch.autograd.Function):
    """F.kl_div(log(qbar), log(1/V), reduction='batchmean', log_target=True) for qbar f32 [B, V] (dvae.py:327-331)."""

    @staticmethod
    def forward(ctx, qbar):
        qbar = qbar.contiguous()
        ctx.save_for_backward(qbar)
        return ops.kl_uniform_fwd(qbar)

    @staticmethod
    def backward(ctx, gout):
        qbar, = ctx.saved_tensors
        return ops.kl_uniform_bwd(qbar, gout.reshape(1))


class CodebookFn(torch.autograd.Function):
    """sampled = einsum('b g n, n c -> b g c', soft_one_hot, codebook) (dvae.py:347) for y [R, V] (activation dtype) and
    codebook f32 [V, C], reading the codebook where it lies: forward = GEMM with an MN-major B operand, dgrad = the plain
    K-major GEMM, wgrad = y^T . dout straight into the codebook's gradient -- no transposed or re-cast copies."""

    @staticmethod
    def forward(ctx, y, codebook):
        ctx.save_for_backward(y, codebook)
        return ops.gemm(y, shadow(codebook), b_mn=True, out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, dout):
        y, codebook = ctx.saved_tensors
        sink = _GradSink()
        C = codebook.shape[1]
        d2 = dout.contiguous().float()
        g = ops.to_act(d2)
        ops.wgrad(y, g, sink.get(codebook, "w"))
        dy = ops.gemm(g, shadow(codebook), out_dtype=y.dtype) if ctx.needs_input_grad[0] else None
        return dy, sink.result("w")


class EdgeLinearFn(torch.autograd.Function):
    """(P | Q) = x . [Wa ; Wb - Wa]^T for a DGCNN edge conv weight W [Cp, 2*Cin] (see dvae.dgcnn_forward): the token-level
    weight is built by one small kernel in the GEMM operand dtype and its gradient folded back into W's by another -- no
    slice / sub / cat / cast / zero-fill / add library passes per layer and step."""

    @staticmethod
    def forward(ctx, x, W):
        W = W.contiguous()
        xb = ops.to_act(x.reshape(-1, x.shape[-1]))
        Wp = ops.edge_weight_fwd(W.detach())
        ctx.save_for_backward(xb, W, Wp)
        ctx.x_dtype, ctx.x_needs = x.dtype, x.requires_grad
        return ops.gemm(xb, Wp, out_dtype=torch.float32)

    @staticmethod
    def backward(ctx, dy):
        xb, W, Wp = ctx.saved_tensors
        sink = _GradSink()
        g = ops.to_act(dy.contiguous())
        dWp = ops.zero_(torch.empty(Wp.shape, dtype=torch.float32, device=Wp.device))
        ops.wgrad(g, xb, dWp)
        ops.edge_weight_bwd(dWp, sink.get(W, "w"))
        dx = None
        if ctx.x_needs:
            dx = ops.gemm(g, Wp, b_mn=True, out_dtype=ctx.x_dtype if ctx.x_dtype == ops.act_dtype() else torch.float32)
        return dx, sink.result("w")


class FoldInputFn(torch.autograd.Function):
    """FoldingNet final_conv.0 over cat([global, seed, point]) (dvae.py:259-266) as z_g (per group) + Ws.seed (per grid
    cell) + Wp.coarse (per coarse point), csrc/folding.cu: one pass writes z in the activation dtype; the backward reduces
    dz to dz_g / dcoarse / the 5 weight columns in one pass.  weight: the conv's full f32 weight [C, c_g + 5] (a view of the
    parameter), only its last 5 columns take part here (the first c_g act through z_g)."""

    @staticmethod
    def forward(ctx, z_g, coarse, weight, seed):
        z_g, coarse = z_g.contiguous().float(), coarse.contiguous().float()
        c_g = weight.shape[1] - 5
        ctx.save_for_backward(coarse, weight, seed)
        return ops.fold_input_fwd(z_g, coarse, weight, c_g, seed)

    @staticmethod
    def backward(ctx, dz):
        coarse, weight, seed = ctx.saved_tensors
        sink = _GradSink()
        dw = sink.get(weight, "w")
        dz_g, dcoarse = ops.fold_input_bwd(ops.to_act(dz.contiguous()), coarse, weight, weight.shape[1] - 5, seed,
                                           dw.view(weight.shape))
        return dz_g, dcoarse, sink.result("w"), None


# --------------------------------------------------------------------------------------- pos-embed MLP
class PosMlpFn(torch.autograd.Function):
    """nn.Sequential(Linear(3,128), GELU, Linear(128,C)) on group centres (act.py:173-177, 1166-1170): the K = 3 layer +
    GELU on CUDA cores (csrc/tokens.cu) writing the bf16 operand of the 128 -> C tcgen05 GEMM; backward: cast + bias
    gradient, wgrad, dgrad, and the first layer's parameter gradients with the pre-activation recomputed from the centres
    (which carry no gradient: they come from FPS)."""

    @staticmethod
    def forward(ctx, x, w0, b0, w2, b2):
        shp = x.shape
        x2 = x.reshape(-1, 3).contiguous().float()
        a = ops.pos_mlp1_fwd(x2, w0, b0, out_dtype=ops.act_dtype())
        y = ops.gemm(a, shadow(w2), bias=b2, out_dtype=torch.float32)
        ctx.save_for_backward(x2, a, w0, b0, w2, b2)
        return y.view(*shp[:-1], w2.shape[0])

    @staticmethod
    def backward(ctx, dy):
        x2, a, w0, b0, w2, b2 = ctx.saved_tensors
        sink = _GradSink()
        C = w2.shape[0]
        g = ops.cast_rows(dy.reshape(-1, C).contiguous().float(), dbias=sink.get(b2, "b2"))
        ops.wgrad(g, a, sink.get(w2, "w2"))
        da = ops.gemm(g, shadow(w2), b_mn=True, out_dtype=ops.act_dtype())
        ops.pos_mlp1_bwd(da, x2, w0, b0, sink.get(w0, "w0"), sink.get(b0, "b0"))
        r = sink.result
        return None, r("w0"), r("b0"), r("w2"), r("b2")


def pos_mlp(seq, x):
    return PosMlpFn.apply(x, seq[0].weight, seq[0].bias, seq[2].weight, seq[2].bias)


# ----------------------------------------------------------------------------- cls / mask-token rows
class AssembleRowsFn(torch.autograd.Function):
    """cat(fill.expand(B,T-n,C), src) (fill_first: cls token / cls pos, act.py:287-290) or cat(src, fill.expand(B,T-n,C))
    (mask tokens, act.py:1222-1224) in one launch.  src: f32 [B, src_T, C] contiguous, of which every cloud's rows
    [src_off, src_off+n) are used (the decoder input reads the encoder output past its cls row without a slice copy).
    Backward: the src rows (zeros elsewhere) and the parameter row's accumulated gradient."""

    @staticmethod
    def forward(ctx, src, fill, B, n, T, fill_first, src_off):
        C = fill.numel()
        shp = src.shape
        src = src.reshape(B, -1, C).contiguous().float()
        out = ops.assemble_rows(src, fill, B, n, T, fill_first, src.shape[1], src_off)
        ctx.save_for_backward(fill)
        ctx.meta = (B, n, T, fill_first, src.shape[1], src_off, shp)
        return out

    @staticmethod
    def backward(ctx, dout):
        (fill,) = ctx.saved_tensors
        B, n, T, fill_first, src_T, src_off, shp = ctx.meta
        sink = _GradSink()
        need_src = ctx.needs_input_grad[0]
        dsrc = ops.assemble_rows_bwd(dout.contiguous().float(), B, n, T, fill_first, need_src, sink.get(fill, "f"), src_T,
                                     src_off)
        return (dsrc.view(shp) if need_src else None), sink.result("f"), None, None, None, None, None


def assemble_rows(src, fill, B, n, T, fill_first, src_off=0):
    """-> f32 [B,T,C]; src may be [B*n, C] / [B, n, C] (src_off 0) or [B, src_T, C] with src_off."""
    return AssembleRowsFn.apply(src, fill, B, n, T, fill_first, src_off)


class LayerNormRowsFn(torch.autograd.Function):
    """nn.LayerNorm on x[:, j0:j0+cnt] of x f32 [B,T,C] (TransformerDecoder.forward, act.py:144: the last
    return_token_num tokens) without the slice copy of the library path: one row-gather launch feeds the LayerNorm kernel;
    the backward scatters into a zero-padded [B,T,C] gradient with the assemble kernel."""

    @staticmethod
    def forward(ctx, x, weight, bias, eps, j0, cnt):
        B, T, C = x.shape
        xs = ops.gather_rows(x.contiguous().float(), None, j0, cnt).view(B * cnt, C)
        y, _, mean, rstd = ops.layernorm_fwd(xs, weight, bias, eps, out_dtype=torch.float32)
        ctx.save_for_backward(xs, mean, rstd, weight, bias)
        ctx.meta = (B, T, C, j0, cnt)
        return y.view(B, cnt, C)

    @staticmethod
    def backward(ctx, dy):
        xs, mean, rstd, weight, bias = ctx.saved_tensors
        B, T, C, j0, cnt = ctx.meta
        if j0 + cnt != T:
            raise NotImplementedError("LayerNormRowsFn backward: the selected rows must be the last ones")
        sink = _GradSink()
        dxs, _ = ops.layernorm_bwd(dy.reshape(B * cnt, C).contiguous().float(), xs, mean, rstd, weight,
                                   sink.get(weight, "w"), sink.get(bias, "b"))
        zero = _zero_row(C, dy.device)
        dx = ops.assemble_rows(dxs, zero, B, cnt, T, True, cnt, 0)
        return dx, sink.result("w"), sink.result("b"), None, None, None


_ZERO_ROWS = {}


def _zero_row(C, device):
    key = (C, str(device))
    if key not in _ZERO_ROWS:        # created during the eager warm-up, before any capture
        _ZERO_ROWS[key] = torch.zeros(C, dtype=torch.float32, device=device)
    return _ZERO_ROWS[key]


def layer_norm_rows(x, weight, bias, eps, j0, cnt):
    return LayerNormRowsFn.apply(x, weight, bias, eps, j0, cnt)


# --------------------------------------------------------------------------------- mini-PointNet Encoder
# BatchNorm2's batch statistics can ride on conv3's GEMM epilogue (ops.gemm(colstats=...): no second pass over the
# [B*G*k, 512] conv output, 268 MB less DRAM traffic per encoder).  MEASURED NEGATIVE on B200 (round 2): that GEMM is bound
# by its epilogue, not by HBM -- the fused kernel takes 213 us against 102 us + 62 us for GEMM + separate statistics pass,
# and the step is 0.05-0.1 ms slower (profiles/r2_fused_bn_stats_ab.txt).  Kept as a tested option, off by default.
FUSED_BN_STATS = os.environ.get("ACT_B200_FUSED_BN_STATS", "0") == "1"


class PointNetEncoderFn(torch.autograd.Function):
    """Encoder.forward (models/dvae.py:201-215) with train-mode BatchNorm1d: four tcgen05 GEMMs + the fused
    elementwise / reduction kernels of csrc/pointnet.cu.  The `cat([global, local])` + conv3 is computed
    hoisted: W3[:, :256] . global (per GROUP) is added to W3[:, 256:] . local (per point) in the GEMM epilogue."""

    @staticmethod
    def forward(ctx, nb, training, momentum, eps, bufs, n_keep, w1, b1, g1, be1, w2, b2, w3, b3, g2, be2, w4, b4):
        # nb [B,G,k,3] or flattened [BG,k,3]; n_keep (None = all): only the first n_keep groups' tokens are
        # produced -- conv4 and everything downstream of BatchNorm2 skip the other groups' rows (their tokens are
        # dead values when the caller drops masked groups), while both BatchNorms still see every point.
        if nb.dim() == 3:
            nb = nb.unsqueeze(0)
        B, G, k, _ = nb.shape
        M = B * G * k
        p = nb.reshape(M, 3).contiguous().float()
        rm1, rv1, nbt1, rm2, rv2, nbt2 = bufs
        W1 = w1.view(128, 3)
        if training:     # statistics, folding and running-stat update in one O(128) kernel
            Wf, bf, mean1, rstd1 = ops.pn_bn1_fold(ops.pn_moments(p), M, W1, b1, g1, be1, eps, momentum, rm1, rv1, nbt1)
        else:
            mean1, var1 = rm1.double(), rv1.double()
            rstd1 = torch.rsqrt(var1 + eps)
            s1 = g1.double() * rstd1
            Wf = (W1.double() * s1[:, None]).float().contiguous()
            bf = ((b1.double() - mean1) * s1 + be1.double()).float()
            mean1, rstd1 = mean1.float(), rstd1.float()
        a1 = ops.pn_conv1(p, Wf, bf, relu=True)                                   # [M,128]
        BG = B * G
        if k == 32:      # max over the group's 32 points fused into the GEMM epilogue (fp32 accumulators)
            adt = ops.act_dtype()
            gmax = torch.empty(BG, 256, dtype=adt, device=p.device)
            arg2 = torch.empty(BG, 256, dtype=torch.uint8, device=p.device)
            f2 = ops.gemm(a1, shadow(w2).view(256, 128), bias=b2, garg=arg2,
                          **({"gmax_f32": gmax} if adt == torch.float32 else {"gmax_bf16": gmax}))   # [M,256]
        else:
            f2 = ops.gemm(a1, shadow(w2).view(256, 128), bias=b2)
            gmax, _, arg2 = ops.group_max(f2, k)                                  # [BG,256]
        w3s = shadow(w3).view(512, 512)
        gpart = ops.gemm(gmax, w3s[:, :256], bias=b3, out_dtype=torch.float32)    # [BG,512]
        if training and FUSED_BN_STATS:   # BatchNorm2's batch statistics ride on conv3's epilogue: h3 is not re-read
            stats = torch.empty(2, 512, dtype=torch.float32, device=p.device)
            h3 = ops.gemm(f2, w3s[:, 256:], resid=gpart, resid_row_div=k, colstats=stats)     # [M,512]
            sc2, sh2, mean2, rstd2 = ops.bn_finalize(stats[0], stats[1], M, g2, be2, eps, momentum, rm2, rv2, nbt2)
        elif training:
            h3 = ops.gemm(f2, w3s[:, 256:], resid=gpart, resid_row_div=k)         # [M,512]
            sm, sq = ops.bn_stats(h3)
            sc2, sh2, mean2, rstd2 = ops.bn_finalize(sm, sq, M, g2, be2, eps, momentum, rm2, rv2, nbt2)
        else:
            h3 = ops.gemm(f2, w3s[:, 256:], resid=gpart, resid_row_div=k)         # [M,512]
            mean2, var2 = rm2.double(), rv2.double()
            rstd2 = torch.rsqrt(var2 + eps)
            sc2d = g2.double() * rstd2
            sc2, sh2 = sc2d.float(), (be2.double() - mean2 * sc2d).float()
            mean2, rstd2 = mean2.float(), rstd2.float()
        a3 = ops.bn_apply(h3, sc2, sh2, relu=True)                                # [M,512]
        C = w4.shape[0]
        GK = BG if n_keep is None else int(n_keep)
        a3k = a3[:GK * k]
        if k == 32:      # conv4's [M,C] output is never written: only its per-group max (+ arg-max) leaves the SM
            tokens = torch.empty(GK, C, dtype=torch.float32, device=p.device)
            arg4 = torch.empty(GK, C, dtype=torch.uint8, device=p.device)
            ops.gemm(a3k, shadow(w4).view(C, 512), bias=b4, gmax_f32=tokens, garg=arg4, no_out=True)
        else:
            f4 = ops.gemm(a3k, shadow(w4).view(C, 512), bias=b4)                  # [GK*k,C]
            _, tokens, arg4 = ops.group_max(f4, k, want_bf16=False, want_f32=True)
        ctx.save_for_backward(p, a1, f2, gmax, arg2, h3, a3, arg4, mean1, rstd1, mean2,
                              rstd2, w1, b1, g1, be1, w2, b2, w3, b3, g2, be2, w4, b4)
        ctx.meta = (B, G, k, C, training, GK)
        return tokens.view(B, G, C) if n_keep is None else tokens

    @staticmethod
    def backward(ctx, dtok):
        (p, a1, f2, gmax, arg2, h3, a3, arg4, mean1, rstd1, mean2, rstd2, w1, b1, g1, be1, w2, b2, w3, b3, g2, be2, w4,
         b4) = ctx.saved_tensors
        B, G, k, C, training, GK = ctx.meta
        if not training:
            raise RuntimeError("act_b200 Encoder: backward is implemented for train-mode BatchNorm only")
        sink = _GradSink()
        BG = B * G
        M = BG * k
        side = _SideStream(dtok.device)
        d = dtok.reshape(GK, C).contiguous().float()
        ops.colsum(d, sink.get(b4, "b4"))
        dF4 = ops.group_max_bwd(d, arg4, k)                                       # [GK*k,C] dense
        a3k = a3[:GK * k]
        gw4 = sink.get(w4, "w4").view(C, 512)
        side.run(lambda: ops.wgrad(dF4, a3k, gw4), dF4)
        if GK == BG:
            dZ3 = ops.gemm(dF4, shadow(w4).view(C, 512), b_mn=True, mul_in=a3, mul_mode=ops.MUL_RELU_MASK)
        else:            # rows of the dropped groups receive no gradient from conv4: dZ3 holds the first GK*k rows only
            dZ3 = ops.gemm(dF4, shadow(w4).view(C, 512), b_mn=True, mul_in=a3k, mul_mode=ops.MUL_RELU_MASK)
        dH3, dbe2, dg2 = ops.bn_bwd(dZ3, h3, mean2, rstd2, g2)
        ops.accumulate_(sink.get(g2, "g2"), dg2)
        ops.accumulate_(sink.get(be2, "be2"), dbe2)
        dGp_b, dGp_f = ops.group_sum(dH3, k, want_bf16=True, want_f32=True)       # [BG,512]
        ops.colsum(dGp_f, sink.get(b3, "b3"))
        gw3 = sink.get(w3, "w3").view(512, 512)
        side.run(lambda: (ops.wgrad(dH3, f2, gw3[:, 256:]), ops.wgrad(dGp_b, gmax, gw3[:, :256])), dH3, dGp_b)
        w3s = shadow(w3).view(512, 512)
        dgmax = ops.gemm(dGp_b, w3s[:, :256], b_mn=True, out_dtype=torch.float32)  # [BG,256]
        dF2 = ops.gemm(dH3, w3s[:, 256:], b_mn=True)                              # [M,256]
        ops.group_max_bwd(dgmax, arg2, k, out=dF2)
        gw2, gb2 = sink.get(w2, "w2").view(256, 128), sink.get(b2, "b2")
        side.run(lambda: (ops.wgrad(dF2, a1, gw2), ops.colsum(dF2, gb2)), dF2)
        dZ1 = ops.gemm(dF2, shadow(w2).view(256, 128), b_mn=True, mul_in=a1, mul_mode=ops.MUL_RELU_MASK)
        dbe1, dg1 = ops.pn_conv1_bwd(dZ1, p, w1.view(128, 3).contiguous(), b1, mean1, rstd1, g1,
                                     sink.get(w1, "w1").view(128, 3), sink.get(b1, "b1"))
        ops.accumulate_(sink.get(g1, "g1"), dg1)
        ops.accumulate_(sink.get(be1, "be1"), dbe1)
        side.join()
        r = sink.result
        return (None, None, None, None, None, None, r("w1"), r("b1"), r("g1"), r("be1"), r("w2"), r("b2"), r("w3"), r("b3"),
                r("g2"), r("be2"), r("w4"), r("b4"))


# --------------------------------------------------------------------------------------------- the loss
class CosineLossFn(torch.autograd.Function):
    """(1/B) sum_b (1 - mean_tok cos(student[b], teacher[b])) -- act.py:1243-1254 without the Python loop."""

    @staticmethod
    def forward(ctx, student, teacher):
        C = student.shape[-1]
        s2 = student.reshape(-1, C).contiguous().float()
        t2 = teacher.reshape(-1, C).contiguous().float()
        loss, grad = ops.cosine_loss(s2, t2, 1e-8, want_grad=True)
        ctx.save_for_backward(grad)
        ctx.shape = student.shape
        return loss.view(())

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        # the saved gradient is consumed once: scale it in place by the incoming d(loss) (1.0 in the training step)
        return ops.scale_by_(grad, dloss.reshape(1)).view(ctx.shape), None


class PointwiseLossFn(torch.autograd.Function):
    """config.loss = 'l2' / 'smoothl1' (act.py:1188-1191, 1255-1256): nn.MSELoss / nn.SmoothL1Loss, reduction 'mean'."""

    @staticmethod
    def forward(ctx, student, teacher, kind):
        loss, grad = ops.pointwise_loss(student, teacher, kind, want_grad=True)
        ctx.save_for_backward(grad)
        ctx.shape = student.shape
        return loss

    @staticmethod
    def backward(ctx, dloss):
        (grad,) = ctx.saved_tensors
        return ops.scale_by_(grad, dloss.reshape(1)).view(ctx.shape), None, None


def pointwise_loss(student, teacher, kind):
    return PointwiseLossFn.apply(student, teacher, kind)


def cosine_loss(student, teacher):
    return CosineLossFn.apply(student, teacher)


# ------------------------------------------------------------------------ Stage-I dVAE: trainable DGCNN layers
class DgcnnEdgeFn(torch.autograd.Function):
    """One DGCNN edge-conv layer after its token-level GEMM (models/dvae.py:91-93: GroupNorm(4) -> LeakyReLU(0.2) -> max
    over the k = 4 neighbours of the conv over cat(x_k - x_q, x_q)), forward and backward fused (csrc/dgcnn_train.cu)."""

    @staticmethod
    def forward(ctx, pq, idx4, gamma, beta, B, G, eps, slope):
        Cp = pq.shape[1] // 2
        pq = pq.contiguous()
        out, argj, stats = ops.dgcnn_edge_gn_train_fwd(pq, idx4, gamma, beta, B, G, Cp, eps, slope)
        ctx.save_for_backward(pq, idx4, argj, stats, gamma, beta)
        ctx.meta = (B, G, Cp, slope)
        return out

    @staticmethod
    def backward(ctx, dout):
        pq, idx4, argj, stats, gamma, beta = ctx.saved_tensors
        B, G, Cp, slope = ctx.meta
        sink = _GradSink()
        dpq = ops.dgcnn_edge_gn_train_bwd(pq, idx4, argj, stats, gamma, beta, dout, B, G, Cp, slope,
                                          sink.get(gamma, "g"), sink.get(beta, "b"))
        return dpq, None, sink.result("g"), sink.result("b"), None, None, None, None


class GroupNormRowsFn(torch.autograd.Function):
    """LeakyReLU(GroupNorm(4)(x)) over the rows of each cloud (DGCNN layer5, models/dvae.py:53-56), x f32 [B*R, C]."""

    @staticmethod
    def forward(ctx, x, gamma, beta, B, R, eps, slope):
        x = x.contiguous()
        out, stats = ops.gn_rows_train_fwd(x, gamma, beta, B, R, eps, slope)
        ctx.save_for_backward(x, stats, gamma, beta)
        ctx.meta = (B, R, slope)
        return out

    @staticmethod
    def backward(ctx, dy):
        x, stats, gamma, beta = ctx.saved_tensors
        B, R, slope = ctx.meta
        sink = _GradSink()
        dx = ops.gn_rows_train_bwd(x, stats, gamma, beta, dy, B, R, slope, sink.get(gamma, "g"), sink.get(beta, "b"))
        return dx, sink.result("g"), sink.result("b"), None, None, None, None
