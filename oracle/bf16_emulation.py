"""CPU emulation of the act_b200 mini-PointNet pipeline with bf16 rounding at exactly the storage points the CUDA
path has (act_b200/layers.py PointNetEncoderFn).  TEST INFRASTRUCTURE ONLY.

Purpose: separate "the kernels are wrong" from "bf16 compute moves the answer".  The Encoder's gradients pass
through two max-pools and two ReLUs whose winners/masks are discrete functions of the forward activations, so
rounding forward tensors to bf16 moves some gradients by 5-15 % relative to the fp32 reference
(oracle/ref_model.Encoder == models/dvae.py:185-215) on adversarially random upstream gradients -- while this
emulation, run in fp32 (bf16=False), matches the reference to 1e-3, and the GPU result matches the bf16 emulation
to ~1e-2.  `python -m oracle.bf16_emulation` prints the table quoted in DESIGN.md."""
import numpy as np
import torch


def emulate(nb, sd, dtok, bf16=True):
    """nb [B,G,k,3] f32, sd = Encoder state_dict, dtok [B,G,C] upstream gradient.
    Returns (tokens [B,G,C], {param name: gradient})."""
    q = (lambda t: t.bfloat16().float()) if bf16 else (lambda t: t)
    B, G, k, _ = nb.shape
    M, BG = B * G * k, B * G
    p = nb.reshape(M, 3)
    W1 = sd['first_conv.0.weight'].view(128, 3); b1 = sd['first_conv.0.bias']
    g1 = sd['first_conv.1.weight']; be1 = sd['first_conv.1.bias']
    W2 = q(sd['first_conv.3.weight'].view(256, 128)); b2 = sd['first_conv.3.bias']
    W3 = q(sd['second_conv.0.weight'].view(512, 512)); b3 = sd['second_conv.0.bias']
    g2 = sd['second_conv.1.weight']; be2 = sd['second_conv.1.bias']
    C = sd['second_conv.3.weight'].shape[0]
    W4 = q(sd['second_conv.3.weight'].view(C, 512)); b4 = sd['second_conv.3.bias']
    h1 = p @ W1.t() + b1
    m1 = h1.mean(0); rs1 = torch.rsqrt(h1.var(0, unbiased=False) + 1e-5)
    xh1 = (h1 - m1) * rs1
    a1 = q(torch.relu(xh1 * g1 + be1))
    f2f = a1 @ W2.t() + b2
    f2 = q(f2f)
    gmaxf, arg2 = f2f.view(BG, k, 256).max(1)          # arg-max on the fp32 accumulators (fused epilogue)
    gmax = q(gmaxf)
    gpart = gmax @ W3[:, :256].t() + b3
    h3 = q(f2 @ W3[:, 256:].t() + gpart.repeat_interleave(k, 0))
    m2 = h3.mean(0); rs2 = torch.rsqrt(h3.var(0, unbiased=False) + 1e-5)
    xh3 = (h3 - m2) * rs2
    a3 = q(torch.relu(xh3 * g2 + be2))
    f4 = a3 @ W4.t() + b4
    tok, arg4 = f4.view(BG, k, C).max(1)
    d = dtok.reshape(BG, C)
    dF4 = torch.zeros(BG, k, C).scatter_(1, arg4[:, None], q(d)[:, None]).view(M, C)
    dZ3 = q((dF4 @ W4) * (a3 > 0))
    s1 = dZ3.sum(0); s2 = (dZ3 * xh3).sum(0)
    dH3 = q(g2 * rs2 * (dZ3 - s1 / M - xh3 * s2 / M))
    dGp = dH3.view(BG, k, 512).sum(1)
    dGpb = q(dGp)
    dgmax = dGpb @ W3[:, :256]
    dF2 = q(dH3 @ W3[:, 256:])
    dF2 = q(dF2 + torch.zeros(BG, k, 256).scatter_(1, arg2[:, None], dgmax[:, None]).view(M, 256))
    dZ1 = q((dF2 @ W2) * (a1 > 0))
    dg1 = (dZ1 * xh1).sum(0); dbe1 = dZ1.sum(0)
    dh1 = g1 * rs1 * (dZ1 - dbe1 / M - xh1 * dg1 / M)
    grads = {
        'first_conv.0.weight': (dh1.t() @ p).view(128, 3, 1), 'first_conv.0.bias': dh1.sum(0),
        'first_conv.1.weight': dg1, 'first_conv.1.bias': dbe1,
        'first_conv.3.weight': (dF2.t() @ a1).view(256, 128, 1), 'first_conv.3.bias': dF2.sum(0),
        'second_conv.0.weight': torch.cat([dGpb.t() @ gmax, dH3.t() @ f2], 1).view(512, 512, 1),
        'second_conv.0.bias': dGp.sum(0), 'second_conv.1.weight': s2, 'second_conv.1.bias': s1,
        'second_conv.3.weight': (dF4.t() @ a3).view(C, 512, 1), 'second_conv.3.bias': d.sum(0),
    }
    return tok.view(B, G, C), grads


def main():
    import os
    from . import ref_model
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    g = np.load(os.path.join(root, "tests", "golden", "encoder.npz"))
    nb = torch.from_numpy(np.load(os.path.join(root, "tests", "golden", "group.npz"))["shapenet/neighborhood"][:2])
    enc = ref_model.fill_params(ref_model.Encoder(384), seed=2)
    rel = lambda a, b: ((a.reshape(-1) - b.reshape(-1)).norm() / b.norm()).item()   # noqa: E731
    for bf16 in (False, True):
        tok, grads = emulate(nb, enc.state_dict(), torch.from_numpy(g["wout"]), bf16)
        print("bf16" if bf16 else "fp32", "tokens", f"{rel(tok, torch.from_numpy(g['out'])):.1e}",
              {k: f"{rel(v, torch.from_numpy(g['grad/' + k])):.3f}" for k, v in grads.items() if "bias" not in k or k.endswith("1.bias") or k.endswith("3.bias") and "second" in k})


if __name__ == "__main__":
    main()
