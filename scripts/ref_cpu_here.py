"""Authoring-container-only evidence for bench.py's reference arm: the UNMODIFIED reference Stage-II model
(/root/reference/models/act.py ACT_PointDistillation incl. its frozen ACTPromptedDiscreteVAEwithVIT teacher, imported
through oracle/shims.py, FPS / kNN on the C oracle) timed on this container's CPU cores next to the oracle PORT that
`bench.py --impl reference` runs on the GPU box (where /root/reference does not exist and the reference sources may not
be copied).  Same batch, same threads, 1 warm-up + 3 timed steps (forward + backward + AdamW), median.

    python scripts/ref_cpu_here.py [batch] > profiles/r2_reference_vs_port_cpu.json
"""
import json
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
B = int(sys.argv[1]) if len(sys.argv) > 1 else 16
THREADS = os.cpu_count() or 1


def time_unmodified_reference():
    from oracle import shims
    from oracle.ref_model import synthetic_clouds
    shims.install()
    import models.act as act
    import models.dvae as dvae
    torch.set_num_threads(THREADS)
    torch.manual_seed(0)
    np.random.seed(0)
    cfg = shims.easydict(dict(
        NAME="ACT_PointDistillation", loss="cosine",
        transformer_config=dict(mask_ratio=0.6, mask_type="rand", proj="linear", embed_dim=384, encoder_dims=384,
                                depth=12, drop_path_rate=0.1, cls_dim=512, replace_pob=0.0, num_heads=6,
                                decoder_depth=2, decoder_num_heads=6, return_all_tokens=False, cls_loss=False,
                                register_shallow_hook=9),
        dvae_config=dict(num_group=64, group_size=32, encoder_dims=384, num_tokens=8192, tokens_dims=384,
                         decoder_dims=384, ckpt="", visual_embed_type="vit_base_patch16_384", visual_embed_dim=768,
                         freeze_visual_embed=True, num_prompt_token=64, use_deep_prompt=True)))

    def build_tokenizer(self, cfg_):          # the reference's build_tokenizer minus torch.load(ckpt): no ckpt offline
        self.dvae_tokenizer = dvae.ACTPromptedDiscreteVAEwithVIT(cfg_)
        for p in self.dvae_tokenizer.parameters():
            p.requires_grad = False

    act.ACT_PointDistillation.build_tokenizer = build_tokenizer
    model = act.ACT_PointDistillation(cfg).train()
    named = [(n, p) for n, p in model.named_parameters() if p.requires_grad]
    nd = lambda n, p: len(p.shape) == 1 or n.endswith(".bias") or "token" in n  # noqa: E731
    opt = torch.optim.AdamW([{"params": [p for n, p in named if nd(n, p)], "weight_decay": 0.0},
                             {"params": [p for n, p in named if not nd(n, p)], "weight_decay": 0.05}], lr=1e-3)
    pts = synthetic_clouds(B, 1024)
    ts = []
    for i in range(4):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        loss = model(pts)
        loss.backward()
        opt.step()
        if i:
            ts.append(time.perf_counter() - t0)
    return statistics.median(ts), float(loss.item())


def main():
    sys.path.insert(0, ROOT)
    import bench
    t_ref, loss = time_unmodified_reference()
    t_port = bench.cpu_student_step_time(B, 3, 1, THREADS, "native")
    print(json.dumps({"where": "authoring container (no GPU)", "threads": THREADS, "batch": B,
                      "unmodified_reference_through_shims": {"s_per_step": round(t_ref, 3), "clouds_per_s": round(B / t_ref, 2),
                                                             "loss": loss},
                      "oracle_port (what bench.py --impl reference runs)": {"s_per_step": round(t_port, 3),
                                                                            "clouds_per_s": round(B / t_port, 2),
                                                                            "phases_ms": bench.cpu_student_step_time.last_phases_ms},
                      "port_over_reference_time": round(t_port / t_ref, 3)}, indent=1))


if __name__ == "__main__":
    main()
