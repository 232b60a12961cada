#!/usr/bin/env python
"""bench.py -- ACT Stage-II masked-point-modeling training step on B200 (BASELINE.json configs[1]/[3]).

One "step" = Group tokenizer (FPS + kNN) -> mini-PointNet embed -> 12-block student encoder -> 2-block decoder
-> frozen teacher forward (mini-PointNet + DGCNN x2 + gumbel/codebook + prompted ViT-B, no grad) -> proj head ->
cosine distillation loss -> backward -> (N>1: one NCCL all-reduce of the flat gradient) -> fused AdamW, on B=128
synthetic ShapeNet-shaped clouds per GPU (N=1024, G=64, k=32, d=384, mask 0.6, drop_path 0.1), bf16 tensor-core
operands / fp32 accumulation and master weights.  `--teacher synthetic` times the student-only step (SURVEY.md 8d
config 2(i)); the default includes the teacher with random weights (config 2(ii): pretrained weights are not
obtainable offline), and the line carries the student-only number as `student_only`.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--batch B]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value`: clouds/s with the batch already resident in HBM; `e2e`: the same step
called with a pinned HOST batch (H2D inside the timed region, D2H of the loss).  `--impl reference` times the
reference's path restated for the host CPU (oracle/ref_model.py + oracle/cpu_ref.c: the reference itself is
Python + CUDA-only extensions and cannot run on the box without a GPU build), all host threads.
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "clouds_per_sec_act_stage2_step"
UNIT = "clouds/s"
N_POINTS, N_GROUP, GROUP_SIZE, MASK_RATIO, DROP_PATH = 1024, 64, 32, 0.6, 0.1


def workload_cfg(batch, n_gpus, teacher="native"):
    t = ("frozen teacher forward included (mini-PointNet + DGCNN x2 + gumbel/codebook + VPT-deep ViT-B x12 on 128 tokens, "
         "random weights: pretrained ones are not obtainable offline)") if teacher == "native" else \
        "synthetic teacher target (student-only step)"
    return {"workload": "ACT Stage-II step: N=1024, G=64 x k=32, 12L d=384 + 2L decoder, mask 0.6, drop_path 0.1, "
                        "cosine loss, fwd+bwd+AdamW; " + t, "teacher": teacher,
            "batch_per_gpu": batch, "global_batch": batch * n_gpus, "n_points": N_POINTS, "num_group": N_GROUP,
            "group_size": GROUP_SIZE, "depth": 12, "embed_dim": 384, "mask_ratio": MASK_RATIO,
            "parallelism": f"dp{n_gpus}",
            "l2": "per-step activation working set (~3 GB) >> 126 MB L2; an extra 256 MB L2-flush write runs "
                  "between timed steps, outside the per-step event pairs"}


def peaks():
    p = {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0, "bf16_tflops_sustained": 1400.0, "src": "fallback"}
    try:
        p.update(json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json"))))
        p["src"] = "measured"
    except Exception:
        pass
    return p


def gemm_traffic():
    """DRAM bytes (read + write) per GEMM launch, averaged over the 256 launches of one step: from the committed ncu
    capture of the same step (profiles/r1_gemm_traffic.json, `dram__bytes_read.sum + dram__bytes_write.sum`); None
    when the capture is absent.  A profiler number, reported beside -- never instead of -- the timed ones."""
    try:
        return round(json.load(open(os.path.join(ROOT, "profiles", "r1_gemm_traffic.json")))["dram_bytes_per_launch"])
    except Exception:
        return None


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,clocks_event_reasons.hw_slowdown,"
         "clocks_event_reasons.hw_thermal_slowdown,clocks_event_reasons.sw_thermal_slowdown,"
         "clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.p = None
        try:
            self.p = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                       "-lms", "100", "-i", str(index)], stdout=subprocess.PIPE,
                                      stderr=subprocess.DEVNULL, text=True)
        except Exception:
            self.p = None

    def stop(self):
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.p.terminate()
        try:
            out, _ = self.p.communicate(timeout=5)
        except Exception:
            self.p.kill()
            out = ""
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in out.strip().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for n, v in zip(names, f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(n)
        # under-load samples: the upper half of the observed clocks
        sm_sorted = sorted(sm)
        load = sm_sorted[len(sm_sorted) // 2:] if sm_sorted else []
        return {"sm_mhz": statistics.median(load) if load else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm)}


# ------------------------------------------------------------------------------------------- our arm
def _log(msg):
    if os.environ.get("ACT_BENCH_VERBOSE"):
        print(f"[bench rank {os.environ.get('RANK', '0')} {time.strftime('%H:%M:%S')}] {msg}", file=sys.stderr, flush=True)


def run_ours(args):
    import torch.distributed as dist
    from act_b200 import _lib, dp, layers, models, ops
    from act_b200.data import synthetic_clouds

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if not torch.cuda.is_available():
        raise SystemExit("bench.py (impl=ours) needs a CUDA device: the act_b200 path has no CPU fallback")
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    _log("process group up")
    B = args.batch
    torch.manual_seed(0)
    np.random.seed(1234 + rank)
    cfg = models.default_config(mask_ratio=MASK_RATIO, drop_path_rate=DROP_PATH, num_group=N_GROUP,
                                group_size=GROUP_SIZE)
    model = models.ACT_PointDistillation(cfg, teacher="native" if args.teacher == "native" else None).to(dev).train()
    fp = layers.FlatParams(model, lr=1e-3, weight_decay=0.05, exclude=model.UNUSED_PARAMETERS)
    dp.broadcast_params(fp)

    n_batches = 4
    host = [synthetic_clouds(B, N_POINTS, seed=dp.shard_seed(20231017, rank, i)).pin_memory() for i in range(n_batches)]
    resident = [h.to(dev) for h in host]
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)

    from act_b200.engine import PretrainStep
    _log("model + flat params built; capturing")
    eng = PretrainStep(model, fp, B, N_POINTS, use_graph=not args.no_graph, device=dev).capture()
    _log("captured")

    def step(i):                                             # batch already resident in HBM
        return eng.run(resident[i % n_batches])

    def step_e2e(i):
        loss = eng.run(host[i % n_batches])                  # pinned HOST batch: H2D inside the timed region
        loss_host.copy_(loss, non_blocking=True)             # D2H of the step's result
        return loss

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def timed(fn, K):
        evs = []
        barrier()
        t0 = time.perf_counter()
        for i in range(K):
            flush.zero_()                                     # L2 flush, outside the event pair
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(i)
            b.record()
            evs.append((a, b))
        barrier()
        wall = time.perf_counter() - t0
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms], dtype=torch.float64, device=dev)
        if world > 1:
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return t.item() / K, wall

    for i in range(args.warmup):
        step(i)
    for i in range(max(1, args.warmup // 2)):
        step_e2e(i)
    barrier()

    _log("warm-up done")
    sampler = ClockSampler(local) if rank == 0 else None
    l0 = ops.LAUNCHES
    ms_step, wall = timed(step, args.steps)
    launches = eng.launches_per_step
    ms_e2e, _ = timed(step_e2e, args.steps)
    _log("timed regions done")
    clocks = sampler.stop() if sampler else None
    eng.flush()                                              # pipelined mode (N>1): the last step's pending update
    last_loss = float(loss_host.item())
    # the same step without the frozen teacher's forward (synthetic target): what the trainable path alone costs
    ms_student = None
    if args.teacher == "native":
        object.__setattr__(model, "teacher", models.SyntheticTeacher(384).to(dev))
        eng_s = PretrainStep(model, fp, B, N_POINTS, use_graph=not args.no_graph, device=dev).capture()
        for i in range(3):
            eng_s.run(resident[i % n_batches])
        ms_student, _ = timed(lambda i: eng_s.run(resident[i % n_batches]), args.steps)
        eng_s.flush()
        object.__setattr__(model, "teacher", model.dvae_tokenizer.forward_tokenizer_features)

    # ---- roofline of the dominant kernel: every launch of the tcgen05 GEMM kernels inside one step.
    # One eager step records each distinct GEMM call (shape, layouts, epilogue, its real operands); each distinct
    # call is then replayed 10x back to back from a CUDA graph and timed with CUDA events on the launching stream
    # (steady-state launch duration without host launch gaps); achieved = sum(count * 2MNK) / sum(count * time).
    pk = peaks()
    roof = None
    if rank == 0:
        calls = {}
        orig = ops.gemm

        def rec_gemm(a, b, **kw):
            out = orig(a, b, **kw)
            K_, M_ = (a.shape if kw.get("a_mn") else a.shape[::-1])
            N_ = b.shape[1] if kw.get("b_mn") else b.shape[0]
            epi = [k for k in ("bias", "preact_out", "mul_in", "resid", "row_scale", "gmax_f32", "gmax_bf16")
                   if kw.get(k) is not None]
            if kw.get("act", 0):
                epi.append("gelu" if kw["act"] == 1 else "relu")
            key = (f"{M_}x{N_}x{K_}{'/Amn' if kw.get('a_mn') else ''}{'/Bmn' if kw.get('b_mn') else ''}"
                   f"{'/splitK' + str(kw['splits']) if kw.get('splits', 1) > 1 else ''}"
                   f"{'+' + '+'.join(epi) if epi else ''}")
            c = calls.setdefault(key, [0, 2.0 * M_ * N_ * K_, a, b, dict(kw, out=kw.get("out", out) if not kw.get("no_out") else None)])
            c[0] += 1
            return out

        ops.gemm = rec_gemm
        try:
            eng._host_prologue(resident[0])
            eng._body_a()
            eng._body_b()
            torch.cuda.synchronize()
        finally:
            ops.gemm = orig
        rows = []
        for key, (cnt, fl, a, b, kw) in calls.items():
            g = torch.cuda.CUDAGraph()
            s_ = torch.cuda.Stream()
            s_.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(s_):
                orig(a, b, **kw)
            torch.cuda.current_stream().wait_stream(s_)
            with torch.cuda.graph(g):
                for _ in range(10):
                    orig(a, b, **kw)
            best = 1e9
            for _ in range(3):
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                g.replay()
                e1.record()
                torch.cuda.synchronize()
                best = min(best, e0.elapsed_time(e1) / 10)
            rows.append((key, cnt, fl, best))
        tot_ms = sum(c * t for _, c, _, t in rows)
        tot_fl = sum(c * f for _, c, f, _ in rows)
        n = sum(c for _, c, _, _ in rows)
        ach = tot_fl / (tot_ms * 1e-3) / 1e12
        peak = pk.get("bf16_tflops_sustained", pk["bf16_tflops"])
        top = [{"gemm": k, "launches_per_step": c, "us": round(t * 1e3, 1), "tflops": round(f / (t * 1e-3) / 1e12, 1)}
               for k, c, f, t in sorted(rows, key=lambda r: -r[1] * r[3])[:8]]
        if os.environ.get("ACT_BENCH_GEMM_TABLE"):
            with open(os.environ["ACT_BENCH_GEMM_TABLE"], "w") as f:
                json.dump([{"gemm": k, "launches_per_step": c, "us": round(t * 1e3, 2),
                            "tflops": round(fl / (t * 1e-3) / 1e12, 1)} for k, c, fl, t in
                           sorted(rows, key=lambda r: -r[1] * r[3])], f, indent=1)
        roof = {"kernel": "gemm_bf16_kernel + gemm_bf16_persistent_kernel (tcgen05/TMA GEMM, all launches of a step)",
                "bound": "tensor", "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s",
                "frac": round(ach / peak, 4), "peak_src": pk["src"] + " bf16 sustained", "launches_per_step": n,
                "flops_per_launch": tot_fl / n, "avg_launch_us": round(tot_ms * 1e3 / n, 2),
                "gemm_ms_per_step_serialised": round(tot_ms, 3),
                "note": "per-launch durations: each distinct GEMM call of the step replayed 10x from a CUDA graph on its "
                        "real operands, CUDA events; in the timed step the weight-gradient GEMMs additionally overlap "
                        "the dgrad chain on a second stream",
                "top_launches": top, "traffic": gemm_traffic()}
    if world > 1:
        dist.barrier()

    if rank == 0:
        cpu = (cpu_baseline(sample_batch=8, steps=1, teacher=args.teacher)
               if world == 1 and not args.no_cpu_baseline else None)
        clouds = B * world
        line = {"metric": METRIC, "value": round(clouds / (ms_step * 1e-3), 1), "unit": UNIT, "n_gpus": world,
                "steps": args.steps, "warmup": args.warmup, "ms_per_step": round(ms_step, 4),
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                "data": "synthetic", "config": workload_cfg(B, world, args.teacher), "impl": "ours",
                "e2e": {"value": round(clouds / (ms_e2e * 1e-3), 1), "unit": UNIT, "ms_per_step": round(ms_e2e, 4),
                        "h2d_bytes_per_step": int(host[0].numel() * 4 + B * N_GROUP + 32), "d2h_bytes_per_step": 4},
                "student_only": (None if ms_student is None else
                                 {"value": round(clouds / (ms_student * 1e-3), 1), "unit": UNIT,
                                  "ms_per_step": round(ms_student, 4),
                                  "what": "same step with a synthetic teacher target (no teacher forward)"}),
                "gpu_launches": int(launches), "cuda_graph": not args.no_graph, "pipelined": bool(eng.pipeline),
                "loss": last_loss,
                "wall_s_timed": round(wall, 3),
                "clocks": clocks, "roofline": roof}
        if cpu is not None:
            line["cpu_baseline"] = cpu
        emit(line)
    if world > 1:
        dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_student_step_time(batch, steps, warmup, threads, teacher="native"):
    """The reference's path restated for the host (oracle/): Group on the C oracle, fp32 PyTorch modules,
    torch.optim.AdamW with the reference's two parameter groups (tools/builder.py:37-55)."""
    from oracle import ref_model
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    np.random.seed(0)
    model = ref_model.ACTPointDistillationStudent(mask_ratio=MASK_RATIO).train()
    decay = [p for n, p in model.named_parameters() if not (p.dim() <= 1 or n.endswith(".bias") or "token" in n)]
    nodecay = [p for n, p in model.named_parameters() if (p.dim() <= 1 or n.endswith(".bias") or "token" in n)]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.05}, {"params": nodecay, "weight_decay": 0.0}],
                            lr=1e-3)
    pts = ref_model.synthetic_clouds(batch, N_POINTS)
    tnet = None
    if teacher == "native":
        from oracle import ref_teacher
        tnet = ref_teacher.TeacherFeatures().train()      # frozen, train mode like the reference (act.py:1151-1160)
        for p in tnet.parameters():
            p.requires_grad = False
    tfeat = torch.randn(batch, N_GROUP, 384)
    times, phases = [], {"group_teacher": 0.0, "student_forward": 0.0, "backward": 0.0, "adamw": 0.0}
    for i in range(warmup + steps):
        t0 = time.perf_counter()
        opt.zero_grad(set_to_none=True)
        if tnet is not None:
            with torch.no_grad():
                nb, center = model.group_divider(pts)
                tfeat = tnet.forward_tokenizer_features(nb, center, return_global=True)
        t1 = time.perf_counter()
        loss = model(pts, tfeat)
        t2 = time.perf_counter()
        loss.backward()
        t3 = time.perf_counter()
        opt.step()
        t4 = time.perf_counter()
        if i >= warmup:
            times.append(t4 - t0)
            for k, v in zip(phases, (t1 - t0, t2 - t1, t3 - t2, t4 - t3)):
                phases[k] += v * 1e3 / steps
    cpu_student_step_time.last_phases_ms = {k: round(v, 1) for k, v in phases.items()}
    return sum(times) / len(times)


def cpu_baseline(sample_batch=8, steps=1, teacher="native"):
    threads = os.cpu_count() or 1
    t = cpu_student_step_time(sample_batch, steps, 1, threads, teacher)
    return {"value": round(sample_batch / t, 3), "unit": UNIT, "cores": threads, "kind": "port",
            "phases_ms": getattr(cpu_student_step_time, "last_phases_ms", None),
            "sample": f"{steps} timed step(s) after 1 warm-up of the oracle restatement (fp32 PyTorch + C FPS/kNN) "
                      f"at batch {sample_batch} (same per-cloud workload; the full batch of 128 would take minutes)"}


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample = 8
    steps = max(1, min(args.steps, 3))
    warm = max(1, min(args.warmup, 1))
    t = cpu_student_step_time(sample, steps, warm, threads, args.teacher)
    world = max(1, args.gpus)
    val = round(sample / t, 3)
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": round(t * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_cfg(args.batch, world, args.teacher),
            "impl": "reference",
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "phases_ms": getattr(cpu_student_step_time, "last_phases_ms", None),
                             "sample": f"{steps} timed step(s) at batch {sample} per step (bounded sample of the "
                                       f"batch-128 workload), oracle restatement of the reference path on all host "
                                       f"threads; the reference's own native ops are CUDA-only"},
            "e2e": {"value": val, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    emit(line)


_REAL_STDOUT = None


def _guard_stdout():
    """stdout carries exactly ONE JSON line: anything else written to fd 1 (NCCL prints its version banner there from
    C code) is sent to stderr instead; emit() writes the line to the saved descriptor."""
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)


def emit(line):
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is None:
        sys.stdout.write(data.decode())
        sys.stdout.flush()
    else:
        os.write(_REAL_STDOUT, data)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--batch", type=int, default=128, help="clouds per GPU")
    ap.add_argument("--teacher", default="native", choices=["native", "synthetic"],
                    help="native: the frozen teacher's forward is part of the step (the reference's full Stage-II step); "
                         "synthetic: student-only step with a synthetic target")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-graph", action="store_true", help="launch the step eagerly instead of replaying a CUDA graph")
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == "ours" else args.warmup
    _guard_stdout()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_ours(args)


if __name__ == "__main__":
    main()
