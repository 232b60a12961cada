#!/usr/bin/env python
"""bench.py -- ACT masked-point-modeling training steps on B200 (BASELINE.json configs[1..4]).

Headline (config 2 / 4, `--config stage2`): one "step" = Group tokenizer (FPS + kNN) -> mini-PointNet embed -> 12-block
student encoder -> 2-block decoder -> frozen teacher forward (mini-PointNet + DGCNN x2 + gumbel/codebook + prompted
ViT-B, no grad) -> proj head -> cosine distillation loss -> backward -> (N>1: one NCCL all-reduce of the flat gradient)
-> fused AdamW, on B=128 synthetic ShapeNet-shaped clouds per GPU (N=1024, G=64, k=32, d=384, mask 0.6, drop_path 0.1),
bf16 tensor-core operands / fp32 accumulation and master weights.  The same line carries, under `configs`, the other two
GPU configs of BASELINE.json measured in the same process: `dvae` (config 3: Stage-I autoencoder step, B=64) and `dense`
(config 5: N=8192, G=512 x k=32, T_enc=206 / T_dec=512, B=16 per GPU; student step with a synthetic target -- the
teacher's ViT is defined for 64 tokens), and `sustained`: the headline step replayed back to back for >= 3 s.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--config all|stage2|dvae|dense]
  N>1: python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...

Prints ONE JSON line (rank 0).  `value`: clouds/s with the batch already resident in HBM (device-timed; the wall clock
of the same loop is reported beside it and the slower of the two is the value); `e2e`: the same step called with a pinned
HOST batch (H2D inside the timed region, D2H of the loss).  `--impl reference` times the reference's path restated for
the host CPU (oracle/ref_model.py + ref_teacher.py + cpu_ref.c, pinned against the unmodified reference modules in the
authoring container): the reference itself is Python over CUDA-only extensions and its sources cannot travel to the GPU
box, so `kind` is "port".
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
MASK_RATIO, DROP_PATH = 0.6, 0.1
SHAPES = {"stage2": dict(n_points=1024, num_group=64, group_size=32, batch=128),
          "dense": dict(n_points=8192, num_group=512, group_size=32, batch=16),
          "dvae": dict(n_points=1024, num_group=64, group_size=32, batch=64)}


def workload_cfg(batch, n_gpus, teacher="native", shape="stage2"):
    sh = SHAPES[shape]
    t = ("frozen teacher forward included (mini-PointNet + DGCNN x2 + gumbel/codebook + VPT-deep ViT-B x12 on 128 tokens, "
         "random weights: pretrained ones are not obtainable offline)") if teacher == "native" else \
        "synthetic teacher target (student-only step)"
    return {"workload": f"ACT Stage-II step: N={sh['n_points']}, G={sh['num_group']} x k={sh['group_size']}, 12L d=384 + "
                        f"2L decoder, mask 0.6, drop_path 0.1, cosine loss, fwd+bwd+AdamW; " + t, "teacher": teacher,
            "batch_per_gpu": batch, "global_batch": batch * n_gpus, "n_points": sh["n_points"],
            "num_group": sh["num_group"], "group_size": sh["group_size"], "depth": 12, "embed_dim": 384,
            "mask_ratio": MASK_RATIO, "parallelism": f"dp{n_gpus}",
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


def parity_report():
    """Errors of both precision modes against the reference goldens, as measured on a B200 by
    tests/test_gpu_parity.py::test_report_errors_of_both_modes (committed copy; None when absent)."""
    try:
        return json.load(open(os.path.join(ROOT, "profiles", "r2_parity_report.json")))
    except Exception:
        return None


def gemm_traffic():
    """DRAM bytes (read + write) per GEMM launch from the committed ncu capture of the same step (newest round first);
    None when absent.  A profiler number, reported beside -- never instead of -- the timed ones."""
    for name in ("r2_gemm_traffic.json", "r1_gemm_traffic.json"):
        try:
            return round(json.load(open(os.path.join(ROOT, "profiles", name)))["dram_bytes_per_launch"])
        except Exception:
            continue
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


class Harness:
    """Process-wide pieces shared by every config: ranks, barrier, the L2 flush buffer and the two timers."""

    def __init__(self, args):
        import torch.distributed as dist
        self.dist = dist
        self.world = int(os.environ.get("WORLD_SIZE", "1"))
        self.rank = int(os.environ.get("RANK", "0"))
        self.local = int(os.environ.get("LOCAL_RANK", "0"))
        if not torch.cuda.is_available():
            raise SystemExit("bench.py (impl=ours) needs a CUDA device: the act_b200 path has no CPU fallback")
        torch.cuda.set_device(self.local)
        self.dev = torch.device("cuda", self.local)
        if self.world > 1:
            dist.init_process_group("nccl", device_id=self.dev)
        self.flush = torch.empty(256 << 20, dtype=torch.uint8, device=self.dev)
        self.args = args
        self._flush_ms = None

    def barrier(self):
        if self.world > 1:
            self.dist.barrier()
        torch.cuda.synchronize()

    def flush_ms(self):
        """Device time of one L2-flush write, measured alone (subtracted from the wall clock of a timed loop)."""
        if self._flush_ms is None:
            for _ in range(2):
                self.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(5):
                self.flush.zero_()
            b.record()
            torch.cuda.synchronize()
            self._flush_ms = a.elapsed_time(b) / 5
        return self._flush_ms

    def timed(self, fn, K):
        """K steps, an L2 flush before each (outside the per-step event pair), barrier + synchronize on both sides.
        -> (device ms/step = sum of the per-step CUDA-event intervals, max over ranks;
            wall ms/step = host wall clock of the loop minus K flush writes, max over ranks).
        The second is what a training loop sees when the host (graph launches, mask loop, staging) is the limiter."""
        evs = []
        fl = self.flush_ms()
        self.barrier()
        t0 = time.perf_counter()
        for i in range(K):
            self.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn(i)
            b.record()
            evs.append((a, b))
        self.barrier()
        wall_ms = (time.perf_counter() - t0) * 1e3
        ms = sum(a.elapsed_time(b) for a, b in evs)
        t = torch.tensor([ms, wall_ms], dtype=torch.float64, device=self.dev)
        self.last_per_rank_ms = [round(ms / K, 4)]
        if self.world > 1:
            every = [torch.zeros_like(t) for _ in range(self.world)]
            self.dist.all_gather(every, t)
            self.last_per_rank_ms = [round(e[0].item() / K, 4) for e in every]      # device ms/step of every rank
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        return t[0].item() / K, max(0.0, t[1].item() / K - fl)

    def sustained(self, fn, seconds):
        """fn(i) back to back for >= `seconds` of device time (no flush: consecutive steps use different batches and a
        multi-GB working set), one event pair around the whole run, clocks sampled meanwhile."""
        sampler = ClockSampler(self.local) if self.rank == 0 else None
        self.barrier()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        n, t0 = 0, time.perf_counter()
        a.record()
        # the same step count on every rank (N>1: each step holds a collective): sized from a 10-step probe on rank 0
        for _ in range(10):
            fn(n)
            n += 1
        torch.cuda.synchronize()
        per = (time.perf_counter() - t0) / n
        more = torch.tensor([max(20, int(seconds / per) - n)], dtype=torch.int64, device=self.dev)
        if self.world > 1:
            self.dist.broadcast(more, 0)
        for _ in range(int(more.item())):
            fn(n)
            n += 1
        b.record()
        self.barrier()
        ms = a.elapsed_time(b)
        t = torch.tensor([ms], dtype=torch.float64, device=self.dev)
        if self.world > 1:
            self.dist.all_reduce(t, op=self.dist.ReduceOp.MAX)
        clocks = sampler.stop() if sampler else None
        return {"seconds": round(t.item() * 1e-3, 2), "steps": n, "ms_per_step": round(t.item() / n, 4),
                "sm_mhz_median": clocks["sm_mhz"] if clocks else None, "clock_reasons": clocks["reasons"] if clocks else None}


def gemm_roofline(record_step, pk, want_table=None):
    """Roofline of the dominant kernel family: every launch of the tcgen05 GEMM kernels inside one step.  One eager step
    (`record_step`) records each distinct GEMM call (shape, layouts, epilogue, its real operands); each distinct call is
    replayed 10x back to back from a CUDA graph and timed with CUDA events on the launching stream (steady-state launch
    duration without host gaps, at burst clocks: the denominator is therefore the BURST bf16 peak);
    achieved = sum(count * 2MNK) / sum(count * time)."""
    from act_b200 import ops
    calls = {}
    orig = ops.gemm

    def rec_gemm(a, b, **kw):
        out = orig(a, b, **kw)
        K_, M_ = (a.shape if kw.get("a_mn") else a.shape[::-1])
        N_ = b.shape[1] if kw.get("b_mn") else b.shape[0]
        epi = [k for k in ("bias", "preact_out", "mul_in", "resid", "row_scale", "gmax_f32", "gmax_bf16", "colstats")
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
        record_step()
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
    peak = pk["bf16_tflops"]
    srt = sorted(rows, key=lambda r: -r[1] * r[3])
    top = [{"gemm": k, "launches_per_step": c, "us": round(t * 1e3, 1), "tflops": round(f / (t * 1e-3) / 1e12, 1)}
           for k, c, f, t in srt[:8]]
    if want_table:
        with open(want_table, "w") as f:
            json.dump([{"gemm": k, "launches_per_step": c, "us": round(t * 1e3, 2),
                        "tflops": round(fl / (t * 1e-3) / 1e12, 1)} for k, c, fl, t in srt], f, indent=1)
    return {"kernel": "gemm_bf16_kernel + gemm_bf16_persistent_kernel + gemm_bf16_pair_kernel (tcgen05/TMA GEMM, all launches "
                      "of a step)",
            "bound": "tensor", "achieved": round(ach, 2), "peak": peak, "unit": "TFLOP/s",
            "frac": round(ach / peak, 4), "peak_src": pk["src"] + " bf16 burst (isolated-replay timings run at burst clocks)",
            "launches_per_step": n, "flops_per_launch": tot_fl / n, "flops_per_step": tot_fl,
            "avg_launch_us": round(tot_ms * 1e3 / n, 2), "gemm_ms_per_step_serialised": round(tot_ms, 3),
            "note": "per-launch durations: each distinct GEMM call of the step replayed 10x from a CUDA graph on its "
                    "real operands, CUDA events; in the timed step the weight-gradient GEMMs additionally overlap "
                    "the dgrad chain on a second stream",
            "top_launches": top, "traffic": gemm_traffic()}


def bench_stage2(h, shape, teacher, steps, warmup, want_roofline, want_student_only, want_sustained, no_graph=False):
    """The Stage-II step at one of SHAPES on this rank's GPU -> result dict (timings are max over ranks)."""
    from act_b200 import dp, layers, models, ops
    from act_b200.data import synthetic_clouds
    from act_b200.engine import PretrainStep
    sh = SHAPES[shape]
    B, NP = h.args.batch if (shape == "stage2" and h.args.batch) else sh["batch"], sh["n_points"]
    dev = h.dev
    torch.manual_seed(0)
    np.random.seed(1234 + h.rank)
    cfg = models.default_config(mask_ratio=MASK_RATIO, drop_path_rate=DROP_PATH, num_group=sh["num_group"],
                                group_size=sh["group_size"])
    model = models.ACT_PointDistillation(cfg, teacher="native" if teacher == "native" else "synthetic").to(dev).train()
    fp = layers.FlatParams(model, lr=1e-3, weight_decay=0.05, exclude=model.UNUSED_PARAMETERS)
    dp.broadcast_params(fp)
    n_batches = 4
    host = [synthetic_clouds(B, NP, seed=dp.shard_seed(20231017, h.rank, i)).pin_memory() for i in range(n_batches)]
    resident = [x.to(dev) for x in host]
    loss_host = torch.zeros((), dtype=torch.float32).pin_memory()
    _log(f"{shape}: model + flat params built; capturing")
    eng = PretrainStep(model, fp, B, NP, use_graph=not no_graph, device=dev).capture()
    _log(f"{shape}: captured")

    look = os.environ.get("ACT_BENCH_LOOKAHEAD") == "1" and eng.pipeline     # experiment: software pipelining across steps

    def step(i):                                             # batch already resident in HBM
        if look:
            return eng.run(resident[i % n_batches], next_points=resident[(i + 1) % n_batches])
        return eng.run(resident[i % n_batches])

    # e2e: every step copies ONE pinned host batch to the device and reads the loss back, inside the timed region.  Default:
    # the in-stream copy `eng.run(host_batch)` (the reference loop's `data.cuda()` at the top of the iteration).
    # ACT_BENCH_E2E_PREFETCH=1: the data-loader prefetch pattern of the engine's public API -- `eng.stage(host_batch)` starts
    # the H2D copy of step i+1's batch on the engine's copy stream before step i is launched, `eng.run(staged)` waits for
    # it on the device.  Measured: no difference (7.11 ms either way) -- the 1.6 MB copy was never what separates e2e from
    # the resident-input number; the e2e loop runs second, on a GPU already at its power cap (compare `sustained`).
    prefetch = os.environ.get("ACT_BENCH_E2E_PREFETCH", "0") == "1"
    staged = {}

    def step_e2e(i):
        if not prefetch:
            loss = eng.run(host[i % n_batches])              # pinned HOST batch: H2D inside the timed region
        else:
            cur = staged.pop("next", None)
            if cur is None:
                cur = eng.stage(host[i % n_batches])
            staged["next"] = eng.stage(host[(i + 1) % n_batches])      # H2D of the next step's batch, beside this step
            loss = eng.run(cur)
        loss_host.copy_(loss, non_blocking=True)             # D2H of the step's result
        return loss

    for i in range(warmup):
        step(i)
    for i in range(max(1, warmup // 2)):
        step_e2e(i)
    h.barrier()
    sampler = ClockSampler(h.local) if h.rank == 0 else None
    if os.environ.get("ACT_BENCH_E2E_FIRST") == "1":         # diagnostic: which of the two loops runs on the cooler GPU
        ms_e2e, wall_e2e = h.timed(step_e2e, steps)
        ms_step, wall_step = h.timed(step, steps)
        per_rank = list(h.last_per_rank_ms)
    else:
        ms_step, wall_step = h.timed(step, steps)
        per_rank = list(h.last_per_rank_ms)
        ms_e2e, wall_e2e = h.timed(step_e2e, steps)
    clocks = sampler.stop() if sampler else None
    sustained = h.sustained(step, h.args.sustain_seconds) if want_sustained else None
    eng.flush()                                              # pipelined mode (N>1): the last step's pending update
    torch.cuda.synchronize()
    res = {"batch": B, "ms_step": ms_step, "wall_step": wall_step, "ms_e2e": ms_e2e, "wall_e2e": wall_e2e,
           "clocks": clocks, "sustained": sustained, "launches": int(eng.launches_per_step),
           "pipelined": bool(eng.pipeline), "loss": float(loss_host.item()), "per_rank_ms": per_rank,
           "grad_comm": "bf16" if getattr(fp, "grad16", None) is not None else "fp32",
           "h2d": int(host[0].numel() * 4 + B * sh["num_group"] + 32)}
    # the same step without the frozen teacher's forward (synthetic target): what the trainable path alone costs
    if want_student_only and teacher == "native":
        object.__setattr__(model, "teacher", models.SyntheticTeacher(384).to(dev))
        eng_s = PretrainStep(model, fp, B, NP, use_graph=not no_graph, device=dev).capture()
        for i in range(3):
            eng_s.run(resident[i % n_batches])
        res["ms_student"], _ = h.timed(lambda i: eng_s.run(resident[i % n_batches]), steps)
        res["launches_student"] = int(eng_s.launches_per_step)
        eng_s.flush()
        del eng_s
        object.__setattr__(model, "teacher", model.dvae_tokenizer.forward_tokenizer_features)
    # tokenizer alone (FPS + kNN/gather): the latency-bound front of the step, first-order in the dense regime
    if h.rank == 0:
        ops.group(resident[0], sh["num_group"], sh["group_size"])
        evs = []
        for i in range(5):
            h.flush.zero_()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            ops.group(resident[i % n_batches], sh["num_group"], sh["group_size"])
            b.record()
            evs.append((a, b))
        torch.cuda.synchronize()
        res["tokenizer_us"] = round(1e3 * statistics.median(a.elapsed_time(b) for a, b in evs), 1)
    if want_roofline and h.rank == 0 and os.environ.get("ACT_BENCH_QUICK") != "1":
        def record():
            eng._host_prologue(resident[0])
            eng._body_a()
            eng._body_b()                 # one extra AdamW update outside the timed regions (recording only)
        res["roofline"] = gemm_roofline(record, peaks(), os.environ.get("ACT_BENCH_GEMM_TABLE") if shape == "stage2" else
                                        os.environ.get("ACT_BENCH_GEMM_TABLE_" + shape.upper()))
    if h.world > 1:
        h.dist.barrier()
    del eng, fp, model
    torch.cuda.empty_cache()
    return res


def bench_dvae(h, steps, warmup):
    """BASELINE config 3: the Stage-I dVAE step (tools/runner_autoencoder.py:130-146), B=64 per GPU."""
    from act_b200 import dp, dvae, engine, layers
    from act_b200.data import synthetic_clouds
    from act_b200.models import Cfg
    sh = SHAPES["dvae"]
    B = sh["batch"]
    cfg = Cfg(NAME="DiscreteVAE", group_size=32, num_group=64, num_tokens=8192, encoder_dims=256, tokens_dims=256,
              decoder_dims=256)
    torch.manual_seed(0)
    model = dvae.DiscreteVAE(cfg).to(h.dev).train()
    fp = layers.FlatParams(model, lr=5e-4, weight_decay=5e-4)
    dp.broadcast_params(fp)
    eng = engine.AutoencoderStep(model, fp, B, sh["n_points"], device=h.dev).capture()
    host = [synthetic_clouds(B, sh["n_points"], seed=dp.shard_seed(777, h.rank, i)).pin_memory() for i in range(4)]
    resident = [x.to(h.dev) for x in host]
    lh = torch.zeros(3, dtype=torch.float32).pin_memory()
    for i in range(warmup):
        eng.run(resident[i % 4])

    prefetch = os.environ.get("ACT_BENCH_E2E_PREFETCH", "0") == "1"
    staged = {}

    def e2e(i):
        if not prefetch:
            lh.copy_(eng.run(host[i % 4]), non_blocking=True)
            return
        cur = staged.pop("next", None)
        if cur is None:
            cur = eng.stage(host[i % 4])
        staged["next"] = eng.stage(host[(i + 1) % 4])        # H2D of the next step's batch, beside this step
        lh.copy_(eng.run(cur), non_blocking=True)

    e2e(0)
    ms, wall = h.timed(lambda i: eng.run(resident[i % 4]), steps)
    ms_e, wall_e = h.timed(e2e, steps)
    torch.cuda.synchronize()
    res = {"batch": B, "ms_step": ms, "wall_step": wall, "ms_e2e": ms_e, "wall_e2e": wall_e,
           "launches": int(eng.launches_per_step), "losses": [round(x, 5) for x in lh.tolist()],
           "h2d": int(host[0].numel() * 4 + 8 + 32)}
    if h.rank == 0:
        def record():
            eng._host_prologue(resident[0])
            eng._body_a()
        res["roofline"] = gemm_roofline(record, peaks(), os.environ.get("ACT_BENCH_GEMM_TABLE_DVAE"))
    if h.world > 1:
        h.dist.barrier()
    del eng, fp, model
    torch.cuda.empty_cache()
    return res


def _summ(res, world, unit=UNIT):
    """value = clouds/s from the slower of (device-event time, wall clock) -- see Harness.timed."""
    clouds = res["batch"] * world
    ms = max(res["ms_step"], res["wall_step"])
    ms_e = max(res["ms_e2e"], res["wall_e2e"])
    return {"value": round(clouds / (ms * 1e-3), 1), "unit": unit, "ms_per_step": round(ms, 4),
            "ms_per_step_device": round(res["ms_step"], 4), "ms_per_step_wall": round(res["wall_step"], 4),
            "batch_per_gpu": res["batch"],
            "e2e": {"value": round(clouds / (ms_e * 1e-3), 1), "unit": unit, "ms_per_step": round(ms_e, 4),
                    "ms_per_step_device": round(res["ms_e2e"], 4), "ms_per_step_wall": round(res["wall_e2e"], 4),
                    "h2d_bytes_per_step": res["h2d"], "d2h_bytes_per_step": 4 if "losses" not in res else 12,
                    "input_staging": ("engine.stage(): the pinned host batch of step i+1 is copied on a copy stream while step "
                                      "i runs (one H2D copy per step inside the timed region, consumed by run() through an "
                                      "event wait + a device-to-device copy)"
                                      if os.environ.get("ACT_BENCH_E2E_PREFETCH", "0") == "1" else
                                      "engine.run(host_batch): blocking in-stream H2D copy at the top of every step")},
            "gpu_launches": res["launches"]}


def _dense_line(res, world, pk):
    s = _summ(res, world)
    sh = SHAPES["dense"]
    s["config"] = workload_cfg(res["batch"], world, "synthetic", "dense")
    s["metric"] = "clouds_per_sec_act_stage2_step_dense"
    if "tokenizer_us" in res:
        B, N, G, k = res["batch"], sh["n_points"], sh["num_group"], sh["group_size"]
        nbytes = B * (12 * N + 16 * G) + B * (12 * N + 12 * G + 20 * G * k)       # FPS + kNN/gather
        s["tokenizer"] = {"us": res["tokenizer_us"], "share_of_step": round(res["tokenizer_us"] * 1e-3 / res["ms_step"], 4),
                          "algorithmic_bytes": nbytes,
                          "hbm_gbs": round(nbytes / (res["tokenizer_us"] * 1e-6) / 1e9, 1),
                          "frac_of_hbm_peak": round(nbytes / (res["tokenizer_us"] * 1e-6) / 1e9 / pk["hbm_gbs"], 5),
                          "note": "latency-bound (G-1 dependent arg-max rounds), not bandwidth-bound"}
    if "roofline" in res:
        s["roofline"] = res["roofline"]
    s["loss"] = res.get("loss")
    return s


def _dvae_line(res, world):
    s = _summ(res, world)
    s["metric"] = "clouds_per_sec_dvae_stage1_step"
    s["config"] = {"workload": "dVAE Stage-I step (BASELINE config 3): N=1024, G=64 x k=32, dims 256, 8192 tokens, "
                               "fwd + ChamferL1 x2 + KL + bwd + AdamW", "batch_per_gpu": res["batch"],
                   "global_batch": res["batch"] * world, "parallelism": f"dp{world}",
                   "l2": "256 MB L2-flush write between timed steps, outside the event pairs"}
    s["losses_recon_kl_total"] = res.get("losses")
    if "roofline" in res:
        s["roofline"] = res["roofline"]
    return s


def run_ours(args):
    h = Harness(args)
    _log("process group up")
    pk = peaks()
    world, rank = h.world, h.rank
    which = args.config
    line = None
    extra = {}
    if which in ("all", "stage2"):
        res = bench_stage2(h, "stage2", args.teacher, args.steps, args.warmup, want_roofline=True,
                           want_student_only=True, want_sustained=args.sustain_seconds > 0, no_graph=args.no_graph)
        s = _summ(res, world)
        if rank == 0:
            clouds = res["batch"] * world
            roof = res.get("roofline")
            sus = res["sustained"]
            if sus is not None and roof is not None:
                # step-level tensor throughput of the seconds-long run against the SUSTAINED cuBLAS figure
                tf = roof["flops_per_step"] / (sus["ms_per_step"] * 1e-3) / 1e12
                sus.update({"value": round(clouds / (sus["ms_per_step"] * 1e-3), 1), "unit": UNIT,
                            "gemm_tflops_step_level": round(tf, 1), "peak_sustained": pk["bf16_tflops_sustained"],
                            "frac_of_sustained_peak": round(tf / pk["bf16_tflops_sustained"], 4)})
            line = {"metric": METRIC, "value": s["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": s["ms_per_step"], "ms_per_step_device": s["ms_per_step_device"],
                    "ms_per_step_wall": s["ms_per_step_wall"],
                    "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "bf16",
                    "data": "synthetic", "config": workload_cfg(res["batch"], world, args.teacher), "impl": "ours",
                    "e2e": s["e2e"],
                    "student_only": (None if "ms_student" not in res else
                                     {"value": round(clouds / (res["ms_student"] * 1e-3), 1), "unit": UNIT,
                                      "ms_per_step": round(res["ms_student"], 4), "gpu_launches": res["launches_student"],
                                      "what": "same step with a synthetic teacher target (no teacher forward)"}),
                    "sustained": sus,
                    "gpu_launches": res["launches"], "cuda_graph": not args.no_graph, "pipelined": res["pipelined"],
                    "per_rank_ms_per_step": res["per_rank_ms"], "grad_comm": res["grad_comm"],
                    "loss": res["loss"], "clocks": res["clocks"], "roofline": roof,
                    "tokenizer_us": res.get("tokenizer_us"),
                    "precision": "bf16 operands, fp32 accumulate (speed mode; ACT_B200_PRECISION=fp32x3 selects the parity mode)",
                    "parity": parity_report()}
    for name in ("dvae", "dense"):
        if which not in ("all", name):
            continue
        try:
            if name == "dvae":
                r = bench_dvae(h, args.steps, args.warmup)
                extra[name] = _dvae_line(r, world) if rank == 0 else None
            else:
                r = bench_stage2(h, "dense", "synthetic", max(5, args.steps // 2), args.warmup, want_roofline=True,
                                 want_student_only=False, want_sustained=False)
                extra[name] = _dense_line(r, world, pk) if rank == 0 else None
        except Exception as e:                          # a secondary config must never take the headline line down
            if which != "all" or world > 1:             # (N>1: ranks must fail together, not diverge)
                raise
            extra[name] = {"error": f"{type(e).__name__}: {e}"[:300]}
    if rank == 0:
        if line is None:                                   # --config dvae / dense: that config IS the line
            name = which
            s = extra[name]
            line = {"metric": s.pop("metric"), "value": s["value"], "unit": UNIT, "n_gpus": world, "steps": args.steps,
                    "warmup": args.warmup, "ms_per_step": s["ms_per_step"], "higher_is_better": True, "scaling": "weak",
                    "vs_baseline": None, "dtype": "bf16", "data": "synthetic", "impl": "ours"}
            line.update({k: v for k, v in s.items() if k not in ("value", "unit", "ms_per_step")})
        else:
            line["configs"] = extra
            if world == 1 and not args.no_cpu_baseline:
                line["cpu_baseline"] = cpu_baseline(sample_batch=16, steps=1, teacher=args.teacher)
        emit(line)
    if world > 1:
        h.dist.destroy_process_group()


# --------------------------------------------------------------------------- CPU baseline / reference arm
def cpu_student_step_time(batch, steps, warmup, threads, teacher="native"):
    """The reference's path restated for the host (oracle/): Group on the C oracle, fp32 PyTorch modules,
    torch.optim.AdamW with the reference's two parameter groups (tools/builder.py:37-55).  -> median step time (s)."""
    from oracle import ref_model
    torch.set_num_threads(threads)
    torch.manual_seed(0)
    np.random.seed(0)
    model = ref_model.ACTPointDistillationStudent(mask_ratio=MASK_RATIO).train()
    decay = [p for n, p in model.named_parameters() if not (p.dim() <= 1 or n.endswith(".bias") or "token" in n)]
    nodecay = [p for n, p in model.named_parameters() if (p.dim() <= 1 or n.endswith(".bias") or "token" in n)]
    opt = torch.optim.AdamW([{"params": decay, "weight_decay": 0.05}, {"params": nodecay, "weight_decay": 0.0}],
                            lr=1e-3)
    pts = ref_model.synthetic_clouds(batch, SHAPES["stage2"]["n_points"])
    tnet = None
    if teacher == "native":
        from oracle import ref_teacher
        tnet = ref_teacher.TeacherFeatures().train()      # frozen, train mode like the reference (act.py:1151-1160)
        for p in tnet.parameters():
            p.requires_grad = False
    tfeat = torch.randn(batch, SHAPES["stage2"]["num_group"], 384)
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
    return statistics.median(times)


def cpu_baseline(sample_batch=16, steps=1, teacher="native"):
    threads = os.cpu_count() or 1
    t = cpu_student_step_time(sample_batch, steps, 1, threads, teacher)
    return {"value": round(sample_batch / t, 3), "unit": UNIT, "cores": threads, "kind": "port",
            "phases_ms": getattr(cpu_student_step_time, "last_phases_ms", None),
            "sample": f"{steps} timed step(s) after 1 warm-up of the oracle restatement (fp32 PyTorch + C FPS/kNN) "
                      f"at batch {sample_batch} (same per-cloud workload; the full batch of 128 would take minutes)"}


def run_reference(args):
    """The reference arm: the reference's own CPU implementation of the path on the box's host cores.  What runs is the
    oracle PORT (oracle/ref_model.py + ref_teacher.py + cpu_ref.c) -- pinned against the UNMODIFIED reference modules in
    the authoring container (tests/test_oracle*.py; scripts/ref_cpu_here.py times the two side by side there) -- because
    the reference's sources may not be copied into this repo and /root/reference does not exist on the GPU box.
    SURVEY 8(d): B=16 per step, 1 warm-up + 3 timed steps, median; the line states the steps / batch it REALLY ran."""
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    threads = os.cpu_count() or 1
    sample, steps, warm = 16, 3, 1
    t = cpu_student_step_time(sample, steps, warm, threads, args.teacher)
    world = max(1, args.gpus)
    val = round(sample / t, 3)
    cfg = workload_cfg(sample, 1, args.teacher)
    cfg["parallelism"] = f"host cpu, {threads} threads (one host: the arm does not scale with --gpus)"
    line = {"metric": METRIC, "value": val, "unit": UNIT, "n_gpus": world, "steps": steps,
            "warmup": warm, "ms_per_step": round(t * 1e3, 2), "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": cfg, "requested": {"steps": args.steps, "warmup": args.warmup},
            "impl": "reference",
            "cpu_baseline": {"value": val, "unit": UNIT, "cores": threads, "kind": "port",
                             "phases_ms": getattr(cpu_student_step_time, "last_phases_ms", None),
                             "sample": f"median of {steps} timed steps after {warm} warm-up at batch {sample} per step "
                                       f"(bounded sample of the batch-128 workload), oracle restatement of the reference "
                                       f"path on all host threads; the reference's own native ops are CUDA-only and its "
                                       f"sources cannot travel to this box"},
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
    ap.add_argument("--config", default="all", choices=["all", "stage2", "dvae", "dense"],
                    help="all: the Stage-II headline line with `configs.{dvae,dense}` attached; else that config alone")
    ap.add_argument("--batch", type=int, default=0, help="clouds per GPU of the stage2 config (default 128)")
    ap.add_argument("--teacher", default="native", choices=["native", "synthetic"],
                    help="native: the frozen teacher's forward is part of the step (the reference's full Stage-II step); "
                         "synthetic: student-only step with a synthetic target")
    ap.add_argument("--sustain-seconds", type=float, default=3.0, help="length of the back-to-back `sustained` run (0: skip)")
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
