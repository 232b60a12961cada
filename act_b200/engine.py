"""The Stage-II training step as ONE replayable CUDA graph.

Reference hot loop: /root/reference/tools/runner_pretrain.py:130-157 (points.cuda(); loss = base_model(points);
loss.backward(); optimizer.step(); zero_grad) -- ~460 kernel launches per step from Python, plus four O(B) host
loops and several host<->device syncs.  Here the whole step -- Group tokenizer, mini-PointNet,
encoder, decoder, loss, backward and the fused AdamW -- is captured once into a CUDA
graph (two graphs around the NCCL gradient all-reduce when N>1) (CUDA streams and graphs instead of a tracing compiler) and replayed; per step the host only (a) draws the
random mask exactly as the reference does (numpy RNG, act.py:244-267) into a pinned buffer, (b) stages the AdamW
scalars, (c) copies the batch into the graph's static input, (d) launches the graph.  Nothing synchronises.
"""
import torch

from . import dp, layers, ops
from .modules import mask_center_rand


class _StateSnapshot:
    """Everything a warm-up step mutates, so capture() can leave the model exactly as it found it (a loaded checkpoint
    must not be perturbed by the two eager warm-up steps the allocator / lazy initialisers need before a graph capture):
    master weights, AdamW moments and step count, every module buffer (BatchNorm running statistics and counters of the
    student AND of the frozen teacher), the numpy global RNG (the mask draw) and the torch CPU / CUDA generators
    (DropPath gates, gumbel seed)."""

    def __init__(self, model, fp, dev):
        import numpy as np
        self.np = np
        self.model, self.fp, self.dev = model, fp, dev
        self.flat = fp.flat.clone()
        self.exp_avg, self.exp_avg_sq = fp.exp_avg.clone(), fp.exp_avg_sq.clone()
        self.step_count = fp.step_count
        mods = [model]
        t = getattr(model, "teacher", None)
        t = getattr(t, "__self__", t)                          # bound method of the (registered or external) teacher module
        if isinstance(t, torch.nn.Module):
            mods.append(t)
        self.bufs, seen = [], set()
        for m in mods:
            for b in m.buffers():
                if b.data_ptr() not in seen:
                    seen.add(b.data_ptr())
                    self.bufs.append((b, b.clone()))
        self.np_state = np.random.get_state()
        import random
        self.py_state = random.getstate()                      # block masking draws from Python's `random`
        self.cpu_rng = torch.get_rng_state()
        self.cuda_rng = torch.cuda.get_rng_state(dev)

    def restore(self):
        fp = self.fp
        with torch.no_grad():
            fp.flat.copy_(self.flat)
            fp.exp_avg.copy_(self.exp_avg)
            fp.exp_avg_sq.copy_(self.exp_avg_sq)
            fp.step_count = self.step_count
            fp.refresh_shadow()
            fp.zero_grad()
            for b, saved in self.bufs:
                b.copy_(saved)
        self.np.random.set_state(self.np_state)
        import random
        random.setstate(self.py_state)
        torch.set_rng_state(self.cpu_rng)
        torch.cuda.set_rng_state(self.cuda_rng, self.dev)


class _InputStaging:
    """Data-loader style input prefetch shared by the two step engines.  `stage(host_batch)` starts the host->device copy of
    a pinned host batch on the engine's copy stream into one of two device slots and returns that device tensor; passing
    it to `run()` makes the consuming stream wait for the copy and then move it into the step's static input buffer (a
    1.5 MB device-to-device copy).  Calling `stage(batch[i+1])` before `run(staged[i])` overlaps the next batch's PCIe
    transfer with the current step (the reference's loop does a blocking `points = data.cuda()` at the top of every
    iteration, runner_pretrain.py:128-131).  Plain `run(host_batch)` still copies in-stream."""
    _copy_stream = None

    def stage(self, points):
        if points.is_cuda:
            return points
        dev = self.dev
        if self._copy_stream is None:
            self._copy_stream = torch.cuda.Stream(device=dev)
            self._slots = [{"buf": torch.empty_like(self.points), "ready": torch.cuda.Event(), "consumed": None}
                           for _ in range(2)]
            self._slot_i = 0
        slot = self._slots[self._slot_i % 2]
        self._slot_i += 1
        cs = self._copy_stream
        if slot["consumed"] is not None:
            cs.wait_event(slot["consumed"])                    # the step that read this slot has copied it out
        with torch.cuda.stream(cs):
            slot["buf"].copy_(points, non_blocking=True)
            slot["ready"].record(cs)
        return slot["buf"]

    def _copy_in(self, points):
        """points -> the static input buffer on the CURRENT stream (no-op when it already is that buffer)."""
        if points is None or points.data_ptr() == self.points.data_ptr():
            return
        cur = torch.cuda.current_stream(self.dev)
        slot = None
        if self._copy_stream is not None:
            slot = next((sl for sl in self._slots if sl["buf"] is points), None)
        if slot is not None:
            cur.wait_event(slot["ready"])
        self.points.copy_(points, non_blocking=True)           # H2D when `points` is a pinned host batch
        if slot is not None:
            slot["consumed"] = torch.cuda.Event()
            slot["consumed"].record(cur)


class PretrainStep(_InputStaging):
    """pipeline (default: on when world > 1): the tokenizer and the FROZEN teacher's forward do not depend on the student's
    weights, so (1) step i's gradient all-reduce and AdamW are deferred to the start of step i+1, where the all-reduce runs
    on its own stream beside the teacher branch, and (2) with `run(points, next_points=...)` the tokenizer + teacher forward
    of step i+1 are issued on the teacher stream at the START of step i (software pipelining across steps: the teacher's
    large GEMMs fill the SMs that the student's small latency-bound kernels -- forward AND backward -- leave idle):
        S_n : all-reduce(grad i-1)
        main: copy(nb, center, tfeat <- next) . wait(S_n) [G3 AdamW(i-1)] [G2a student forward(i)] [G2b loss + backward(i)]
        T   :   wait(copy) [H2D points(i+1)] [G0 Group(i+1)] [G1 teacher forward(i+1) ........................]
    Without next_points the teacher of step i is issued at the start of step i and joined before the loss (G2a overlaps it).
    Five graphs (forward and backward of the student are captured separately, the autograd graph spanning both); G0 / G1
    write a "next" buffer set from their own memory pool because they replay concurrently with G3 / G2a / G2b, which read
    the "current" set.  Same arithmetic as the serial order (every student forward sees the weights updated by all earlier
    steps, every loss the teacher features of its own batch); call flush() to apply the last pending update (before
    reading parameters, saving a checkpoint, or switching engines)."""

    def __init__(self, model, flat_params, batch, n_points, use_graph=True, device=None, pipeline=None, device_mask=False):
        import os
        self.model, self.fp = model, flat_params
        if pipeline is None:
            env = os.environ.get("ACT_B200_PIPELINE")
            pipeline = (dp.world_size() > 1) if env is None else env == "1"
        self.pipeline = bool(pipeline) and use_graph
        self._pending = False
        self._prefetched = None                              # the announced batch (the tensor itself, kept alive) whose
                                                             # tokenizer + teacher forward are in flight
        self._nccl_stream = None
        self.dev = device or flat_params.flat.device
        self.B, self.N = batch, n_points
        self.G = model.num_group
        self.mask_ratio = model.mask_ratio
        self.points = torch.zeros(batch, n_points, 3, dtype=torch.float32, device=self.dev)
        self.mask = torch.zeros(batch, self.G, dtype=torch.bool, device=self.dev)
        self.loss = torch.zeros((), dtype=torch.float32, device=self.dev)
        # per-step 64-bit seeds staged from the host like the mask: [0] DropPath gates, [1] the teacher's gumbel / prompt
        # dropout draws (single-graph mode) -- the captured step then contains no library RNG kernel
        self._seeds = torch.zeros(3, dtype=torch.int64, device=self.dev)
        # device_mask: draw the random mask inside the captured step (csrc/augment.cu act_mask_rand, seed [2]) instead of the
        # reference's host loop of B numpy shuffles (act.py:255-264).  Same distribution, not the reference's random stream:
        # off by default, so that a seeded run masks the same groups as the reference does.
        self.device_mask = bool(device_mask)
        # mask_type 'block' (act.py:215-243): the mask depends on the centres, so it is formed inside the step from the
        # random centre indices, which are drawn on the host with Python's `random` like the reference and staged here
        enc = getattr(model, "ACT_encoder", None)
        self.block_mask = getattr(enc, "mask_type", "rand") == "block"
        if self.block_mask:
            self._block_idx = torch.zeros(batch, dtype=torch.int32, device=self.dev)
            enc.block_index = self._block_idx
        self.use_graph = use_graph
        self.graph = None
        self.graph_b = None
        self.launches_per_step = None
        if dp.world_size() > 1:
            self.fp.enable_bf16_comm()

    # the device work of one step (capturable: no host sync, no pageable copy), in two halves so that for N>1 the
    # step's one collective sits BETWEEN two graphs (NCCL's watchdog thread and stream capture do not mix safely;
    # the all-reduce is one eager launch on the same stream, ordered after graph A and before graph B)
    def _body_a(self):
        self.fp.zero_grad()
        with layers.drop_path_seed(self._seeds[:1]):
            loss = self.model(self.points, mask=self._mask())
        loss.backward()
        self.loss.copy_(loss.detach())

    def _mask(self):
        if self.block_mask:
            return None                              # formed inside the encoder from the staged indices
        if self.device_mask:
            return ops.mask_rand(self._seeds[2:], self.B, self.G, int(self.mask_ratio * self.G))
        return self.mask

    def _body_b(self):
        self.fp.step()                              # fused AdamW (+ bf16 shadow refresh); scalars read from device

    def _body(self):
        self._body_a()
        dp.sync_gradients(self.fp)                  # N>1: flat fp32 gradient all-reduce (NCCL over NVLink)
        self._body_b()

    # pipelined mode: the step in five graphs
    def _body_group(self):                           # G0 -> the "next" neighbourhoods / centres
        with torch.no_grad():
            self._nb_n, self._center_n = self.model.group_divider(self.points)

    def _body_teacher(self):                         # G1: frozen teacher on G0's outputs -> the "next" features
        with torch.no_grad():
            self._tfeat_n = self.model.teacher(self._nb_n, self._center_n)

    def _body_fwd(self):                             # G2a: student forward (autograd graph kept for G2b)
        self.fp.zero_grad()
        with layers.drop_path_seed(self._seeds[:1]):
            self._student, self._order, self._nvis = self.model.forward_student(self._nb, self._center, self._mask())

    def _body_bwd(self):                             # G2b: loss against the teacher's features + backward
        loss = self.model.distill_loss(self._student, self._tfeat, self._order, self._nvis)
        loss.backward()
        self.loss.copy_(loss.detach())
        self._student = None

    def _host_prologue(self, points, hyper=True):
        m = None if (self.device_mask or self.block_mask) else mask_center_rand(self.B, self.G, self.mask_ratio, "cpu")
        if self.block_mask:
            import random
            idx = torch.tensor([random.randint(0, self.G - 1) for _ in range(self.B)], dtype=torch.int32)
            self._block_idx.copy_(idx.pin_memory(), non_blocking=True)
        # a FRESH pinned staging tensor per step: the host runs many replays ahead of the GPU, and a reused staging buffer
        # would be overwritten before its asynchronous copy has executed (torch's caching host allocator recycles a
        # pinned block only after the copies recorded on it have completed)
        if m is not None:
            self.mask.copy_(m.pin_memory(), non_blocking=True)
        self._seeds.copy_(torch.randint(0, 2 ** 62, (3,), dtype=torch.int64).pin_memory(), non_blocking=True)
        if hyper:
            self.fp.set_hyper(grad_scale=1.0 / dp.world_size())
        self._copy_in(points)

    def _capture_pipeline(self):
        l0 = ops.LAUNCHES
        snap = _StateSnapshot(self.model, self.fp, self.dev)
        # G1 replays concurrently with G3 / G2a, which draw the DropPath gates from torch's default CUDA generator: the
        # teacher's per-call seed is therefore staged from the host (like the mask and the AdamW scalars), not drawn in G1
        tmod = getattr(getattr(self.model, "teacher", None), "__self__", None)
        self._tseed = None
        if tmod is not None and hasattr(tmod, "seed_buffer"):
            self._tseed = torch.zeros(1, dtype=torch.int64, device=self.dev)
            tmod.seed_buffer = self._tseed
        self._host_prologue(None)
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._body_group()
                self._body_teacher()
                self._nb, self._center, self._tfeat = self._nb_n.clone(), self._center_n.clone(), self._tfeat_n.clone()
                self._body_fwd()
                self._body_bwd()
                dp.sync_gradients(self.fp)
                self._body_b()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        snap.restore()                                       # the warm-up leaves no trace (see capture())
        self.launches_per_step = (ops.LAUNCHES - l0) // 2
        self.graph = torch.cuda.CUDAGraph()                  # G0 \ the teacher stream's graphs share one pool of their own:
        with torch.cuda.graph(self.graph):                   #    } they replay concurrently with the main stream's
            self._body_group()
        self.graph_t = torch.cuda.CUDAGraph()                # G1 /
        with torch.cuda.graph(self.graph_t, pool=self.graph.pool()):
            self._body_teacher()
        self.graph_f = torch.cuda.CUDAGraph()                # G2a
        with torch.cuda.graph(self.graph_f):
            self._body_fwd()
        self.graph_s = torch.cuda.CUDAGraph()                # G2b
        with torch.cuda.graph(self.graph_s, pool=self.graph_f.pool()):
            self._body_bwd()
        self.graph_b = torch.cuda.CUDAGraph()                # G3 (replayed BEFORE G2a: not in the shared pool's order)
        with torch.cuda.graph(self.graph_b):
            self._body_b()
        self._teacher_stream = torch.cuda.Stream(device=self.dev)
        self._ev_group, self._ev_teacher, self._ev_copied = torch.cuda.Event(), torch.cuda.Event(), torch.cuda.Event()
        self._ev_copied.record(torch.cuda.current_stream(self.dev))
        if dp.world_size() > 1:
            self._nccl_stream = torch.cuda.Stream(device=self.dev)
        return self

    def _enqueue_teacher(self, points):
        """Tokenizer + frozen teacher of `points` on the teacher stream, into the "next" buffer set."""
        T = self._teacher_stream
        T.wait_event(self._ev_copied)                        # the previous contents have been copied out
        with torch.cuda.stream(T):
            self._copy_in(points)
            if self._tseed is not None:                      # fresh pinned staging per step (see _host_prologue)
                self._tseed.copy_(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).pin_memory(), non_blocking=True)
            self.graph.replay()                              # G0
            self._ev_group.record(T)
            self.graph_t.replay()                            # G1
            self._ev_teacher.record(T)

    def _run_pipeline(self, points, next_points=None):
        main = torch.cuda.current_stream(self.dev)
        if self._pending and self._nccl_stream is not None:
            self._nccl_stream.wait_stream(main)              # the previous step's backward (G2b) is complete
            with torch.cuda.stream(self._nccl_stream):
                dp.sync_gradients(self.fp)
        self._host_prologue(None, hyper=False)               # this step's mask
        # tokenizer + teacher of THIS batch were issued a step ago (a different batch than announced: start over)
        # identity of the tensor object, not its address: the caching allocators recycle addresses, so a freed batch and a
        # later one can share a data_ptr; the reference held in _prefetched keeps the announced batch alive meanwhile
        ahead = self._prefetched is not None and self._prefetched is points
        if not ahead:
            self._enqueue_teacher(points)
        main.wait_event(self._ev_group)
        self._nb.copy_(self._nb_n)
        self._center.copy_(self._center_n)
        if ahead:
            main.wait_event(self._ev_teacher)
            self._tfeat.copy_(self._tfeat_n)
            self._ev_copied.record(main)
            if next_points is not None:
                self._enqueue_teacher(next_points)           # runs beside this step's AdamW / forward / backward
        if self._pending:
            if self._nccl_stream is not None:
                main.wait_stream(self._nccl_stream)
            self.fp.set_hyper(grad_scale=1.0 / dp.world_size())
            self.graph_b.replay()                            # G3: AdamW of the previous step
        self.graph_f.replay()                                # G2a
        if not ahead:
            main.wait_event(self._ev_teacher)                # teacher of this step: joined only before the loss
            self._tfeat.copy_(self._tfeat_n)
            self._ev_copied.record(main)
            if next_points is not None:
                self._enqueue_teacher(next_points)
        self._prefetched = next_points
        self.graph_s.replay()                                # G2b
        self._pending = True
        return self.loss

    def flush(self):
        """Apply the pending update of the last step (pipelined mode); a no-op otherwise."""
        if self._pending:
            dp.sync_gradients(self.fp)
            self.fp.set_hyper(grad_scale=1.0 / dp.world_size())
            self.graph_b.replay()
            self._pending = False


    def checkpoint(self, epoch=0, metrics=None, best_metrics=None):
        """The dict tools/builder.py:132-144 saves ({'base_model', 'optimizer', 'epoch', 'metrics', 'best_metrics'}), with
        any pending pipelined update applied first (flush()), so that nothing is lost between the last step and the save;
        `optimizer` is FlatParams.state_dict() = torch.optim.AdamW's layout."""
        self.flush()
        torch.cuda.synchronize(self.dev)
        return {"base_model": self.model.state_dict(), "optimizer": self.fp.state_dict(), "epoch": epoch,
                "metrics": metrics if metrics is not None else {}, "best_metrics": best_metrics if best_metrics is not None else {}}

    def save_checkpoint(self, path, **kw):
        torch.save(self.checkpoint(**kw), path)

    def capture(self):
        """Warm up on a side stream (allocator + lazily initialised state), then capture the step.  capture() leaves the
        model, the optimizer state, every BatchNorm buffer and the RNG streams exactly as it found them: the two eager
        warm-up steps run on the all-zero static batch and are undone (snapshot before, restore after), so a loaded
        checkpoint is not perturbed and step counts / bias corrections start where the caller left them."""
        if self.pipeline:
            return self._capture_pipeline()
        tmod = getattr(getattr(self.model, "teacher", None), "__self__", None)
        if tmod is not None and hasattr(tmod, "seed_buffer"):
            tmod.seed_buffer = self._seeds[1:2]              # single graph: the teacher's seed is staged with the mask
        l0 = ops.LAUNCHES
        snap = _StateSnapshot(self.model, self.fp, self.dev)
        self._host_prologue(None)
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._body()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        snap.restore()
        l1 = ops.LAUNCHES
        self.launches_per_step = (l1 - l0) // 2
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            if dp.world_size() == 1:
                with torch.cuda.graph(self.graph):
                    self._body()
            else:
                with torch.cuda.graph(self.graph):
                    self._body_a()
                self.graph_b = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_b, pool=self.graph.pool()):
                    self._body_b()
        return self

    def run(self, points, next_points=None):
        """One training step on `points` ([B,N,3] f32: device tensor, or pinned host tensor).  Returns the (device,
        asynchronous) loss scalar of this step.  next_points (pipelined mode): the batch of the NEXT call -- its tokenizer
        and teacher forward are issued now, beside this step's student work; the next call must then pass that batch."""
        if self.pipeline:
            return self._run_pipeline(points, next_points)
        self._host_prologue(points)
        if self.graph is None:
            self._body()
        elif self.graph_b is None:
            self.graph.replay()
        else:
            self.graph.replay()
            dp.sync_gradients(self.fp)
            self.graph_b.replay()
        return self.loss


class AutoencoderStep(_InputStaging):
    """The Stage-I dVAE training step (tools/runner_autoencoder.py:130-146: temp = get_temp(n_itr); ret = model(points,
    temperature=temp, hard=False); loss_1, loss_2 = get_loss(ret, points); loss_1 + kld_weight * loss_2; backward;
    optimizer.step; zero_grad) as one replayable CUDA graph.  The two schedules are evaluated on the host exactly like
    the reference (dvae.get_temp / dvae.get_kld_weight) and staged into a 2-element device tensor the graph reads."""

    def __init__(self, model, flat_params, batch, n_points, use_graph=True, device=None, temp_cfg=None, kld_cfg=None):
        from . import dvae
        self._dvae = dvae
        self.model, self.fp = model, flat_params
        self.dev = device or flat_params.flat.device
        self.B, self.N = batch, n_points
        self.points = torch.zeros(batch, n_points, 3, dtype=torch.float32, device=self.dev)
        self.sched = torch.ones(2, dtype=torch.float32, device=self.dev)          # [temperature, kld_weight]
        self.losses = torch.zeros(3, dtype=torch.float32, device=self.dev)         # [recon, klv, total]
        # seed of the in-kernel gumbel noise (csrc/gumbel.cu), drawn on the host from torch's CPU generator every step
        self._seed = torch.zeros(1, dtype=torch.int64, device=self.dev)
        core = getattr(model, "module", model)
        if hasattr(core, "gumbel_seed"):
            core.gumbel_seed = self._seed
        self.temp_cfg = temp_cfg or dict(start=1.0, target=0.0625, ntime=100000)   # pointbert_dvae.yaml:27-30
        self.kld_cfg = kld_cfg or dict(start=0.0, target=0.1, ntime=100000)        # pointbert_dvae.yaml:33-36
        self.use_graph = use_graph
        self.graph = self.graph_b = None
        self.n_itr = 0
        self.launches_per_step = None
        if dp.world_size() > 1:
            self.fp.enable_bf16_comm()

    def _body_a(self):
        self.fp.zero_grad()
        ret = self.model(self.points, temperature=self.sched[0], hard=False)
        l1, l2 = self.model.get_loss(ret, self.points)
        loss = l1 + self.sched[1] * l2
        loss.backward()
        self.losses.copy_(torch.stack([l1.detach(), l2.detach(), loss.detach()]))

    def _body_b(self):
        self.fp.step()

    def _body(self):
        self._body_a()
        dp.sync_gradients(self.fp)
        self._body_b()

    def _host_prologue(self, points):
        sched = torch.tensor([self._dvae.get_temp(self.n_itr, **self.temp_cfg),
                              self._dvae.get_kld_weight(self.n_itr, **self.kld_cfg)], dtype=torch.float32)
        self.sched.copy_(sched.pin_memory(), non_blocking=True)      # fresh pinned staging per step (see PretrainStep)
        self._seed.copy_(torch.randint(0, 2 ** 62, (1,), dtype=torch.int64).pin_memory(), non_blocking=True)
        self.fp.set_hyper(grad_scale=1.0 / dp.world_size())
        self._copy_in(points)

    def flush(self):
        """Serial schedule: nothing is ever pending (kept for interface parity with PretrainStep)."""

    checkpoint = PretrainStep.checkpoint
    save_checkpoint = PretrainStep.save_checkpoint

    def capture(self):
        """Same contract as PretrainStep.capture(): the warm-up steps are undone."""
        l0 = ops.LAUNCHES
        snap = _StateSnapshot(self.model, self.fp, self.dev)
        self._host_prologue(None)
        s = torch.cuda.Stream(device=self.dev)
        s.wait_stream(torch.cuda.current_stream(self.dev))
        with torch.cuda.stream(s):
            for _ in range(2):
                self._body()
        torch.cuda.current_stream(self.dev).wait_stream(s)
        torch.cuda.synchronize(self.dev)
        snap.restore()
        self.launches_per_step = (ops.LAUNCHES - l0) // 2
        if self.use_graph:
            self.graph = torch.cuda.CUDAGraph()
            if dp.world_size() == 1:
                with torch.cuda.graph(self.graph):
                    self._body()
            else:
                with torch.cuda.graph(self.graph):
                    self._body_a()
                self.graph_b = torch.cuda.CUDAGraph()
                with torch.cuda.graph(self.graph_b, pool=self.graph.pool()):
                    self._body_b()
        return self

    def run(self, points):
        """One Stage-I step on `points`; returns the device tensor [loss_recon, loss_klv, loss] of this step."""
        self._host_prologue(points)
        self.n_itr += 1
        if self.graph is None:
            self._body()
        elif self.graph_b is None:
            self.graph.replay()
        else:
            self.graph.replay()
            dp.sync_gradients(self.fp)
            self.graph_b.replay()
        return self.losses
