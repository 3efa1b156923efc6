"""Training-step driver for the engine: the reference's step semantics (train.py:139-168) without
its host synchronisations, plus data parallelism (sharded optimiser over NVLink peer memory, or NCCL all-reduce).

  step(x):  VAE.forward -> alpha*recon + beta*sum(kl) -> backward
            -> (DP) bucketed all-reduce of the gradient arena, overlapped with the rest of backward
            -> ONE fused multi-tensor launch pair: spectral-norm gradient + AdamW + global grad-norm
               (sg_opt_step; AdamW defaults of torch, train.py:92).

Gradients never take the reference layout on this path: the wgrad GEMMs write into a persistent flat
arena (engine.GradSink, laid out in backward-completion order), NCCL reduces contiguous slices of that
arena in place, and the optimiser kernel reads it directly.  The reference computes the gradient norm
with one `.item()` per parameter (126 blocking D2H copies per step, train.py:156-161) and four more
for logging (train.py:171-174); here all scalars stay on the device and are fetched only when the
caller asks (`Trainer.scalars()`).

`fused=False` keeps the per-parameter path (p.grad materialised, one AdamW launch per tensor); it is the
cross-check for the fused path in the tests and what the untouched reference train.py effectively runs
with torch.optim.AdamW.
"""
from __future__ import annotations

import os

import torch

from . import engine
from . import kernels as K


def warmup_beta(epoch: int, epochs: int, init_beta: float = 1e-4, beta_target: float = 1.0) -> float:
    """WarmupKLLoss schedule of the reference (train.py:18-41,75-81): a function of the epoch."""
    start, end = int(epochs * 0.3), int(epochs * 0.8)
    if epoch < start:
        return init_beta
    if epoch < end:
        return (epoch - start) * (beta_target - init_beta) / (end - start) + init_beta
    return beta_target


def _gemm_layout_elems(mod):
    """Elements of the wgrad output for a Conv1d / ConvTranspose1d / Linear (GEMM layout, Cin padded to 8)."""
    w = mod.weight_orig if hasattr(mod, "weight_orig") else mod.weight
    if w.dim() == 2:
        return w.numel()
    if isinstance(mod, torch.nn.ConvTranspose1d):
        cin, cout, k = w.shape
    else:
        cout, cin, k = w.shape
    return k * cout * ((cin + 7) // 8 * 8)


def shard_item(it, rank, world):
    """The part of one optimiser item (engine.GradSink.items: p, g[, u, vv, sigma, Cout, Cin, Cin_p, k, flip]) that rank
    `rank` of `world` owns under the sharded optimiser, or None when its share is empty.  Spectral-norm weights are
    split by whole rows of the native layout (Conv1d / Linear: output channels; ConvTranspose1d: input channels), which
    are also contiguous row ranges of the k tap planes of the GEMM-layout gradient, so the shard is described by views
    that simply START later - the kernels index them as if they were the whole tensor; everything else is split into
    16-byte-aligned ranges.  Returns a dict with the views, "n" (elements of the shard) and "rows" = (lo, hi)."""
    p, g = it["p"], it["g"]
    if it.get("u") is not None:
        Cout, Cin, Cin_p, k, flip = it["Cout"], it["Cin"], it["Cin_p"], it["k"], bool(it["flip"])
        rows, rowlen = (Cin, Cout * k) if flip else (Cout, Cin * k)
        lo, hi = rows * rank // world, rows * (rank + 1) // world
        if hi <= lo:
            return None
        out = dict(it)
        out.update(n=(hi - lo) * rowlen, rows=(lo, hi), full=it,
                   p=p.reshape(-1)[lo * rowlen:], g=g.reshape(-1)[(lo if flip else lo * Cin_p):],
                   u=it["u"] if flip else it["u"][lo:], vv=it["vv"][lo * k:] if flip else it["vv"])
        return out
    n = p.numel()
    blocks = (n + 3) // 4
    lo, hi = min(n, blocks * rank // world * 4), min(n, blocks * (rank + 1) // world * 4)
    if hi <= lo:
        return None
    out = dict(it)
    out.update(n=hi - lo, rows=(lo, hi), full=it, p=p.reshape(-1)[lo:], g=g.reshape(-1)[lo:])
    return out


class PeerMemory:
    """Symmetric (peer-mapped) device buffers of a data-parallel group on one NVLink / NVSwitch box: every rank allocates
    the same buffers and, after a rendezvous, holds device pointers to all of them (torch.distributed._symmetric_memory:
    CUDA VMM allocations exported between the ranks' processes)."""

    def __init__(self, group, device):
        import torch.distributed._symmetric_memory as symm_mem
        self.symm_mem, self.group, self.device = symm_mem, group, device
        self.handles = []
        self.group_name = (group if group is not None else torch.distributed.group.WORLD).group_name

    def empty(self, n, dtype=torch.float32):
        """(tensor, [device pointer of rank r's copy for every r], NVSwitch multicast address of the buffer or 0)"""
        t = self.symm_mem.empty(int(n), dtype=dtype, device=self.device)
        h = self.symm_mem.rendezvous(t, self.group_name)
        self.handles.append(h)
        ptrs = [int(x) for x in h.buffer_ptrs]
        assert ptrs[h.rank] == t.data_ptr(), "symmetric memory: own pointer mismatch"
        try:
            mc = int(h.multicast_ptr or 0)
        except Exception:
            mc = 0
        return t, ptrs, mc


class Trainer:
    def __init__(self, model, lr=1e-3, alpha=1.0e6, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.01,
                 process_group=None, bucket_mb=64, fused=True, materialize_xhat=False, loss_scale=None,
                 broadcast_init=True, growth_interval=200, single_process=False, dp_mode=None, cuda_graph=None):
        self.model = model
        self.lr, self.alpha, self.betas, self.eps, self.wd = lr, alpha, betas, eps, weight_decay
        self.params = [p for p in model.parameters() if p.requires_grad]
        self.m = {}
        self.v = {}
        self.step_count = 0
        dev = self.params[0].device
        self.dev = dev
        self.gnorm_sq = torch.zeros(1, dtype=torch.float64, device=dev)
        self.pg = process_group
        self.world = 1
        # single_process: ignore an initialised torch.distributed (a rank-local reference run next to a data-parallel one)
        if not single_process and (process_group is not None or
                                   (torch.distributed.is_available() and torch.distributed.is_initialized())):
            self.world = torch.distributed.get_world_size(process_group)
        bucket_mb = float(os.environ.get("SIMULGEN_B200_BUCKET_MB", bucket_mb))
        self.bucket_elems = int(bucket_mb * (1 << 20)) // 4
        # SMs kept free for NCCL's kernels while collectives overlap the backward pass (the persistent GEMM grids shrink
        # by that many SMs; see sg_set_sm_limit) and whether the all-reduce overlaps backward at all (0: after backward)
        self.reserve_sms = int(os.environ.get("SIMULGEN_B200_DP_RESERVE_SMS", "0"))
        self.overlap = os.environ.get("SIMULGEN_B200_DP_OVERLAP", "1") != "0"
        # data-parallel exchange: "peer" (default on CUDA) = sharded optimiser over NVLink peer memory (PeerMemory,
        # sg_peer_reduce_dot / sg_opt_step: no gradient all-reduce at all), "nccl" = bucketed ncclAllReduce of the gradient
        # arena overlapped with backward + replicated optimiser (round 1); "peer" falls back to "nccl" when symmetric
        # memory cannot be set up (and always on CPU / gloo)
        self.dp_mode = os.environ.get("SIMULGEN_B200_DP", dp_mode or "peer")
        # overlap the decoder's share of the exchange with the encoder's backward / the next forward.  Off by default:
        # measured on 2 B200s it LOSES 0.2-0.9 ms (profiles/r2_dp_peer_pipeline_2gpu.txt) - the GPU is power-bound, a
        # co-running exchange kernel slows the GEMMs by more than it hides
        self.pipeline = os.environ.get("SIMULGEN_B200_DP_PIPELINE", "0") != "0"
        self.peer = None
        self.fused = fused
        if self.world > 1 and broadcast_init:
            self._broadcast_replica()
        # fp16 mode: gradients are stored as fp16 GEMM operands, so the loss is multiplied by a power of two before
        # backward and the optimiser divides it out again (everything in between is linear).
        #   loss_scale=None (default): DYNAMIC, device-resident (kernels.make_scaler_state / sg_scaler_state): starts at
        #     numel/alpha rounded to a power of two (reconstruction-loss gradient at O(1)); a step whose gradients
        #     overflowed is skipped inside sg_opt_step and the scale halves, `growth_interval` clean steps double it -
        #     no host synchronisation.  Fused path only (the per-tensor path keeps the static initial value).
        #   loss_scale=<float>: static.            bf16 / fp32 modes: no scaling.
        self.loss_scale = loss_scale
        self.growth_interval = growth_interval
        self.scaler = None                             # created at the first fp16 step (needs numel of a batch)
        self.materialize_xhat = materialize_xhat      # the step only needs the losses; x_hat stays in registers
        self._last = None
        self.sink = None
        self.plan = None
        self._works = []
        self._launched = 0
        self._token = torch.zeros(1, dtype=torch.float32, device=dev)
        self._token2 = torch.zeros(1, dtype=torch.float32, device=dev)
        self.gnorm_dec = torch.zeros(1, dtype=torch.float64, device=dev)
        self.plan_dec, self._dots, self._v_split, self._pipelined_ok = None, None, 0, False
        self._ev_dec_done, self._ev_reduced, self._step_scaler = None, None, None
        self._peer_stream = torch.cuda.Stream(device=dev) if dev.type == "cuda" else None
        # launches that run underneath the step on the peer stream keep a small resident footprint (blocks per SM x SMs)
        self._bg_blocks = int(os.environ.get("SIMULGEN_B200_DP_BG_BLOCKS_PER_SM", "2")) * \
            (torch.cuda.get_device_properties(dev).multi_processor_count if dev.type == "cuda" else 1)
        # replay the step as one CUDA graph once it has run eagerly (single GPU, fp16 mode): opt-in, SIMULGEN_B200_GRAPH=1
        self.cuda_graph = bool(int(os.environ.get("SIMULGEN_B200_GRAPH", "0"))) if cuda_graph is None else bool(cuda_graph)
        self._graphs, self._graph_pool, self._dev_counter, self._dev_counter_host = {}, None, None, -1
        # fused path: drive the engine's tapes directly instead of going through torch.autograd (SIMULGEN_B200_DIRECT=0:
        # the autograd Functions, as the reference's train.py uses them)
        self.direct = os.environ.get("SIMULGEN_B200_DIRECT", "1") != "0"
        if fused:
            w_elems = v_elems = n_layers = 0
            wparams = set()
            for mod in model.modules():
                if isinstance(mod, (torch.nn.Conv1d, torch.nn.ConvTranspose1d, torch.nn.Linear)):
                    w = mod.weight_orig if hasattr(mod, "weight_orig") else mod.weight
                    wparams.add(id(w))
                    n_layers += 1
                    if hasattr(mod, "weight_orig"):
                        w_elems += engine.GradSink._round(_gemm_layout_elems(mod))
                    else:
                        v_elems += engine.GradSink._round(w.numel())
            for p in self.params:
                if id(p) not in wparams:
                    v_elems += engine.GradSink._round(p.numel())
            arenas = None
            if self.world > 1 and self.dp_mode == "peer" and dev.type == "cuda" and self.world <= 8:
                arenas = self._setup_peer_memory(w_elems, v_elems)
            self.sink = engine.GradSink(w_elems, v_elems, dev, arenas=arenas)
            if self.world > 1 and self.peer is None:
                self.dp_mode = "nccl"
                self.sink.on_commit = self._on_commit

    # -- data parallel --------------------------------------------------------------------------
    def _setup_peer_memory(self, w_elems, v_elems):
        """Gradient arenas and ONE flat buffer for all parameters in symmetric memory; the parameters are re-homed into
        the flat buffer (p.data becomes a view of it: same Parameter objects, same state dict).  Returns the two arenas,
        or None (with a warning) when symmetric memory is unavailable - the NCCL path takes over."""
        dist = torch.distributed
        try:
            pm = PeerMemory(self.pg, self.dev)
            weights, wptrs, wmc = pm.empty(max(w_elems, 1))
            vecs, vptrs, vmc = pm.empty(max(v_elems, 1))
            allp = list(self.model.parameters())
            offs, total = [], 0
            for p in allp:
                offs.append(total)
                total += (p.numel() + 63) // 64 * 64
            flat, pptrs, pmc = pm.empty(max(total, 1))
            ok = torch.ones(1, device=self.dev)
        except Exception as e:  # pragma: no cover - depends on the box
            ok = torch.zeros(1, device=self.dev)
            err = e
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=self.pg)
        if float(ok) == 0:
            import warnings
            warnings.warn("simulgen_b200: symmetric peer memory unavailable (%s); data parallel falls back to NCCL all-reduce"
                          % (locals().get("err", "another rank failed"),))
            return None
        weights.zero_()
        vecs.zero_()
        with torch.no_grad():
            for p, off in zip(allp, offs):
                view = flat[off:off + p.numel()].view(p.shape)
                view.copy_(p.data)
                p.data = view
        rank = dist.get_rank(self.pg)
        self.peer_mem = pm
        # NVSwitch multicast (NVLS), opt-in: gradients summed INSIDE the switch (multimem.ld_reduce), parameters written to
        # every rank with one store (multimem.st) - NVLink traffic per GPU drops from (W-1)/W to 1/W of the buffers.
        # Correct (dp_check) but measured SLOWER than plain P2P loads / stores on this box at 2 and at 8 GPUs
        # (41.9 vs 39.5 ms and 41.7 vs 40.7 ms per step, profiles/r2_dp_multicast_vs_p2p.txt), so it stays off.
        use_mc = os.environ.get("SIMULGEN_B200_DP_MULTICAST", "0") != "0" and wmc and vmc and pmc
        self.multicast = bool(use_mc)
        self.peer = K.make_peer(rank, wptrs, vptrs, pptrs, (wmc, vmc, pmc) if use_mc else None)
        self.rank = rank
        self._flat_params = flat
        torch.cuda.synchronize(self.dev)
        dist.barrier(group=self.pg)
        return weights, vecs

    def _broadcast_replica(self):
        """Data parallel needs IDENTICAL replicas: the reference's entry point builds the model on every rank without
        a seed (SimulGen-VAE.py never calls manual_seed; train.py:65-72 draws the He weights and the spectral-norm
        u / v vectors from the global RNG), so rank 0's parameters and buffers - weight_orig, bias, GroupNorm affine,
        weight_u, weight_v - are broadcast once before the first step."""
        dist = torch.distributed
        src = dist.get_global_rank(self.pg, 0) if self.pg is not None else 0
        with torch.no_grad():
            for t in list(self.model.parameters()) + list(self.model.buffers()):
                dist.broadcast(t.data, src=src, group=self.pg)

    def _on_commit(self, committed, force=False):
        """Gradient arena filled up to `committed` elements: all-reduce complete buckets right away.  The
        collective is enqueued behind the wgrad GEMMs already issued on the compute stream and runs on
        NCCL's own stream while backward continues.  force: a chunk of a large gradient - reduce it now."""
        if not self.overlap:
            return
        if committed - self._launched >= (1 if force else self.bucket_elems) and committed > self._launched:
            self._reduce(self.sink.weights[self._launched:committed])
            self._launched = committed

    def _reduce(self, flat):
        dist = torch.distributed
        # a bucket holds gradients written on the compute stream (heads, latent layers) and on the weight-gradient
        # side stream (conv wgrads): NCCL orders itself after the CURRENT stream only, so join the other one first
        engine.order_after_side_stream(flat.device)
        self._works.append(dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True))

    def _finish_reduce(self):
        if self.world == 1:
            return
        sink = self.sink
        if sink.committed > self._launched:
            self._reduce(sink.weights[self._launched:sink.committed])
        self._launched = 0
        if sink.v_used:
            self._reduce(sink.vecs[:sink.v_used])
        for w in self._works:
            w.wait()
        self._works = []

    def _allreduce_grads(self):
        """Unfused path: bucketed sum all-reduce of p.grad (mean is folded into AdamW's grad_scale)."""
        if self.world == 1:
            return
        grads = [p.grad for p in reversed(self.params) if p.grad is not None]
        bucket, size, handles = [], 0, []
        for g in grads:
            bucket.append(g)
            size += g.numel()
            if size >= self.bucket_elems:
                handles.append(self._launch_bucket(bucket))
                bucket, size = [], 0
        if bucket:
            handles.append(self._launch_bucket(bucket))
        for flat, bucket, work in handles:
            work.wait()
            off = 0
            for g in bucket:
                n = g.numel()
                g.copy_(flat[off:off + n].view_as(g))
                off += n

    def _launch_bucket(self, bucket):
        dist = torch.distributed
        if len(bucket) == 1:
            flat = bucket[0].view(-1)
        else:
            flat = torch.cat([g.reshape(-1) for g in bucket])
        work = dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=self.pg, async_op=True)
        return flat, ([] if len(bucket) == 1 else bucket), work

    # Sharded optimiser over peer memory (sg_peer_reduce_dot / sg_opt_step with sg_peer).  Backward runs with no collective
    # in flight.  The optimiser's items are split into the decoder's (complete after the decoder's backward) and the
    # encoder's, over one shared scratch buffer, and the exchange is pipelined against the step itself:
    #
    #   main stream   ... decoder bwd | encoder bwd ............ | barrier, reduce(enc), all-reduce(dots), scalars, step(enc), barrier | next fwd: encoder ... | decoder
    #   peer stream                   | barrier, reduce(dec) ...                                        | step(dec) + stores ........ barrier |  (waited for before the next decoder forward)
    #
    #   reduce(x): every rank LOADS its shard of x's gradients from all ranks' arenas (NVLink), sums, takes its share of <G, W>
    #   step(x)  : spectral-norm gradient + AdamW on the shard, updated parameters STORED to all ranks (NVLink)
    # so three quarters of the NVLink traffic (the decoder holds 75 % of the parameters) run underneath the encoder's backward
    # and the next step's encoder forward.  "barrier" = an all-reduce of one scalar (stream-ordered, no host sync).
    def before_decoder_forward(self):
        self.sync()                                   # the decoder's parameters of the previous step have landed everywhere
        if self._v_split > 0:
            self.sink.vecs[:self._v_split].zero_()

    def sync(self):
        """Make the current stream wait for the part of the last optimiser step that runs on the peer stream (data
        parallel, peer mode: the decoder's share).  Trainer.step does it itself where it matters; call it before reading
        the model outside the trainer (validation forward, state_dict(), checkpoints)."""
        if self._ev_dec_done is not None:
            torch.cuda.current_stream(self.dev).wait_event(self._ev_dec_done)

    def after_decoder_backward(self):
        dist = torch.distributed
        engine.order_after_side_stream(self.dev)
        cur = torch.cuda.current_stream(self.dev)
        ps = self._peer_stream
        ps.wait_stream(cur)
        with torch.cuda.stream(ps):
            dist.all_reduce(self._token2, group=self.pg)               # every rank's decoder gradients are complete
            K.peer_reduce_dot(self.plan_dec, self._step_scaler is not None, self.peer, clear_dots=True,
                              max_blocks=self._bg_blocks)
            self._ev_reduced = torch.cuda.Event()
            self._ev_reduced.record(ps)

    def _peer_exchange(self, b1, b2, scale, scaler, pipelined):
        dist = torch.distributed
        engine.order_after_side_stream(self.dev)
        cur = torch.cuda.current_stream(self.dev)
        ps = self._peer_stream
        if pipelined:
            cur.wait_event(self._ev_reduced)                            # reduce(dec) done: the shared dots are ours now
        dist.all_reduce(self._token, group=self.pg)                     # every rank's (encoder) gradients are complete
        if not pipelined and self.plan_dec is not None:
            K.peer_reduce_dot(self.plan_dec, scaler is not None, self.peer, clear_dots=True)
        K.peer_reduce_dot(self.plan, scaler is not None, self.peer, clear_dots=self.plan_dec is None)
        n_red = self._dots.numel() - 5                                  # every <G, W> + the overflow flag
        dist.all_reduce(self._dots[:n_red], group=self.pg)
        K.opt_step(self.plan, self.lr, b1, b2, self.eps, self.wd, self.step_count, scale, self.gnorm_sq, scaler,
                   peer=self.peer, phase=2)
        ps.wait_stream(cur)
        with torch.cuda.stream(ps):
            self.gnorm_dec.zero_()
            if self.plan_dec is not None:
                K.opt_step(self.plan_dec, self.lr, b1, b2, self.eps, self.wd, self.step_count, scale, self.gnorm_dec, scaler,
                           peer=self.peer, phase=3, max_blocks=self._bg_blocks if pipelined else 0)
            dist.all_reduce(self.gnorm_dec, group=self.pg)              # closing barrier of the decoder's share
            self._ev_dec_done = torch.cuda.Event()
            self._ev_dec_done.record(ps)
        dist.all_reduce(self.gnorm_sq, group=self.pg)                   # closing barrier of the encoder's share + norm
        if not pipelined:
            cur.wait_event(self._ev_dec_done)

    # -- optimiser ---------------------------------------------------------------------------------
    def _state(self, p):
        if p not in self.m:
            self.m[p] = torch.zeros_like(p)
            self.v[p] = torch.zeros_like(p)
        return self.m[p], self.v[p]

    def _build_plan(self):
        items = []
        vlo = self.sink.vecs.data_ptr()
        vhi = vlo + self.sink.vecs.numel() * 4
        for key in self.sink.order:
            it = dict(self.sink.items[key])
            p = it.pop("param")
            it.update(p=p.data, vec_arena=vlo <= it["g"].data_ptr() < vhi, param_id=id(p))
            if self.peer is not None:
                it = shard_item(it, self.rank, self.world)      # sharded optimiser: this rank's rows only
                if it is None:
                    continue
                n = it["n"]
                it.update(m=torch.zeros(n, dtype=torch.float32, device=self.dev),
                          v=torch.zeros(n, dtype=torch.float32, device=self.dev))
            else:
                m, v = self._state(p)
                it.update(m=m, v=v)
            items.append(it)
        self.sink.frozen = True
        if self.peer is None:
            self.plan = K.OptPlan(items, self.dev)
            return
        # two item tables over one scratch buffer: the decoder's share (first in backward order) and the encoder's
        dec_ids = set(id(p) for p in self.model.decoder.parameters()) if hasattr(self.model, "decoder") else set()
        dec = [it for it in items if it["param_id"] in dec_ids]
        enc = [it for it in items if it["param_id"] not in dec_ids]
        n_sn = sum(1 for it in items if it.get("u") is not None)
        self._dots = torch.zeros(n_sn + 6, dtype=torch.float64, device=self.dev)
        n_dec_sn = sum(1 for it in dec if it.get("u") is not None)
        self.plan_dec = K.OptPlan(dec, self.dev, dots=self._dots, dot_base=0) if dec else None
        self.plan = K.OptPlan(enc, self.dev, dots=self._dots, dot_base=n_dec_sn)
        # the decoder's vectors must form a prefix of the vector arena (they are allocated in backward order)
        sink = self.sink
        dec_v = [sink.v_slots[k] for k in sink.v_slots if k in dec_ids]
        enc_v = [sink.v_slots[k] for k in sink.v_slots if k not in dec_ids]
        self._v_split = max((off + sink._round(n) for off, n in dec_v), default=0)
        if enc_v and min(off for off, _ in enc_v) < self._v_split:
            self._v_split = 0
        self._pipelined_ok = self.plan_dec is not None and self._v_split > 0 and len(enc) > 0

    # -- CUDA graph of the whole step -----------------------------------------------------------------
    def _graph_ok(self, x):
        return (self.cuda_graph and self.fused and self.direct and not self.materialize_xhat and self.world == 1 and self.plan is not None and self.dev.type == "cuda" and
                engine.get_precision() == "fp16" and self.loss_scale is None and self.scaler is not None and
                engine._rng_state().fixed is None and torch.randn_like is engine._ORIG_RANDN_LIKE)

    def _graph_step(self, x, beta, sample_offset):
        """One optimisation step as ONE cudaGraphLaunch.  The step issues ~190 C-ABI calls and ~400 allocations from
        Python (~9 ms of host time): at the reference's batch size of 16 the GPU finishes its ~11 ms of work before the
        host has issued it.  Everything that changes from step to step already lives in device memory - loss scale and
        AdamW step count (sg_scaler_state), the Philox draw counter (sg_counter_add), the batch (a fixed input buffer:
        the loader's ring slots are captured as they are, other tensors are copied into a static buffer) - so a
        captured step can be replayed as is.  beta and the learning rate are kernel / operator arguments: a change
        (once per epoch, train.py:144,237) captures a new graph."""
        st = engine._rng_state()
        packed = isinstance(x, engine.PackedBatch)
        src = x.operand if packed else x
        key = (src.data_ptr() if packed else "copy", tuple(src.shape), src.dtype, float(beta), float(self.lr), int(sample_offset))
        ent = self._graphs.get(key)
        if ent is None:
            if len(self._graphs) >= 16:
                self._graphs.clear()                         # schedules walk through many (beta, lr) pairs: keep the cache small
            static = src if packed else torch.empty_like(src)
            if not packed:
                static.copy_(src)
            if self._dev_counter is None:
                self._dev_counter = torch.zeros(1, dtype=torch.int64, device=self.dev)
            inp = engine.PackedBatch(static, x.T) if packed else static
            g = torch.cuda.CUDAGraph()
            st.dev_counter, st.graph_draw = self._dev_counter, 0
            # inside a graph the weight-gradient side stream is not used (its fork / join would have to be captured from
            # the autograd worker thread; the replay has no launch gaps for it to fill anyway)
            overlap = engine._OVERLAP_WGRAD
            engine._OVERLAP_WGRAD = False
            try:
                with torch.cuda.graph(g, pool=self._graph_pool):
                    out = self._eager_step(inp, beta, sample_offset)
                    draws = st.graph_draw
                    K.counter_add(self._dev_counter, draws)
            finally:
                st.dev_counter = None
                engine._OVERLAP_WGRAD = overlap
            if self._graph_pool is None:
                self._graph_pool = g.pool()
            ent = self._graphs[key] = (g, static, out, draws, packed)
            self.step_count -= 1                             # the capture pass did not execute anything
        g, static, out, draws, packed = ent
        if not packed:
            static.copy_(src)
        seed = torch.initial_seed()
        if st.seed != seed:
            st.seed, st.counter = seed, 0
        if self._dev_counter_host != st.counter:             # eager steps (or a reseed) moved the draw counter meanwhile
            self._dev_counter.fill_(st.counter)
        g.replay()
        st.counter += draws                                  # host mirror of the device counter (eager steps continue from it)
        self._dev_counter_host = st.counter
        self.step_count += 1
        self._last = out
        return out

    # -- one optimisation step -------------------------------------------------------------------
    def step(self, x, beta=1e-4, sample_offset=0, packed=None):
        """packed: optional bf16 operand of x written by the batch-assembly kernel (augment.B200AugmentedLoader with
        emit_operand=True): the encoder then skips its own input packing pass."""
        if packed is None and self._graph_ok(x):
            return self._graph_step(x, beta, sample_offset)
        return self._eager_step(x, beta, sample_offset, packed)

    def _eager_step(self, x, beta=1e-4, sample_offset=0, packed=None):
        model = self.model
        engine.set_sample_offset(sample_offset)
        engine.set_packed_input(packed)
        b1, b2 = self.betas
        S = self.loss_scale
        scaler = None
        if S is None:
            S = 1.0
            if engine.get_precision() == "fp16":
                import math
                S = 2.0 ** max(-24, min(24, round(math.log2(x.numel() / max(self.alpha, 1e-30)))))
                if self.fused:
                    if self.scaler is None:
                        self.scaler = K.make_scaler_state(self.dev, S, growth_interval=self.growth_interval)
                    scaler = self.scaler
                    S = scaler[:1].view(torch.float32)      # live device scalar: loss * S without a host read
        scale = 1.0 / self.world if scaler is not None else 1.0 / (self.world * S)
        self.gnorm_sq.zero_()
        self._works, self._launched = [], 0     # an exception that escaped a previous backward must not leak buckets
        pipelined = False
        if self.fused:
            will_pipeline = (self.peer is not None and self.plan is not None and self._pipelined_ok and self.pipeline and
                             self.direct and not self.materialize_xhat)
            self.sink.begin_step(zero_vecs_from=self._v_split if will_pipeline else 0)
            if self.peer is not None and not will_pipeline:
                self.sync()
            engine.set_grad_sink(self.sink)
            engine.set_materialize_xhat(self.materialize_xhat)
            try:
                limit = self.world > 1 and self.peer is None and self.overlap and self.reserve_sms > 0 and self.dev.type == "cuda"
                if limit:
                    K.set_sm_limit(torch.cuda.get_device_properties(self.dev).multi_processor_count - self.reserve_sms)
                try:
                    if self.direct and not self.materialize_xhat:
                        # the engine's own forward + backward, no torch.autograd in between (engine.train_step_direct)
                        pipelined = self.peer is not None and self.plan is not None and self._pipelined_ok and self.pipeline
                        self._step_scaler = scaler
                        loss, recon, kl_sum, mse = engine.train_step_direct(
                            model, x, self.alpha, beta, S if (scaler is not None or S != 1.0) else None,
                            hooks=self if pipelined else None)
                    else:
                        x_hat, recon, kls, mse = model(x)
                        kl_sum = kls[0]
                        for k in kls[1:]:
                            kl_sum = kl_sum + k
                        loss = recon * self.alpha + kl_sum * beta
                        (loss * S if (scaler is not None or S != 1.0) else loss).backward()
                finally:
                    if limit:
                        K.set_sm_limit(0)
            finally:
                engine.set_grad_sink(None)
                engine.set_materialize_xhat(True)
            if self.plan is None and self.peer is None:
                self._finish_reduce()
                self._build_plan()
            elif self.plan is None:
                self._build_plan()
            elif self.peer is None:
                self._finish_reduce()
            self.step_count += 1
            if self.peer is not None:
                self._peer_exchange(b1, b2, scale, scaler, pipelined)
            else:
                K.opt_step(self.plan, self.lr, b1, b2, self.eps, self.wd, self.step_count, scale, self.gnorm_sq, scaler)
        else:
            for p in self.params:
                p.grad = None
            x_hat, recon, kls, mse = model(x)
            kl_sum = kls[0]
            for k in kls[1:]:
                kl_sum = kl_sum + k
            loss = recon * self.alpha + kl_sum * beta
            (loss * S if S != 1.0 else loss).backward()
            self._allreduce_grads()
            self.step_count += 1
            for p in self.params:
                g = p.grad
                if g is None:
                    continue
                m, v = self._state(p)
                K.adamw_step(p.data, g, m, v, self.lr, b1, b2, self.eps, self.wd, self.step_count, scale, self.gnorm_sq)
        self._last = (loss.detach(), recon.detach(), kl_sum.detach(), mse.detach())
        return self._last

    def scaler_state(self):
        """Dynamic loss scaler as a dict (scale, applied steps, skipped steps, ...), or None; synchronises."""
        return K.read_scaler_state(self.scaler) if self.scaler is not None else None

    def scalars(self):
        """(loss, recon, kl_sum, mse, grad_norm) as Python floats - the only host synchronisation."""
        loss, recon, kl, mse = self._last
        gn = self.gnorm_sq
        if self.peer is not None:                      # the decoder's share of the norm was accumulated on the peer stream
            torch.cuda.current_stream(self.dev).wait_stream(self._peer_stream)
            gn = gn + self.gnorm_dec
        return float(loss), float(recon), float(kl), float(mse), float(gn.sqrt())
