#!/usr/bin/env python
"""bench.py - VAE training throughput (samples/s) of the B200 engine, 1..8 B200s.

  python bench.py --gpus N --steps K --warmup W            # the engine (this repo), BASELINE.json configs[1]
  python bench.py --config {2,3,4,5} ...                    # the other BASELINE.json configs at their real size
  python bench.py --impl reference --gpus N --steps K ...   # the reference's own CPU path on the host cores

One "step" = forward + backward + AdamW over one batch of synthetic fields.  Prints ONE JSON line (rank 0).

  value        steps with the inputs resident in HBM in the engine's input format (fp16 mode: engine.PackedBatch, the
               packed operand of the batch that the resident-dataset loader emits; bf16 mode: fp32 [B, N, T])
  e2e.value    HOST buffers: every step copies its batch from pinned host memory (H2D inside the timed region,
               double-buffered on a copy stream) and reads the loss back - PCIe-bound at this shape
  e2e.resident the reference's own mode (`load_all`, modules/utils.py:41-43): dataset resident in HBM, the host sends
               sample indices + augmentation decisions per step, sg_assemble_batch gathers / augments / packs
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

import torch  # noqa: E402

HEADLINE = dict(latent_dim=32, hierarchical_dim=8, enc=[1024, 512, 256, 128], num_node=95008, num_time=200,
                small=True, lossfun="MSE")
# BASELINE.json configs[1..4].  `batch` = default per-GPU batch, `samples` = dataset size P.
CONFIGS = {
    2: dict(HEADLINE, name="preset1 --size=small, 484x200x95008 field", batch=64, samples=484),
    3: dict(HEADLINE, small=False, name="preset1 --size=large, 484x200x95008 field", batch=64, samples=484),
    4: dict(HEADLINE, num_node=1000000, num_time=1, name="static (Dim2=1) 4096x1x1,000,000-node field", batch=512, samples=4096),
    # num_var is parsed and never read by the reference (modules/utils.py:311): the only meaning a [B, N, T] Conv1d model
    # can give 4 variables is to fold them into the node axis, N = 4 x 95008 = 380032
    5: dict(HEADLINE, num_node=380032, num_time=400, name="multi-variable num_var=4 (N = 4 x 95008 = 380032), 1024x400 field",
            batch=16, samples=1024),
}
ALPHA, BETA, LR = 1.0e6, 1.0e-4, 1.0e-3
METRIC = "VAE train samples/s @200x95008 fields"


def fwd_bwd_gflop_per_sample(cfg):
    """Closed form of BASELINE.md 2: sum 2*Cin*Cout*k*T over conv / 2*in*out over linear; backward = 2x
    forward except encoder conv0 (no dgrad) and the heads whose outputs are unused."""
    N, T, enc, small = cfg["num_node"], cfg["num_time"], cfg["enc"], cfg["small"]
    dec = enc[::-1]
    L, H = cfg["latent_dim"], cfg["hierarchical_dim"]
    fwd, bwd = 0.0, 0.0

    def conv(cin, cout, k, dgrad=True):
        nonlocal fwd, bwd
        f = 2.0 * cin * cout * k * T
        fwd += f
        bwd += f * (2 if dgrad else 1)

    def lin(i, o, live=True):
        nonlocal fwd, bwd
        f = 2.0 * i * o
        fwd += f
        if live:
            bwd += 2 * f
    w = [N] + enc
    for i in range(len(enc)):
        conv(w[i], w[i + 1], 1, dgrad=i > 0)
        if not small:
            conv(w[i + 1], w[i + 1], 3)
        conv(w[i + 1], w[i + 1], 3)
        if not small:
            conv(w[i + 1], w[i + 1], 3)
        lin(w[i + 1] * T, H, live=0 < i < len(enc) - 1)
    lin(enc[-1] * T, 2 * L)
    lin(L, L * T)
    conv(L, dec[0], 5)
    for i in range(len(dec) - 1):
        c = dec[i + 1]
        conv(dec[i], c, 3)
        if small:
            conv(c, 5 * c, 1); conv(5 * c, 5 * c, 5); conv(5 * c, c, 1)
        else:
            conv(c, c, 1); conv(c, 5 * c, 5); conv(5 * c, 5 * c, 5); conv(5 * c, c, 1)
        if i < len(dec) - 2:
            for width_in, width_out in ((c, 2 * c), (2 * c, 2 * c)):
                conv(width_in, width_in, 3)
                if not small:
                    conv(width_in, width_in, 3)
                conv(width_in, width_out, 3)
            lin(H, H * T)
            conv(H, c, 5)
    conv(dec[-1], N, 1)
    return (fwd + bwd) / 1e9


def synthetic_fields(P, N, T, device, seed, chunk=16):
    """SURVEY.md 8d generator, on the device, fp32 [P, N, T] in [-0.7, 0.7] (written chunk by chunk)."""
    g = torch.Generator(device=device).manual_seed(seed)
    a = torch.rand(N, generator=g, device=device) * 0.8 + 0.2
    phi = torch.rand(N, generator=g, device=device)
    t = torch.arange(T, dtype=torch.float32, device=device) / max(T, 1)
    out = torch.empty(P, N, T, dtype=torch.float32, device=device)
    for p0 in range(0, P, chunk):
        n = min(chunk, P - p0)
        f = torch.rand(n, generator=g, device=device) * 3.5 + 0.5
        x = 0.7 * a[None, :, None] * torch.sin(6.283185307179586 * (f[:, None, None] * t[None, None, :] + phi[None, :, None]))
        x += 0.02 * torch.randn(n, N, T, generator=g, device=device)
        out[p0:p0 + n] = x.clamp_(-0.7, 0.7)
        del x
    return out


def synthetic_batches(n_batches, B, N, T, device, seed):
    """n_batches fp32 [B, N, T] batches of the generator above (helper of the scripts under scripts/)."""
    xs = synthetic_fields(n_batches * B, N, T, device, seed).view(n_batches, B, N, T)
    return [xs[i] for i in range(n_batches)]


class ClockSampler:
    """nvidia-smi clocks + throttle reasons sampled DURING the timed region (B200_PROFILING.md)."""
    Q = ("clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.rows, self.proc, self.index = [], None, index

    def start(self):
        try:
            self.proc = subprocess.Popen(["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q,
                                          "--format=csv,noheader,nounits", "-lms", "100"],
                                         stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.th = threading.Thread(target=self._read, daemon=True)
            self.th.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.rows.append([c.strip() for c in line.split(",")])

    def stop(self):
        if self.proc is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, smax, reasons, power = [], None, set(), []
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                power.append(float(r[2]))
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "power_w": statistics.median(power) if power else None, "samples": len(sm)}


def build_engine_model(cfg, batch, device, seed=0):
    import simulgen_vae_b200 as sg
    VAE = sg.load_vae_class()
    from modules.common import add_sn, initialize_weights_He
    torch.manual_seed(seed)
    m = VAE(cfg["latent_dim"], cfg["hierarchical_dim"], list(cfg["enc"]), list(cfg["enc"])[::-1], cfg["num_node"],
            cfg["num_time"], lossfun=cfg["lossfun"], batch_size=batch, small=cfg["small"])
    m.apply(initialize_weights_He)
    m.apply(add_sn)
    return m.to(device).train(True)


# ------------------------------------------------------------------------------------------------
# CPU arm: the reference's own implementation (or its oracle port) on the host cores
# ------------------------------------------------------------------------------------------------
def cpu_reference_run(cfg, batch, steps, warmup):
    """Returns (samples/s, kind, cores, s/step).  Runs the UNMODIFIED reference modules (from /root/reference in the build
    container, from the staged byte-for-byte copy oracle/_ref on the GPU box: oracle/make_ref.sh) - kind "reference";
    only if neither exists the oracle port - kind "port"."""
    from oracle import ref_import, vae_oracle as O
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    B, N, T = batch, cfg["num_node"], cfg["num_time"]
    x = O.synthetic_field(B, N, T, seed=1234)
    g = torch.Generator().manual_seed(123)
    times = []
    if ref_import.available():
        kind = "reference"
        model = ref_import.build_reference_vae(dict(cfg, batch=B), seed=0)
        model.train(True)
        opt = torch.optim.AdamW(model.parameters(), lr=LR)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            opt.zero_grad(set_to_none=True)
            _, rl, kls, _ = model(x)
            (rl * ALPHA + sum(kls) * BETA).backward()
            opt.step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    else:
        kind = "port"
        m = build_engine_model(cfg, B, "cpu")           # parameter container only: forward is never called
        p = O.params_from_state_dict(m.state_dict())
        del m
        leaves = [v for k, v in p.items() if v.requires_grad]
        opt = torch.optim.AdamW(leaves, lr=LR)
        for it in range(warmup + steps):
            t0 = time.perf_counter()
            opt.zero_grad(set_to_none=True)
            eps = [torch.randn(s, generator=g) for s in O.eps_shapes(cfg, B)]
            _, rl, kls, _ = O.vae_forward(p, x, eps, cfg["latent_dim"], cfg["lossfun"], training=True)
            O.total_loss(rl, kls, ALPHA, BETA).backward()
            opt.step()
            if it >= warmup:
                times.append(time.perf_counter() - t0)
    return B * len(times) / sum(times), kind, cores, sum(times) / len(times)


def main():
    # everything except the final JSON line goes to stderr (the reference's modules print while moving)
    real_stdout = sys.stdout
    sys.stdout = sys.stderr
    try:
        line = _main()
    finally:
        sys.stdout = real_stdout
    if line is not None:
        print(json.dumps(line), flush=True)


def _timed_steps(step_fn, n, world, dev):
    """n calls of step_fn(i) between barriers; device time (CUDA events), max over ranks."""
    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(n):
        step_fn(i)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    return ms


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--config", type=int, default=2, choices=sorted(CONFIGS), help="BASELINE.json configs[n-1]; 2 = headline")
    ap.add_argument("--batch", type=int, default=int(os.environ.get("SIMULGEN_BENCH_BATCH", "0")), help="per-GPU batch (0: the config's default)")
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--nodes", type=int, default=0, help="override the node count (debugging)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--batch-sweep", default=os.environ.get("SIMULGEN_BENCH_SWEEP", "16,32"),
                    help="extra per-GPU batch sizes timed after the main run (config 2, 1 GPU; '' = none)")
    ap.add_argument("--precision", default=None, choices=["bf16", "fp16"],
                    help="16-bit operand format of the tensor-core path (default: the engine's default, fp16)")
    args = ap.parse_args()
    cfg = dict(CONFIGS[args.config])
    if args.nodes:
        cfg["num_node"] = args.nodes
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    gflop = fwd_bwd_gflop_per_sample(cfg)
    N, T = cfg["num_node"], cfg["num_time"]
    workload = "config %d: %s (N=%d nodes, T=%d), fwd+bwd+AdamW" % (args.config, cfg["name"], N, T)

    if args.impl == "reference":
        if rank != 0:
            return None
        steps, warm = max(1, min(args.steps, 3)), min(args.warmup, 1)
        sps, kind, cores, sec = cpu_reference_run(cfg, args.cpu_batch, steps, warm)
        sample = "%d steps of batch %d at the shape of config %d (fwd+bwd+AdamW, fp32, all host threads)" % (steps, args.cpu_batch, args.config)
        return {"impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
                "steps": steps, "warmup": warm, "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "batch": args.cpu_batch, "device": "host CPU"},
                "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": kind, "sample": sample},
                "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}

    assert torch.cuda.is_available(), "bench.py (engine arm) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import simulgen_vae_b200 as sg
    from simulgen_vae_b200 import augment, engine, kernels as K
    from simulgen_vae_b200.trainer import Trainer
    precision = args.precision or sg.DEFAULT_PRECISION
    sg.set_precision(precision)
    B = args.batch or cfg["batch"]
    # replicas: every rank draws its own weights (different seeds on purpose); Trainer broadcasts rank 0's
    model = build_engine_model(cfg, B, dev, seed=rank)
    trainer = Trainer(model, lr=LR, alpha=ALPHA)
    packed_mode = engine.loss_target(T) == "operand"

    def make_pool(batch, n_pool, seed):
        xs = synthetic_fields(n_pool * batch, N, T, dev, seed).view(n_pool, batch, N, T)
        if not packed_mode:
            return [xs[i] for i in range(n_pool)], xs[0].numel() * 4
        pool = []
        for i in range(n_pool):
            op = torch.empty(1, N, batch, sg.tp_of(T), dtype=torch.float16 if precision == "fp16" else torch.bfloat16, device=dev)
            K.pack_input(xs[i].contiguous(), op, T)
            pool.append(engine.PackedBatch(op, T))
        return pool, pool[0].operand.numel() * 2

    def run_value(batch, steps, warmup, tr):
        n_pool = 4
        pool, nbytes = make_pool(batch, n_pool, 1234 + rank)
        for i in range(warmup):
            tr.step(pool[i % n_pool], beta=BETA, sample_offset=rank * batch)
        ms = _timed_steps(lambda i: tr.step(pool[i % n_pool], beta=BETA, sample_offset=rank * batch), steps, world, dev)
        return ms, nbytes

    # ---- timed region: inputs resident in HBM -----------------------------------------------------------------------
    n_pool = 4
    pool, pool_bytes = make_pool(B, n_pool, 1234 + rank)
    for i in range(args.warmup):
        trainer.step(pool[i % n_pool], beta=BETA, sample_offset=rank * B)
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    K.PROFILE = []
    K.PROFILE_BYTES = 0
    l0 = K.LAUNCHES
    ms = _timed_steps(lambda i: trainer.step(pool[i % n_pool], beta=BETA, sample_offset=rank * B), args.steps, world, dev)
    clocks = sampler.stop() if rank == 0 else None
    launches = K.LAUNCHES - l0
    prof = K.PROFILE
    gemm_alg_bytes = K.PROFILE_BYTES
    K.PROFILE = None
    gemm_ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in prof)
    gemm_flops = sum(f for _, f, _, _ in prof) * (T / sg.tp_of(T, "bf16"))   # valid columns only
    scalars = trainer.scalars()
    scaler = trainer.scaler_state()
    value = world * B * args.steps / (ms / 1e3)

    # ---- e2e ------------------------------------------------------------------------------------------------------------
    e2e = None
    if not args.no_e2e:
        k_e2e = max(args.steps, 20)
        # (1) HOST buffers: pinned host batches (fp16 staging in fp16 mode - the loader keeps its host-side copy of the
        # dataset in the operand format, halving the PCIe bytes; fp32 otherwise), H2D of every batch inside the timed
        # region on a copy stream, double-buffered against the previous step; the loss is read back every step.
        src = synthetic_fields(2 * B, N, T, dev, 4321 + rank).view(2, B, N, T)
        host_dtype = torch.float16 if (packed_mode and precision == "fp16") else (torch.bfloat16 if packed_mode else torch.float32)
        host = [src[i].to(host_dtype).cpu().pin_memory() for i in range(2)]
        del src
        dbuf = [torch.empty(B, N, T, dtype=host_dtype, device=dev) for _ in range(2)]
        ops = [torch.empty(1, N, B, sg.tp_of(T), dtype=host_dtype, device=dev) for _ in range(2)] if packed_mode else None
        copy_stream = torch.cuda.Stream()
        done = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        losses = []

        def host_loop():
            with torch.cuda.stream(copy_stream):
                dbuf[0].copy_(host[0], non_blocking=True)
                done[0].record(copy_stream)
            for i in range(k_e2e):
                cur, nxt = i % 2, (i + 1) % 2
                if i + 1 < k_e2e:
                    with torch.cuda.stream(copy_stream):
                        if i >= 1:
                            copy_stream.wait_event(free[nxt])
                        dbuf[nxt].copy_(host[nxt], non_blocking=True)     # overlaps the step below
                        done[nxt].record(copy_stream)
                torch.cuda.current_stream().wait_event(done[cur])
                if packed_mode:
                    K.pack_input(dbuf[cur], ops[cur], T)                  # [B, N, T] -> operand layout, on the device
                    out = trainer.step(engine.PackedBatch(ops[cur], T), beta=BETA, sample_offset=rank * B)
                else:
                    out = trainer.step(dbuf[cur], beta=BETA, sample_offset=rank * B)
                free[cur].record()
                losses.append(out[0].to("cpu", non_blocking=True))        # D2H of the step's loss
        ms_host = _timed_steps(lambda i: host_loop() if i == 0 else None, 1, world, dev)
        h2d = dbuf[0].numel() * dbuf[0].element_size()
        del dbuf, ops, host
        torch.cuda.empty_cache()
        # (2) RESIDENT dataset (the reference's load_all mode): P samples fp32 in HBM, per step the host draws the
        # sampling order + augmentation decisions and sends them (a few hundred bytes); sg_assemble_batch gathers,
        # augments and emits the packed operand one batch ahead on its own stream (augment.B200AugmentedLoader).
        resident = None
        free_b, _ = torch.cuda.mem_get_info(dev)
        P = cfg["samples"]
        need = P * N * T * 4
        if need > 0.55 * free_b:
            P = max(2 * B, int(0.55 * free_b / (N * T * 4)))
        try:
            data = synthetic_fields(P, N, T, dev, 99 + rank)
            idx = torch.arange(P)
            loader = augment.B200AugmentedLoader(data, idx, B, shuffle=True, augment=True, seed=7, rank=0, world=1)
            loader.yield_packed = packed_mode
            loader.prefetch = True
            state = {"it": iter(loader), "n": 0, "h2d": 0}
            losses_r = []

            def resident_step(i):
                try:
                    batch = next(state["it"])
                except StopIteration:
                    state["it"] = iter(loader)
                    batch = next(state["it"])
                if batch.shape[0] != B:                                   # ragged last batch of an epoch: skip it
                    return resident_step(i)
                out = trainer.step(batch, beta=BETA, sample_offset=rank * B)
                losses_r.append(out[0].to("cpu", non_blocking=True))
                state["n"] += 1
            for i in range(2):
                resident_step(i)
            state["n"] = 0
            ms_res = _timed_steps(resident_step, k_e2e, world, dev)
            resident = {"value": world * B * state["n"] / (ms_res / 1e3), "unit": "samples/s", "steps": state["n"],
                        "dataset_samples_in_hbm": P, "h2d_bytes_per_step": B * (2 * 4 + 4 * 4), "d2h_bytes_per_step": 4,
                        "vs_value": (world * B * state["n"] / (ms_res / 1e3)) / value}
            del data, loader, state
        except torch.cuda.OutOfMemoryError as e:  # pragma: no cover
            resident = {"value": None, "error": "dataset does not fit next to the model: %s" % str(e)[:80]}
        e2e = {"value": world * B * k_e2e / (ms_host / 1e3), "unit": "samples/s", "h2d_bytes_per_step": h2d,
               "d2h_bytes_per_step": 4, "steps": k_e2e, "h2d_gb_per_s_per_gpu": h2d * k_e2e / (ms_host / 1e3) / 1e9,
               "note": "host-buffer mode is bound by the host -> device link: pinned host batches (%s [B, N, T]), H2D "
                       "double-buffered on a copy stream, re-layout on the device, loss read back every step" % str(host_dtype).replace("torch.", ""),
               "resident": resident}

    # ---- other per-GPU batch sizes (the reference's Batch_size is 16) ----------------------------------------------------
    sweep = []
    if world == 1 and args.config == 2 and args.batch_sweep and not args.nodes:
        del pool
        torch.cuda.empty_cache()
        n_sw = max(10, args.steps)
        for bs in [int(v) for v in args.batch_sweep.split(",") if v.strip()]:
            if bs == B:
                continue
            ms_b, _ = run_value(bs, n_sw, args.warmup, trainer)
            ent = {"per_gpu_batch": bs, "value": bs * n_sw / (ms_b / 1e3), "ms_per_step": ms_b / n_sw}
            # the same steps replayed as CUDA graphs (Trainer(cuda_graph=True)): at small batches the ~9 ms of host work
            # per step (190 C-ABI calls from Python) is what bounds the eager rate, not the GPU
            try:
                trainer.cuda_graph = True
                ms_g, _ = run_value(bs, n_sw, args.warmup + 4, trainer)
                ent["cuda_graph"] = {"value": bs * n_sw / (ms_g / 1e3), "ms_per_step": ms_g / n_sw}
            except Exception as e:  # pragma: no cover
                ent["cuda_graph"] = {"value": None, "error": str(e)[:120]}
            finally:
                trainer.cuda_graph = False
                trainer._graphs.clear()
                torch.cuda.empty_cache()
            sweep.append(ent)

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return None
    peaks = {}
    try:
        peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
    except Exception:
        pass
    peak = peaks.get("bf16_tflops_sustained", 1400.0)
    peak_src = "MEASURED_PEAKS.json bf16_tflops_sustained" if peaks else "fallback (B200_PROFILING.md sustained ~1.4 PF)"
    hbm_peak = peaks.get("hbm_gbs", 6500.0)
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None          # DRAM bytes per GEMM launch from the committed ncu capture of this very command line
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r2_gemm_traffic_b64.json")))
        if tr.get("per_gpu_batch") == B and args.config == 2 and tr.get("precision") == precision:
            traffic = tr["gemm_dram_bytes_per_launch"]
    except Exception:
        pass
    for s_ in sweep:
        s_["step_tensor_frac"] = s_["value"] * gflop * 1e9 / (peak * 1e12)
        if s_.get("cuda_graph", {}).get("value"):
            s_["cuda_graph"]["step_tensor_frac"] = s_["cuda_graph"]["value"] * gflop * 1e9 / (peak * 1e12)
    if args.config == 4:
        # static fields: the step streams the 2.29 G parameters (spectral-norm preparation 14 B, forward + dgrad reads of the
        # 16-bit copy 4 B, weight-gradient write 4 B, optimiser 36 B per element) and is bound by HBM, not by the tensor pipe
        # plus the node-axis streams of the batch, counted on VALID elements only (no layout padding): x 4 B, its 16-bit
        # operand written once and read by conv0 fprop / wgrad and by both head passes (5 x 2 B), the recon conv's output
        # written once and read by the statistics, forward and backward passes (4 x 2 B), its gradient written once and read
        # by dgrad and wgrad (3 x 2 B) = 28 B per (sample, node)
        n_par = sum(p.numel() for p in model.parameters())
        step_bytes = 58.0 * n_par + 28.0 * B * N * T
        gbs = step_bytes / (ms / args.steps / 1e3) / 1e9
        roofline = {"bound": "hbm", "kernel": "whole step (weight / optimiser streams dominate at T = 1)", "achieved": gbs,
                    "peak": hbm_peak, "unit": "GB/s", "frac": gbs / hbm_peak, "traffic": None,
                    "algorithmic_bytes_per_step": step_bytes, "peak_source": "MEASURED_PEAKS.json hbm_gbs" if peaks else "fallback",
                    "gemm_tflops": achieved, "gemm_share_of_step": gemm_ms / ms if ms > 0 else None}
    else:
        roofline = {"bound": "tensor", "kernel": "conv_gemm_tc2_kernel (tcgen05 cta_group::2 implicit-GEMM fprop/dgrad/wgrad)",
                    "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                    "algorithmic_flop_per_launch": gemm_flops / max(len(prof), 1),
                    "algorithmic_bytes_per_launch": gemm_alg_bytes / max(len(prof), 1),
                    "traffic_source": "profiles/r2_gemm_traffic_b64.json: ncu dram__bytes_read.sum + dram__bytes_write.sum per GEMM launch of "
                                      "`python bench.py --steps 2 --warmup 3 --no-e2e --no-cpu-baseline --batch-sweep ''` (config 2, batch 64, fp16); "
                                      "algorithmic_bytes_per_launch = every operand plane and the output moved once",
                    "peak_source": peak_src, "launches_timed": len(prof),
                    "share_of_step": gemm_ms / ms if ms > 0 else None}
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": precision,
        "data": "synthetic",
        "config": {"workload": workload, "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2_policy": "inputs larger than L2 (%.0f MB per batch, 4 batches cycled)" % (pool_bytes / 1e6),
                   "gflop_per_sample_fwd_bwd": gflop, "loss": scalars[0], "grad_norm": scalars[4],
                   "resident_input_format": "engine.PackedBatch: fp16 operand [N, B, T] written by the loader kernel" if packed_mode else "fp32 [B, N, T]",
                   "loss_scaler": scaler,
                   "dp_exchange": (None if world == 1 else
                                   ("peer memory: sharded optimiser, %s" % ("NVSwitch multicast (multimem.ld_reduce / multimem.st)"
                                                                            if getattr(trainer, "multicast", False) else "P2P loads / stores")
                                    if trainer.peer is not None else "nccl all-reduce overlapped with backward")),
                   "batch_sweep": sweep,
                   "notes": "operands of the tensor-core path are %s with fp32 accumulation (tcgen05 kind::f16); x_hat is not written to "
                            "HBM by the training step (train.py:142 discards it); pre-norm conv outputs and interior activation "
                            "gradients are stored in the 16-bit operand format in fp16 mode; the loss target is the packed operand of "
                            "x in fp16 mode (DESIGN.md section 3); gpu_launches counts C-ABI calls (each >= 1 kernel)" % precision},
        "clocks": clocks,
        "gpu_launches": launches,
        "step_tensor_frac": value / world * gflop * 1e9 / (peak * 1e12),
        "roofline": roofline,
    }
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:           # the CPU baseline is timed on rank 0 of the 1-GPU run only
        try:
            sps, kind, cores, sec = cpu_reference_run(cfg if args.config in (2, 3) else CONFIGS[2], args.cpu_batch, 2, 1)
            line["cpu_baseline"] = {"value": sps, "unit": "samples/s", "cores": cores, "kind": kind,
                                    "sample": "2 steps of batch %d at the headline shape (fwd+bwd+AdamW, fp32)" % args.cpu_batch}
        except Exception as e:  # pragma: no cover
            line["cpu_baseline"] = {"value": None, "unit": "samples/s", "cores": os.cpu_count(), "kind": "port",
                                    "sample": "failed: %s" % e}
    if world > 1:
        torch.distributed.destroy_process_group()
    return line


if __name__ == "__main__":
    main()
