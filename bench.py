#!/usr/bin/env python
"""bench.py - VAE training throughput (samples/s) on the headline shape, 1..8 B200s.

  python bench.py --gpus N --steps K --warmup W            # the engine (this repo)
  python bench.py --impl reference --gpus N --steps K ...   # the reference's CPU path on the host cores

One "step" = forward + backward + AdamW over one batch of synthetic [B, 95008, 200] fields
(BASELINE.json configs[1]: preset 1, --size=small, bf16).  Prints ONE JSON line (rank 0).
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


def synthetic_batches(n_batches, B, N, T, device, seed):
    """SURVEY.md 8d generator, on the device, fp32 [B, N, T] in [-0.7, 0.7]."""
    g = torch.Generator(device=device).manual_seed(seed)
    a = torch.rand(N, generator=g, device=device) * 0.8 + 0.2
    phi = torch.rand(N, generator=g, device=device)
    t = torch.arange(T, dtype=torch.float32, device=device) / T
    out = []
    for _ in range(n_batches):
        f = torch.rand(B, generator=g, device=device) * 3.5 + 0.5
        x = 0.7 * a[None, :, None] * torch.sin(6.283185307179586 * (f[:, None, None] * t[None, None, :] + phi[None, :, None]))
        x += 0.02 * torch.randn(B, N, T, generator=g, device=device)
        out.append(x.clamp_(-0.7, 0.7).contiguous())
    return out


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
        sm, smax, reasons = [], None, set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for r in self.rows:
            try:
                sm.append(float(r[0]))
                smax = float(r[1])
                for n, v in zip(names, r[3:7]):
                    if v.lower().startswith("active"):
                        reasons.add(n)
            except Exception:
                pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": smax, "reasons": sorted(reasons),
                "samples": len(sm)}


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
    """Returns (samples/s, kind, cores).  Uses the unmodified reference modules when /root/reference is
    present (build container), else the oracle port (GPU box)."""
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


def _main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--batch", type=int, default=int(os.environ.get("SIMULGEN_BENCH_BATCH", "64")), help="per-GPU batch")
    ap.add_argument("--impl", default="engine", choices=["engine", "reference"])
    ap.add_argument("--nodes", type=int, default=HEADLINE["num_node"])
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--cpu-batch", type=int, default=2)
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp16"],
                    help="16-bit operand format of the tensor-core path (fp16: same kernels, loss scaling, ~8x smaller rounding error)")
    args = ap.parse_args()
    cfg = dict(HEADLINE, num_node=args.nodes)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    gflop = fwd_bwd_gflop_per_sample(cfg)
    workload = "preset1 --size=small, %dx%d fields (N=%d nodes, T=%d), fwd+bwd+AdamW" % (
        cfg["num_time"], cfg["num_node"], cfg["num_node"], cfg["num_time"])

    if args.impl == "reference":
        if rank != 0:
            return None
        sps, kind, cores, sec = cpu_reference_run(cfg, args.cpu_batch, max(1, min(args.steps, 3)), min(args.warmup, 1))
        line = {"impl": "reference", "metric": METRIC, "value": sps, "unit": "samples/s", "n_gpus": args.gpus,
                "steps": max(1, min(args.steps, 3)), "warmup": min(args.warmup, 1), "ms_per_step": sec * 1e3,
                "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
                "config": {"workload": workload, "batch": args.cpu_batch, "device": "host CPU"},
                "cpu_baseline": {"value": sps, "unit": "samples/s", "cores": cores, "kind": kind,
                                 "sample": "%d steps of batch %d at the headline shape" % (max(1, min(args.steps, 3)), args.cpu_batch)},
                "e2e": {"value": sps, "unit": "samples/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
                "gpu_launches": 0}
        return line

    assert torch.cuda.is_available(), "bench.py (engine arm) needs a CUDA device: there is no CPU fallback"
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    import simulgen_vae_b200 as sg
    from simulgen_vae_b200 import kernels as K
    from simulgen_vae_b200.trainer import Trainer
    sg.set_precision(args.precision)
    B = args.batch
    model = build_engine_model(cfg, B, dev, seed=0)          # same seed on every rank = identical replicas
    trainer = Trainer(model, lr=LR, alpha=ALPHA, process_group=pg)
    n_pool = 4
    pool = synthetic_batches(n_pool, B, cfg["num_node"], cfg["num_time"], dev, seed=1234 + rank)

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # ---- warm-up -----------------------------------------------------------------------------------
    for i in range(args.warmup):
        trainer.step(pool[i % n_pool], beta=BETA, sample_offset=rank * B)
    barrier()
    # ---- timed region: inputs resident in HBM ----------------------------------------------------------
    sampler = ClockSampler(local_rank)
    if rank == 0:
        sampler.start()
    K.PROFILE = []
    l0 = K.LAUNCHES
    ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    barrier()
    ev0.record()
    for i in range(args.steps):
        trainer.step(pool[i % n_pool], beta=BETA, sample_offset=rank * B)
    ev1.record()
    barrier()
    ms = ev0.elapsed_time(ev1)
    clocks = sampler.stop() if rank == 0 else None
    launches = K.LAUNCHES - l0
    prof = K.PROFILE
    K.PROFILE = None
    gemm_ms = sum(e0.elapsed_time(e1) for _, _, e0, e1 in prof)
    gemm_flops = sum(f for _, f, _, _ in prof) * (cfg["num_time"] / sg.tp_of(cfg["num_time"], "bf16"))   # valid columns only
    scalars = trainer.scalars()
    if world > 1:
        t = torch.tensor([ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
        ms = float(t)
    value = world * B * args.steps / (ms / 1e3)

    # ---- e2e: host buffers, H2D of every batch + D2H of the loss inside the timed region -------------------
    e2e = None
    if not args.no_e2e:
        host = [p.cpu().pin_memory() for p in pool[:2]]
        dbuf = [torch.empty_like(pool[0]) for _ in range(2)]
        copy_stream = torch.cuda.Stream()
        done = [torch.cuda.Event() for _ in range(2)]
        free = [torch.cuda.Event() for _ in range(2)]
        k_e2e = max(2, min(args.steps, 4))
        barrier()
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        t0.record()
        with torch.cuda.stream(copy_stream):
            dbuf[0].copy_(host[0], non_blocking=True)
            done[0].record(copy_stream)
        losses = []
        for i in range(k_e2e):
            cur, nxt = i % 2, (i + 1) % 2
            if i + 1 < k_e2e:
                with torch.cuda.stream(copy_stream):
                    if i >= 1:
                        copy_stream.wait_event(free[nxt])
                    dbuf[nxt].copy_(host[nxt], non_blocking=True)     # overlaps the step below
                    done[nxt].record(copy_stream)
            torch.cuda.current_stream().wait_event(done[cur])
            out = trainer.step(dbuf[cur], beta=BETA, sample_offset=rank * B)
            free[cur].record()
            losses.append(out[0].to("cpu", non_blocking=True))        # D2H of the step's loss
        t1.record()
        barrier()
        ms_e2e = t0.elapsed_time(t1)
        if world > 1:
            t = torch.tensor([ms_e2e], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(t, op=torch.distributed.ReduceOp.MAX)
            ms_e2e = float(t)
        e2e = {"value": world * B * k_e2e / (ms_e2e / 1e3), "unit": "samples/s",
               "h2d_bytes_per_step": pool[0].numel() * 4, "d2h_bytes_per_step": 4, "steps": k_e2e,
               "note": "pinned host fp32 batches, H2D double-buffered on a copy stream, loss read back every step"}

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
    achieved = gemm_flops / (gemm_ms / 1e3) / 1e12 if gemm_ms > 0 else 0.0
    traffic = None          # DRAM bytes per GEMM launch from the committed ncu capture of this very command line
    try:
        tr = json.load(open(os.path.join(ROOT, "profiles", "r1_gemm_traffic_b64.json")))
        if tr.get("per_gpu_batch") == B and cfg["num_node"] == HEADLINE["num_node"]:
            traffic = tr["gemm_dram_bytes_per_launch"]
    except Exception:
        pass
    line = {
        "metric": METRIC, "value": value, "unit": "samples/s", "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": args.precision,
        "data": "synthetic",
        "config": {"workload": workload, "per_gpu_batch": B, "global_batch": B * world, "parallelism": "dp%d" % world,
                   "l2_policy": "inputs larger than L2 (%.0f MB per batch, 4 batches cycled)" % (pool[0].numel() * 4 / 1e6),
                   "gflop_per_sample_fwd_bwd": gflop, "loss": scalars[0], "grad_norm": scalars[4],
                   "notes": "x_hat is not written to HBM by the training step (train.py:142 discards it); the recon layer's "
                            "pre-norm output is stored in the 16-bit operand format; gpu_launches counts C-ABI calls (each >= 1 kernel)"},
        "clocks": clocks,
        "gpu_launches": launches,
        "step_tensor_frac": value / world * gflop * 1e9 / (peak * 1e12),
        "roofline": {"bound": "tensor", "kernel": "conv_gemm_tc_kernel (tcgen05 implicit-GEMM fprop/dgrad/wgrad)",
                     "achieved": achieved, "peak": peak, "unit": "TFLOP/s", "frac": achieved / peak, "traffic": traffic,
                     "algorithmic_flop_per_launch": gemm_flops / max(len(prof), 1),
                     "peak_source": peak_src, "launches_timed": len(prof),
                     "share_of_step": gemm_ms / ms if ms > 0 else None},
    }
    if e2e:
        line["e2e"] = e2e
    if not args.no_cpu_baseline and world == 1:           # the CPU baseline is timed on rank 0 of the 1-GPU run only
        try:
            sps, kind, cores, sec = cpu_reference_run(cfg, args.cpu_batch, 2, 1)
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
