"""Stand-alone timing of the HBM-streaming kernels at the headline shapes (per-GPU batch B): CUDA-event time and
achieved algorithmic GB/s against the measured HBM copy bandwidth.  Small memory footprint so that it can also
run under `ncu --set full`."""
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
from simulgen_vae_b200 import kernels as K  # noqa: E402
from simulgen_vae_b200.engine import tp_of  # noqa: E402

B = int(sys.argv[1]) if len(sys.argv) > 1 else 32
REPS = int(sys.argv[2]) if len(sys.argv) > 2 else 5
ONLY = sys.argv[3] if len(sys.argv) > 3 else ""
T, N, G = 200, 95008, 8
Tp = tp_of(T, "bf16")
dev = torch.device("cuda")
BF = torch.bfloat16
try:
    PEAK = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))["hbm_gbs"]
except Exception:
    PEAK = 6650.0


def timed(fn):
    fn()
    torch.cuda.synchronize()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    ts = []
    for _ in range(REPS):
        flush.zero_()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        ts.append(e0.elapsed_time(e1))
    return sorted(ts)[len(ts) // 2]


def report(name, ms, nbytes):
    gbs = nbytes / (ms * 1e-3) / 1e9
    print("%-34s B=%d %8.3f ms  %7.2f GB  %7.0f GB/s  %.3f of measured HBM peak (%.0f)" % (name, B, ms, nbytes / 1e9, gbs, gbs / PEAK, PEAK),
          flush=True)


def want(name):
    return not ONLY or ONLY in name


if want("assemble"):
    Pn = 6
    data = torch.rand(Pn, N, T, device=dev) * 1.4 - 0.7
    g = torch.Generator().manual_seed(0)
    idx = torch.randint(0, Pn, (B,), generator=g)
    oth = torch.where(torch.rand(B, generator=g) < 0.5, torch.randint(0, Pn, (B,), generator=g), torch.full((B,), -1))
    ids = torch.stack([idx, oth]).to(torch.int32).to(dev)
    nl = torch.where(torch.rand(B, generator=g) < 0.5, torch.full((B,), 0.05), torch.zeros(B))
    lam = torch.rand(B, generator=g) * 0.8 + 0.1
    table = torch.stack([nl, torch.rand(B, generator=g) * 0.2 + 0.9, lam, 1 - lam]).float().to(dev)
    outb = torch.empty(B, N, T, device=dev)
    opb = torch.empty(1, N, B, Tp, device=dev, dtype=BF)
    nmix = int((oth >= 0).sum())
    per = N * T * 4
    report("assemble_batch (aug, fp32 out)", timed(lambda: K.assemble_batch(data, ids, table, None, outb, 1, 0)),
           (2 * B + nmix) * per)
    report("assemble_batch (+ bf16 operand)", timed(lambda: K.assemble_batch(data, ids, table, None, outb, 1, 0, opb)),
           (2 * B + nmix) * per + opb.numel() * 2)
    del data, outb, opb
if want("recon") or want("pack"):
    x = torch.rand(B, N, T, device=dev) * 1.4 - 0.7
if want("pack"):
    op = torch.empty(1, N, B, Tp, device=dev, dtype=BF)
    report("pack_input", timed(lambda: K.pack_input(x, op, T)), x.numel() * 4 + op.numel() * 2)
    del op
if want("recon"):
    y = torch.randn(N, B, Tp, device=dev)
    y[..., T:] = 0
    gamma, beta = torch.ones(N, device=dev), torch.zeros(N, device=dev)
    stats = torch.empty(B, G, 2, device=dev)
    report("gn_stats (recon y)", timed(lambda: K.gn_stats(y, stats, T, G)), y.numel() * 4)
    y = y.to(BF)                                   # bf16 mode stores the recon layer's pre-norm output as bf16
    ybytes = 2
    x_hat = torch.empty(B, N, T, device=dev)
    sums = torch.empty(2, device=dev, dtype=torch.float64)
    rows = torch.empty(N * B, 4, device=dev)
    report("recon_fwd (x_hat + rowsums)", timed(lambda: K.recon_fwd(y, stats, gamma, beta, x, x_hat, sums, T, G, 0, rows)),
           y.numel() * ybytes + x.numel() * 8 + rows.numel() * 4)
    report("recon_fwd (no x_hat)", timed(lambda: K.recon_fwd(y, stats, gamma, beta, x, None, sums, T, G, 0, rows)),
           y.numel() * ybytes + x.numel() * 4 + rows.numel() * 4)
    del x_hat
    dy = torch.empty(1, N, B, Tp, device=dev, dtype=BF)
    dg, db, dbi = (torch.empty(N, device=dev) for _ in range(3))
    gl, gm = torch.tensor([1e6], device=dev), torch.tensor([0.0], device=dev)
    inv = 1.0 / (B * N * T)
    report("recon_bwd (one pass)", timed(lambda: K.recon_bwd(y, stats, gamma, beta, x, gl, gm, inv, None, dy, dg, db, dbi, T, G, 0, rows)),
           y.numel() * ybytes + x.numel() * 4 + dy.numel() * 2 + rows.numel() * 4)
    # round 2: the target read from the packed 16-bit operand of x ([N, B, Tp], the layout of y; engine.loss_target)
    opx = torch.empty(1, N, B, Tp, device=dev, dtype=BF)
    K.pack_input(x, opx, T)
    report("recon_fwd (packed 16-bit target)", timed(lambda: K.recon_fwd(y, stats, gamma, beta, opx[0], None, sums, T, G, 0, rows)),
           y.numel() * ybytes + opx.numel() * 2 + rows.numel() * 4)
    report("recon_bwd (packed 16-bit target)", timed(lambda: K.recon_bwd(y, stats, gamma, beta, opx[0], gl, gm, inv, None, dy, dg, db, dbi, T, G, 0, rows)),
           y.numel() * ybytes + opx.numel() * 2 + dy.numel() * 2 + rows.numel() * 4)
    del opx
    del y, dy, rows, x
if want("gn_act"):
    # (channels, operand planes, residual, 16-bit y, 16-bit incoming gradient): round-1 fp32 hand-offs and the
    # round-2 16-bit ones (engine.store16) side by side.  Bytes = algorithmic minimum of the kernel as called.
    for C, P, res, y16, d16 in ((5120, 5, False, False, False), (5120, 5, False, True, True), (5120, 1, False, False, False),
                                (5120, 1, False, True, True), (1024, 1, True, False, False), (1024, 1, True, True, False)):
        OPD = torch.float16
        tag = "C=%d planes=%d res=%d y%d d%d" % (C, P, res, 16 if y16 else 32, 16 if d16 else 32)
        y = torch.randn(C, B, Tp, device=dev)
        y[..., T:] = 0
        gamma, beta = torch.ones(C, device=dev), torch.zeros(C, device=dev)
        stats = torch.empty(B, G, 2, device=dev)
        if not y16:
            report("gn_stats C=%d" % C, timed(lambda: K.gn_stats(y, stats, T, G)), y.numel() * 4)
        else:
            K.gn_stats(y, stats, T, G)
            y = y.to(OPD)
        yb, db_ = (2 if y16 else 4), (2 if d16 else 4)
        out = torch.empty(P, C, B, Tp, device=dev, dtype=OPD)
        of = torch.empty(C, B, Tp, device=dev) if res else None
        r = torch.randn(C, B, Tp, device=dev) if res else None
        nb = y.numel() * yb + out.numel() * 2 + (y.numel() * 8 if res else 0)
        report("gn_act_fwd " + tag,
               timed(lambda: K.gn_act_fwd(y, stats, gamma, beta, r, 0.1 if res else 1.0, K.ACT_GELU, False, out, of, T, G)), nb)
        dout = torch.randn(C, B, Tp, device=dev).to(OPD if d16 else torch.float32)
        dyo = torch.empty(P, C, B, Tp, device=dev, dtype=OPD)
        dg, db, dbi = (torch.empty(C, device=dev) for _ in range(3))
        dres = torch.zeros(C, B, Tp, device=dev) if res else None
        # pass 1: y + dout read, dz written (+ dres read-modify-write); pass 2: y + dz read, dy planes written
        nb = y.numel() * (2 * yb + 3 * db_) + dyo.numel() * 2 + (y.numel() * 8 if res else 0)
        report("gn_act_bwd " + tag,
               timed(lambda: K.gn_act_bwd(y, stats, gamma, beta, r, 0.1 if res else 1.0, K.ACT_GELU, False, dout, dyo, dg, db, dbi,
                                          dres, 1, T, G)), nb)
        del y, out, dout, dyo, of, r, dres
if want("opt"):
    items = []
    for Cout, Cin, k in ((1024, 95008, 1), (5120, 5120, 5)):
        Cin_p = (Cin + 7) // 8 * 8
        p = torch.randn(Cout, Cin, k, device=dev) * 0.01
        items.append(dict(p=p, g=torch.randn(k, Cout, Cin_p, device=dev), m=torch.zeros_like(p), v=torch.zeros_like(p),
                          u=torch.randn(Cout, device=dev), vv=torch.randn(Cin * k, device=dev), sigma=torch.ones(1, device=dev),
                          Cout=Cout, Cin=Cin, Cin_p=Cin_p, k=k, flip=0))
    gn = torch.zeros(1, device=dev, dtype=torch.float64)
    for i, it in enumerate(items):
        plan = K.OptPlan([it], dev)
        n = it["p"].numel()
        report("opt_step k=%d (%dM elems)" % (it["k"], n // 1000000),
               timed(lambda: K.opt_step(plan, 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 1.0, gn)), n * 36)
if want("sn_prepare"):
    layers = []
    for Cout, Cin, k in ((1024, 95008, 1), (5120, 5120, 5), (95008, 1024, 1)):
        Cin_p = (Cin + 7) // 8 * 8
        w = torch.randn(Cout, Cin, k, device=dev) * 0.01
        layers.append(dict(w=w, u=torch.randn(Cout, device=dev), v=torch.randn(Cin * k, device=dev), sigma=torch.ones(1, device=dev),
                           wg=torch.empty(k, Cout, Cin_p, device=dev, dtype=BF), H=Cout, Cin=Cin, k=k, Cin_p=Cin_p, so=Cin * k, si=k,
                           flip=0))
    plan = K.SnPlan(layers, dev, BF)
    n = sum(L["w"].numel() for L in layers)
    report("sn_prepare (3 big layers, %dM)" % (n // 1000000), timed(lambda: K.sn_prepare(plan, True)), n * 14)
if want("minmax"):
    # the preprocessing scan over B parameter sets of the headline field, float64 like the reference
    import time
    import numpy as np
    R = B * T
    xd = torch.rand(R, N, device=dev, dtype=torch.float64) * 40 - 20
    mn, mx = torch.empty(N, dtype=torch.float64, device=dev), torch.empty(N, dtype=torch.float64, device=dev)
    report("minmax_fit (f64, all rows)", timed(lambda: K.minmax_fit(xd, None, mn, mx)), R * N * 8)
    rows = torch.sort(torch.randperm(R, device=dev)[:R // 10]).values
    report("minmax_fit (f64, 10% rows)", timed(lambda: K.minmax_fit(xd, rows, mn, mx)), (R // 10) * N * 8)
    K.minmax_fit(xd, None, mn, mx)
    scale, minv = 1.4 / (mx - mn), -0.7 - mn * (1.4 / (mx - mn))
    out_t = torch.empty(B, N, T, dtype=torch.float32, device=dev)
    report("minmax_transform (f64 in place)", timed(lambda: K.minmax_transform(xd, scale, minv, out=xd)), 2 * R * N * 8)
    report("minmax_transform (-> f32 [P,N,T])", timed(lambda: K.minmax_transform(xd, scale, minv, out=None, out_t=out_t, T=T)),
           R * N * 12)
    report("minmax_transform (both)", timed(lambda: K.minmax_transform(xd, scale, minv, out=xd, out_t=out_t, T=T)), R * N * 20)
    # the reference's CPU path on a bounded sample (4 parameter sets): MinMaxScaler.fit + transform + transpose/cast
    from sklearn.preprocessing import MinMaxScaler
    xs = xd[:4 * T].cpu().numpy()
    t0 = time.perf_counter()
    sc = MinMaxScaler(feature_range=(-0.7, 0.7)).fit(xs)
    t1 = time.perf_counter()
    ys = sc.transform(xs)
    x32 = np.float32(ys.reshape(4, T, N).transpose((0, 2, 1)))
    t2 = time.perf_counter()
    print("cpu (sklearn/numpy, %d cores, 4 parameter sets): fit %.1f GB/s, transform+transpose+cast %.1f GB/s" %
          (os.cpu_count(), xs.nbytes / (t1 - t0) / 1e9, (xs.nbytes * 2.5) / (t2 - t1) / 1e9), flush=True)
    del xd, out_t
