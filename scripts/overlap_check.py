"""Single-GPU check that running the weight-gradient GEMMs on the side stream does not change the result:
2 Trainer steps with the overlap off and on, fp32 validation mode (and bf16), weights compared."""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch  # noqa: E402
import bench  # noqa: E402
import simulgen_vae_b200 as sg  # noqa: E402
from simulgen_vae_b200 import engine  # noqa: E402
from simulgen_vae_b200.trainer import Trainer  # noqa: E402

dev = torch.device("cuda")
cfg = dict(bench.HEADLINE, num_node=2048, enc=[256, 128, 64, 32])
B = 4
data = bench.synthetic_batches(2, B, cfg["num_node"], cfg["num_time"], dev, seed=7)


def run(overlap):
    engine._OVERLAP_WGRAD = overlap
    engine._OVERLAP_MAX_FLOP = None
    torch.manual_seed(11)
    engine._rng_state().seed = None
    m = bench.build_engine_model(cfg, B, dev, seed=0)
    tr = Trainer(m, lr=1e-3, alpha=1e6)
    for x in data:
        tr.step(x, beta=1e-4)
    torch.cuda.synchronize()
    return {k: v.double().clone() for k, v in m.state_dict().items()}


ok = True
for precision in sys.argv[1:] or ["fp32", "bf16"]:
    sg.set_precision(precision)
    a, a2, b = run(False), run(False), run(True)
    for tag, x, y in (("off vs off", a, a2), ("off vs on ", a, b)):
        per = sorted(((float((x[k] - y[k]).norm() / (x[k].norm() + 1e-30)), k) for k in x), reverse=True)
        print("overlap_check[%s] %s worst %.3e %s" % (precision, tag, per[0][0], per[0][1]), flush=True)
        for e, k in per[1:4]:
            print("     %.3e %s" % (e, k))
