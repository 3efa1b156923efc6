"""TEST INFRASTRUCTURE ONLY - loads the *unmodified* reference modules from /root/reference, or from
the staged byte-for-byte copy oracle/_ref (oracle/make_ref.sh; that is what travels to the GPU box).

Only `oracle/make_golden*.py`, tests that pin the oracle or run the reference's own drivers on the
engine's overlay, and `bench.py --impl reference` / `cpu_baseline` may use this.  The product path
(`simulgen_vae_b200`) never imports it.

The reference's hot-path files import four non-numerical packages that are missing from this
image (`matplotlib`, `torchinfo`, `natsort`, `skimage`).  They are replaced by empty
stub modules.  The reference files use absolute imports (`from modules.common import *`,
/root/reference/modules/encoder.py:12), so they are imported under the name `modules` and then
detached from `sys.modules`, which lets the engine's own overlay package (also called `modules`)
live in the same process.
"""
import importlib
import os
import sys
import types

_STAGED = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_ref")      # written by oracle/make_ref.sh


def _find_root():
    """The reference checkout: $SIMULGEN_REFERENCE_ROOT, else /root/reference (build container), else the byte-for-byte
    staged copy oracle/_ref (what the GPU box has; oracle/make_ref.sh)."""
    env = os.environ.get("SIMULGEN_REFERENCE_ROOT")
    if env:
        return env
    for cand in ("/root/reference", _STAGED):
        if os.path.isfile(os.path.join(cand, "modules", "VAE_network.py")):
            return cand
    return "/root/reference"


REFERENCE_ROOT = _find_root()

_cache = None


def available() -> bool:
    return os.path.isfile(os.path.join(REFERENCE_ROOT, "modules", "VAE_network.py"))


def _install_stubs():
    def stub(name, **attrs):
        if name in sys.modules:
            return
        try:
            importlib.import_module(name)
            return
        except Exception:
            pass
        m = types.ModuleType(name)
        m.__dict__.update(attrs)
        sys.modules[name] = m

    stub("matplotlib")
    stub("matplotlib.pyplot")
    if "matplotlib.pyplot" in sys.modules and isinstance(sys.modules.get("matplotlib"), types.ModuleType):
        setattr(sys.modules["matplotlib"], "pyplot", sys.modules["matplotlib.pyplot"])
    stub("torchinfo", summary=lambda *a, **k: None)
    stub("natsort", natsorted=sorted)
    stub("skimage")
    stub("skimage.util", random_noise=lambda x, *a, **k: x)


def load():
    """Return a namespace with the reference's hot-path modules (VAE_network, encoder, decoder,
    common, losses).  `train` is loaded lazily by `load_train()` (needs tensorboard)."""
    global _cache
    if _cache is not None:
        return _cache
    if not available():
        raise RuntimeError("reference checkout not found at %s" % REFERENCE_ROOT)
    _install_stubs()
    saved = {k: v for k, v in sys.modules.items() if k == "modules" or k.startswith("modules.")}
    for k in saved:
        del sys.modules[k]
    saved_path = list(sys.path)
    sys.path.insert(0, REFERENCE_ROOT)
    try:
        ns = types.SimpleNamespace()
        for name in ("common", "losses", "encoder", "decoder", "VAE_network"):
            setattr(ns, name, importlib.import_module("modules." + name))
        ns.VAE = ns.VAE_network.VAE
        ns._sysmods = {k: v for k, v in sys.modules.items() if k == "modules" or k.startswith("modules.")}
    finally:
        sys.path[:] = saved_path
        for k in [k for k in sys.modules if k == "modules" or k.startswith("modules.")]:
            del sys.modules[k]
        sys.modules.update(saved)
    _cache = ns
    return ns


def build_reference_vae(cfg, seed=0, device="cpu"):
    """Build the reference VAE exactly like /root/reference/modules/train.py:65-72:
    construct, `apply(initialize_weights_He)`, `apply(add_sn)` (which draws u, v)."""
    import torch
    ref = load()
    torch.manual_seed(seed)
    m = ref.VAE(cfg["latent_dim"], cfg["hierarchical_dim"], list(cfg["enc"]), list(cfg["enc"])[::-1],
                cfg["num_node"], cfg["num_time"], lossfun=cfg.get("lossfun", "MSE"),
                batch_size=cfg.get("batch", 1), small=cfg.get("small", True))
    m.apply(ref.common.initialize_weights_He)
    m.apply(ref.common.add_sn)
    return m.to(device)


class patched_randn_like:
    """Context manager that feeds a fixed list of eps tensors to `torch.randn_like`
    (the only RNG call on the path, /root/reference/modules/decoder.py:221)."""

    def __init__(self, eps_list):
        self.eps = list(eps_list)
        self.i = 0

    def __enter__(self):
        import torch
        self._orig = torch.randn_like

        def fake(t, *a, **k):
            e = self.eps[self.i]
            self.i += 1
            assert tuple(e.shape) == tuple(t.shape), (e.shape, t.shape)
            return e.to(t.device, t.dtype)

        torch.randn_like = fake
        return self

    def __exit__(self, *exc):
        import torch
        torch.randn_like = self._orig
        return False
