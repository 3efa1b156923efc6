"""Decoder.forward modes against the reference's own Decoder (decoder.py:170-216) with identical parameters: mode
"random", mode "fix" and mode "fix" with freeze_level >= 0 (decoder.py:202-207: the first calls cache z in self.zs, later
calls replay self.zs[i + 1]).  The noise is forced to zero on both sides.  CPU: kernels replaced by their torch models;
GPU: the real kernels in fp32 validation mode."""
import pytest
import torch

import kernel_emulator as emu
import simulgen_vae_b200 as sg
from conftest import rel_l2
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout (or its staged copy oracle/_ref) not present")


def _decoders():
    ref = ref_import.load()
    args = (32, 8, [8, 16, 24, 32], 48, 20, 4, True)
    torch.manual_seed(2)
    rd = ref.decoder.Decoder(*args)
    rd.apply(ref.common.initialize_weights_He)
    rd.apply(ref.common.add_sn)
    sg.install_overlay()
    from modules.decoder import Decoder
    from modules.common import add_sn, initialize_weights_He
    od = Decoder(*args)
    od.apply(initialize_weights_He)
    od.apply(add_sn)
    od.load_state_dict(rd.state_dict())
    return rd.eval(), od.eval()


def _run(device, monkeypatch):
    rd, od = _decoders()
    od.to(device)
    g = torch.Generator().manual_seed(4)
    monkeypatch.setattr(torch, "randn_like", lambda t, *a, **k: torch.zeros_like(t))
    calls = [(torch.randn(3, 32, generator=g), [torch.randn(3, 8, generator=g) for _ in range(3)]) for _ in range(5)]
    with torch.no_grad():
        for mode in ("random", "fix"):
            z, xs = calls[0]
            xr, klr = rd(z, xs, mode=mode)
            xo, klo = od(z.to(device), [t.to(device) for t in xs], mode=mode)
            assert rel_l2(xo, xr) < 1e-4 and len(klo) == len(klr) == 2
            for a, b in zip(klo, klr):
                assert rel_l2(a, b) < 1e-4
        freeze_level = 1
        rd.zs, od.zs = [], []
        for z, xs in calls:
            xr, _ = rd(z, xs, mode="fix", freeze_level=freeze_level)
            xo, _ = od(z.to(device), [t.to(device) for t in xs], mode="fix", freeze_level=freeze_level)
            assert len(od.zs) == len(rd.zs)
            assert rel_l2(xo, xr) < 1e-4, rel_l2(xo, xr)
        assert len(rd.zs) == freeze_level + 1
        for a, b in zip(od.zs, rd.zs):
            assert tuple(a.shape) == tuple(b.shape) and rel_l2(a, b) < 1e-4
        # freeze_level = 2: the reference replays self.zs[i + 1], which at level 1 is a level-0 latent of the wrong width:
        # it raises RuntimeError on the second call (decoder.py:179,207) - and so does the engine
        rd.zs, od.zs = [], []
        z, xs = calls[0]
        rd(z, xs, mode="fix", freeze_level=2)
        od(z.to(device), [t.to(device) for t in xs], mode="fix", freeze_level=2)
        assert len(rd.zs) == len(od.zs) == 2
        with pytest.raises(RuntimeError):
            rd(z, xs, mode="fix", freeze_level=2)
        with pytest.raises(RuntimeError):
            od(z.to(device), [t.to(device) for t in xs], mode="fix", freeze_level=2)


def test_decoder_modes_and_freeze_level_cpu(monkeypatch):
    sg.set_precision("fp32")
    try:
        with emu.install():
            _run("cpu", monkeypatch)
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)


@pytest.mark.gpu
def test_decoder_modes_and_freeze_level_gpu(monkeypatch):
    sg.set_precision("fp32")
    try:
        _run("cuda", monkeypatch)
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
