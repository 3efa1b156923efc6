"""Standalone calls of the overlay's block classes (simulgen_vae_b200.blocks.run_block) against the reference's own
block modules (common.py:78-162, encoder.py:14-94, decoder.py:17-82) with the same parameters: output, input gradient and
parameter gradients.  CPU: kernels replaced by their torch models (host wiring); GPU: the real kernels, all precisions."""
import pytest
import torch

import kernel_emulator as emu
import simulgen_vae_b200 as sg
from conftest import rel_l2
from oracle import ref_import

pytestmark = pytest.mark.skipif(not ref_import.available(), reason="reference checkout (or its staged copy oracle/_ref) not present")

CASES = [
    ("ResidualBlock", "common", (16, True), 16),
    ("ResidualBlock", "common", (16, False), 16),
    ("EncoderResidualBlock", "common", (24, 24, True), 24),
    ("DecoderResidualBlock", "common", (8, True), 8),
    ("DecoderResidualBlock", "common", (8, False), 8),
    ("ConvBlock", "encoder", (40, 16, False), 40),
    ("EncoderBlock", "encoder", ([40, 16, 8], True), 40),
    ("UpsampleBlock", "decoder", (8, 16), 8),
    ("DecoderBlock", "decoder", ([8, 16, 24], True), 8),
]


def _pair(cls_name, where, args):
    """(reference block, overlay block) with identical parameters and spectral-norm state."""
    ref = ref_import.load()
    torch.manual_seed(5)
    rb = getattr(getattr(ref, where), cls_name)(*args)
    rb.apply(ref.common.initialize_weights_He)
    rb.apply(ref.common.add_sn)
    sg.install_overlay()
    import importlib
    mod = importlib.import_module("modules." + where)
    ob = getattr(mod, cls_name)(*args)
    from modules.common import add_sn, initialize_weights_He
    ob.apply(initialize_weights_He)
    ob.apply(add_sn)
    assert list(ob.state_dict().keys()) == list(rb.state_dict().keys())
    ob.load_state_dict(rb.state_dict())
    return rb.train(True), ob.train(True)


def _check(cls_name, where, args, cin, device, tol_out, tol_grad):
    rb, ob = _pair(cls_name, where, args)
    g = torch.Generator().manual_seed(9)
    x = torch.randn(3, cin, 20, generator=g)
    w = None
    xr = x.clone().requires_grad_(True)
    yr = rb(xr)
    w = torch.randn(yr.shape, generator=g)
    (yr * w).sum().backward()
    ob.to(device)
    xo = x.clone().to(device).requires_grad_(True)
    yo = ob(xo)
    assert tuple(yo.shape) == tuple(yr.shape)
    (yo * w.to(device)).sum().backward()
    assert rel_l2(yo, yr) < tol_out, rel_l2(yo, yr)
    assert rel_l2(xo.grad, xr.grad) < tol_grad, rel_l2(xo.grad, xr.grad)
    rp, op = dict(rb.named_parameters()), dict(ob.named_parameters())
    for k, p in rp.items():
        assert op[k].grad is not None, k
        assert rel_l2(op[k].grad, p.grad) < tol_grad, (k, rel_l2(op[k].grad, p.grad))
    for k, b in rb.named_buffers():                                     # the power iteration advanced identically
        assert rel_l2(dict(ob.named_buffers())[k], b) < 1e-5, k


@pytest.mark.parametrize("cls_name,where,args,cin", CASES)
def test_standalone_block_wiring_cpu(cls_name, where, args, cin):
    sg.set_precision("fp32")
    try:
        with emu.install():
            _check(cls_name, where, args, cin, "cpu", 1e-5, 1e-4)
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)


@pytest.mark.gpu
@pytest.mark.parametrize("precision,tol_out,tol_grad", [("fp32", 1e-5, 1e-4), ("fp16", 2e-3, 5e-3), ("bf16", 1.5e-2, 3e-2)])
@pytest.mark.parametrize("cls_name,where,args,cin", CASES)
def test_standalone_block_gpu(cls_name, where, args, cin, precision, tol_out, tol_grad):
    sg.set_precision(precision)
    try:
        _check(cls_name, where, args, cin, "cuda", tol_out, tol_grad)
    finally:
        sg.set_precision(sg.DEFAULT_PRECISION)
