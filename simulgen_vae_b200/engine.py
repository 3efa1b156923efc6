"""Host side of the B200 engine: forward/backward orchestration of the hierarchical VAE over the
hand-written CUDA kernels (kernels.py -> libsimulgen_b200.so).

The reference runs ~300 ATen/cuDNN launches per step through autograd
(/root/reference/modules/VAE_network.py:79-117, encoder.py:146-167, decoder.py:170-216).  Here the
encoder and the decoder are each ONE torch.autograd.Function; inside, a small tape records the
backward of every fused block, so no ATen kernel runs on the hot path and gradients are produced in
a known order (which the data-parallel driver uses to overlap the NCCL all-reduce).

Data layout inside the engine ("CR"): activations are [C, B, Tp], the tail t >= T of every row zero, so that
every Conv1d is one implicit GEMM  D[Cout, B*Tp] = sum_taps Wg[tap] @ shift(act).
  16-bit (tensor-core) modes: Tp = roundup(T, 8); the operand of a k-tap conv is stored as k pre-shifted planes
      [planes, C, B, Tp] (zero-filled at the row ends = the conv padding), tap j selects a plane.
  fp32 validation mode: Tp = roundup(T+2, 8); the SIMT GEMMs shift along the flattened (sample, time) axis and the
      >= 2 zero columns between samples are the conv padding (tp_of() below).
GEMM operands are bf16 / fp16 (tcgen05, fp32 accumulation) or fp32 (validation mode); norm statistics, the
residual stream, latents, KL and losses stay fp32.
"""
from __future__ import annotations

import contextlib
import os
import threading
import weakref

import torch

from . import kernels as K

# Default operand format of the tensor-core path.  tcgen05.mma.kind::f16 runs IEEE fp16 and bf16 operands at the same
# rate with the same fp32 accumulation; fp16 carries 3 more mantissa bits, and that is what it takes to meet
# north_star's "per-layer outputs AND gradients within 1e-2 relative L2" on this ~40-GEMM-deep model (measured at the
# headline shape: worst gradient 1.7e-3 with fp16 operands, 1.5e-2 with bf16 - DESIGN.md section 2).  Its narrower
# exponent is handled by the device-resident dynamic loss scaler of the Trainer (sg_scaler_state).
DEFAULT_PRECISION = "fp16"
_PRECISION = os.environ.get("SIMULGEN_B200_PRECISION", DEFAULT_PRECISION)


# 16-bit storage of the tensors BETWEEN the GEMMs and the GroupNorm / activation kernels (16-bit operand modes only):
#   y   the pre-norm conv output (GroupNorm statistics come from the fp32 accumulators in the GEMM epilogue), and
#   dx  the gradient of an activation whose only consumer is one conv (the dgrad epilogue stores the operand format).
# Both halve the bytes of three HBM passes per layer.  Default: on in fp16 mode (10 mantissa bits: adds ~3e-4 per
# layer, the gradient parity stays inside north_star's 1e-2, tests/test_parity_gpu.py), off in bf16 mode (7 bits).
_STORE16 = os.environ.get("SIMULGEN_B200_STORE16", "auto")


def store16() -> bool:
    if _STORE16 in ("0", "1"):
        return _STORE16 == "1" and _PRECISION in ("bf16", "fp16")
    return _PRECISION == "fp16"


def set_precision(p: str):
    """'fp16' (tcgen05 tensor cores, IEEE fp16 operands, fp32 accumulation; default - the backward needs loss scaling,
    which Trainer applies on the device), 'bf16' (the same kernels with bf16 operands: no loss scaling, ~8x larger
    rounding error) or 'fp32' (validation mode, SIMT)."""
    global _PRECISION
    if p not in ("bf16", "fp16", "fp32"):
        raise ValueError("precision must be 'bf16', 'fp16' or 'fp32'")
    _PRECISION = p


def get_precision() -> str:
    return _PRECISION


def tp_of(T: int, precision: str = None) -> int:
    """Row pitch of the CR layout.  bf16 (tensor-core) mode: T rounded up to 8 elements (TMA pitches are multiples
    of 16 bytes) - the taps of a k-tap conv read pre-shifted operand planes, so no zero gap between samples is
    needed.  fp32 validation mode: the SIMT GEMMs shift along the flattened (sample, time) axis and rely on a
    zero gap of >= 2 columns (k <= 5) between samples."""
    if (precision or _PRECISION) in ("bf16", "fp16"):
        return (T + 7) // 8 * 8
    return (T + 2 + 7) // 8 * 8


# ------------------------------------------------------------------------------------------------
# eps source: counter-based Philox kernel by default; tests may inject a fixed stream
# ------------------------------------------------------------------------------------------------
_ORIG_RANDN_LIKE = torch.randn_like
_rng = threading.local()


def _rng_state():
    if not hasattr(_rng, "counter"):
        _rng.counter = 0
        _rng.seed = None
        _rng.sample0 = 0
        _rng.fixed = None
    return _rng


class fixed_eps:
    """Context manager: feed the given eps tensors (reference layout, draw order) to the engine."""

    def __init__(self, eps_list):
        self.eps = list(eps_list)

    def __enter__(self):
        st = _rng_state()
        self._prev = st.fixed
        st.fixed = list(self.eps)
        return self

    def __exit__(self, *exc):
        _rng_state().fixed = self._prev
        return False


def set_sample_offset(sample0: int):
    """Global index of the first sample of the local batch (data parallel: rank * local_batch) so
    that eps depends on the global sample id, not on how the batch is split over ranks."""
    _rng_state().sample0 = int(sample0)


def draw_eps(shape, device):
    """Standard-normal tensor of the reference's shape ([B, L] or [B, C, T]), replacing
    torch.randn_like (decoder.py:221).  Order of precedence: fixed_eps() stream, a patched
    torch.randn_like (the oracle harness patches it), else the Philox kernel."""
    st = _rng_state()
    if st.fixed is not None:
        if not st.fixed:
            raise RuntimeError("fixed_eps stream exhausted")
        e = st.fixed.pop(0)
        if tuple(e.shape) != tuple(shape):
            raise RuntimeError("fixed eps has shape %s, expected %s" % (tuple(e.shape), tuple(shape)))
        return e.to(device=device, dtype=torch.float32).contiguous()
    if torch.randn_like is not _ORIG_RANDN_LIKE:
        proto = torch.empty(shape, device=device, dtype=torch.float32)
        return torch.randn_like(proto).to(torch.float32).contiguous()
    seed = torch.initial_seed()
    if st.seed != seed:
        st.seed, st.counter = seed, 0
    out = torch.empty(shape, device=device, dtype=torch.float32)
    dev_counter = getattr(st, "dev_counter", None)
    if dev_counter is not None:
        # CUDA-graph capture (Trainer): the draw index of the STEP lives in device memory (advanced inside the graph),
        # the kernel argument only carries the index of the draw within the step
        K.philox_normal(out, seed, st.graph_draw, st.sample0, dev_counter)
        st.graph_draw += 1
        return out
    K.philox_normal(out, seed, st.counter, st.sample0)
    st.counter += 1
    return out


# ------------------------------------------------------------------------------------------------
# gradient sink: persistent destinations for parameter gradients (Trainer's fused optimiser path)
# ------------------------------------------------------------------------------------------------
class GradSink:
    """Owns one flat fp32 arena for the wgrad GEMM outputs (GEMM layout, consumed as is by the fused
    spectral-norm-gradient + AdamW kernel) and one for the small vectors (biases, GroupNorm affine).
    Buffers are bump-allocated in the order backward first produces them, so the arena is laid out in
    backward-completion order: contiguous slices are all-reduced while the rest of backward runs.
    With a sink installed the autograd Functions return no parameter gradients (p.grad stays None)."""

    ALIGN = 64

    def __init__(self, weight_elems, vec_elems, device, arenas=None):
        """arenas: optional preallocated (weights, vecs) buffers - the data-parallel trainer hands in symmetric
        (peer-mapped) memory so that the other ranks can read the gradients over NVLink."""
        self.dev = device
        if arenas is not None:
            self.weights, self.vecs = arenas
            assert self.weights.numel() >= max(weight_elems, 1) and self.vecs.numel() >= max(vec_elems, 1)
        else:
            self.weights = torch.zeros(max(weight_elems, 1), dtype=torch.float32, device=device)
            self.vecs = torch.zeros(max(vec_elems, 1), dtype=torch.float32, device=device)
        self.scratch = torch.zeros(1 << 18, dtype=torch.float64, device=device)
        self.scratch_used = 0
        self.w_used = self.v_used = 0
        self.w_slots, self.v_slots = {}, {}
        self.items = {}                 # id(param) -> optimiser item description (first backward)
        self.order = []                 # params in commit order
        self.committed = 0              # arena elements whose producers have been enqueued this step
        self.on_commit = None           # callback(committed_elems) for the data-parallel driver
        self.frozen = False

    @staticmethod
    def _round(n):
        return (n + GradSink.ALIGN - 1) // GradSink.ALIGN * GradSink.ALIGN

    def begin_step(self, zero_vecs_from=0):
        """zero_vecs_from: leave the first elements of the vector arena alone (the data-parallel trainer zeroes the
        decoder's part later: the previous step's optimiser may still be reading it on another stream)."""
        self.committed = 0
        # the per-channel gradient vectors (bias, GroupNorm gain/shift) are accumulated with atomics: one memset of
        # the whole arena per step instead of three per layer; same for the fp64 scratch of the GroupNorm backward
        if zero_vecs_from > 0:
            self.vecs[zero_vecs_from:].zero_()
        else:
            self.vecs.zero_()
        self.scratch.zero_()
        self.scratch_used = 0

    def zeroed_scratch(self, n):
        """n zeroed float64 elements (valid until the next begin_step), or None when the arena is exhausted"""
        n = (n + 1) // 2 * 2
        if self.scratch_used + n > self.scratch.numel():
            return None
        out = self.scratch[self.scratch_used:self.scratch_used + n]
        self.scratch_used += n
        return out

    def weight_buffer(self, param, shape):
        key = id(param)
        n = 1
        for d in shape:
            n *= d
        if key not in self.w_slots:
            if self.frozen:
                raise RuntimeError("simulgen_b200: a parameter produced a gradient for the first time after the "
                                   "optimiser plan was built")
            if self.w_used + n > self.weights.numel():
                raise RuntimeError("simulgen_b200: GradSink weight arena exhausted")
            self.w_slots[key] = (self.w_used, n)
            self.w_used += self._round(n)
        off, n0 = self.w_slots[key]
        if n0 != n:
            raise RuntimeError("simulgen_b200: weight-gradient shape changed between steps")
        return self.weights[off:off + n].view(shape)

    def vec_buffer(self, param, n):
        key = id(param)
        if key not in self.v_slots:
            if self.frozen:
                raise RuntimeError("simulgen_b200: a parameter produced a gradient for the first time after the "
                                   "optimiser plan was built")
            if self.v_used + n > self.vecs.numel():
                raise RuntimeError("simulgen_b200: GradSink vector arena exhausted")
            self.v_slots[key] = (self.v_used, n)
            self.items[key] = dict(param=param, g=self.vecs[self.v_used:self.v_used + n])
            self.order.append(key)
            self.v_used += self._round(n)
        off, n0 = self.v_slots[key]
        if n0 != n:
            raise RuntimeError("simulgen_b200: gradient size changed between steps")
        return self.vecs[off:off + n0]

    def commit_weight(self, prep, dwg, upto=None):
        """Called once the wgrad GEMM writing `dwg` (a weight_buffer) has been enqueued.  upto: only the first `upto`
        elements of the buffer are final so far (a weight gradient produced in row chunks, so that the all-reduce of the
        first chunks overlaps the GEMMs of the later ones)."""
        key = id(prep.w)
        if key not in self.items:
            self.items[key] = dict(param=prep.w, g=dwg, u=prep.u, vv=prep.v, sigma=prep.sigma, Cout=prep.Cout,
                                   Cin=prep.Cin, Cin_p=prep.Cin_p, k=prep.k, flip=prep.flip)
            self.order.append(key)
        off, n = self.w_slots[key]
        done = off + (self._round(n) if upto is None or upto >= n else (upto // self.ALIGN * self.ALIGN))
        self.committed = max(self.committed, done)
        if self.on_commit is not None:
            self.on_commit(self.committed, force=upto is not None)


_sink = threading.local()


def set_grad_sink(sink):
    """Install (or remove, with None) the gradient sink used by the next forward/backward of this thread."""
    _sink.value = sink


def get_grad_sink():
    return getattr(_sink, "value", None)


def set_materialize_xhat(flag: bool):
    """False: VAE.forward(x) returns x_hat = None - the reconstruction only exists in registers inside the fused
    tanh + loss kernel (train.py:142 discards it); saves one fp32 [B, N, T] write per step.  Thread-local."""
    _sink.xhat = bool(flag)


def _materialize_xhat():
    return getattr(_sink, "xhat", True)


def note_grad_mode():
    """Called by the overlay modules right before Function.apply: inside Function.forward grad mode is always off and
    needs_input_grad only says that the parameters COULD take a gradient, so under torch.no_grad() (validation,
    export sweeps) the engine would otherwise record a backward tape that keeps every activation alive."""
    _sink.grad_enabled = torch.is_grad_enabled()


def _take_grad_mode():
    g = getattr(_sink, "grad_enabled", True)
    _sink.grad_enabled = True
    return g


def set_freeze_level(level: int):
    """freeze_level of the next Decoder.forward of this thread (decoder.py:170,202-207); consumed by that call."""
    _sink.freeze = int(level)


def _take_freeze_level():
    lvl = getattr(_sink, "freeze", -1)
    _sink.freeze = -1
    return lvl


class PackedBatch:
    """A batch that exists only as the packed 16-bit operand of the first encoder conv, [1, N, B, Tp] in the engine's
    operand format - what sg_assemble_batch (the resident-dataset loader) or sg_pack_input write.  VAE.forward,
    Trainer.step and the engine's train() accept it in place of the fp32 [B, N, T] tensor: the encoder consumes the
    operand as is and the reconstruction loss reads the same values (loss_target() == "operand"), so the fp32 batch is
    never materialised.  fp16 operand mode with T % 8 == 0 only."""

    def __init__(self, operand, T):
        if operand.dim() != 4 or operand.shape[0] != 1 or operand.dtype not in (torch.float16, torch.bfloat16):
            raise ValueError("PackedBatch needs a 16-bit operand of shape [1, N, B, Tp]")
        self.operand, self.T = operand, int(T)
        self.shape = (operand.shape[2], operand.shape[1], int(T))
        self.device = operand.device

    def numel(self):
        return self.shape[0] * self.shape[1] * self.shape[2]


# Target of the reconstruction losses (VAE_network.py:110-111).  "input": the fp32 tensor x [B, N, T], as the reference.
# "operand": the packed 16-bit operand of x that the first encoder conv consumes anyway ([N, B, Tp], the layout of the
# recon conv's output): 2 instead of 4 bytes per element in both passes of the reconstruction head, coalesced like y, and
# the fp32 batch need not exist at all (PackedBatch).  The target is then x rounded to fp16 (|x| <= 0.7: absolute error
# <= 2.4e-4, a fixed perturbation of the data far below the model's reconstruction error).  Default: "operand" in fp16
# mode when T % 8 == 0, "input" otherwise (bf16 has 3 bits less: the rounded target would cost gradient parity).
_LOSS_TARGET = os.environ.get("SIMULGEN_B200_LOSS_TARGET", "auto")


def set_loss_target(kind: str):
    global _LOSS_TARGET
    if kind not in ("auto", "input", "operand"):
        raise ValueError("loss target must be 'auto', 'input' or 'operand'")
    _LOSS_TARGET = kind


def loss_target(T: int = 8) -> str:
    if _PRECISION == "fp32" or T % 8:
        return "input"
    if _LOSS_TARGET == "auto":
        return "operand" if _PRECISION == "fp16" else "input"
    return _LOSS_TARGET


# Static fields (Dim2 = 1, T = 1).  The CR layout pads every row to 8 elements, i.e. 1 valid column in 8 on the two
# N-channel layers (encoder conv0, reconstruction conv + head) - the GEMMs would multiply 7 zero columns per sample and the
# head kernels stream 8x the bytes.  For T = 1 a k = 1 conv is a plain matrix product over the batch, so those two layers
# run on COMPACT tensors [C][B] instead (csrc/static_ops.cu): to the GEMM entry points a compact tensor is an activation
# with B / 8 samples of 8 valid columns; small [C <= 1024][B] tensors are converted between the two forms on the way in
# and out.  Needs B % 8 == 0, B <= 2048 and a 16-bit operand mode; SIMULGEN_B200_STATIC_COMPACT=0 keeps the padded path.
_STATIC_COMPACT = os.environ.get("SIMULGEN_B200_STATIC_COMPACT", "1") != "0"


def static_compact(B: int, T: int) -> bool:
    return _STATIC_COMPACT and T == 1 and B % 8 == 0 and 8 <= B <= 2048 and _PRECISION != "fp32"


class StaticTarget:
    """What the compact encoder input leaves for the reconstruction loss: xc [N, B] in the operand format and / or
    xt [N, B] fp32 (the transposed batch)."""
    __slots__ = ("xc", "xt")

    def __init__(self, xc, xt):
        self.xc, self.xt = xc, xt

    @staticmethod
    def operand_policy() -> bool:
        """loss_target() without its T % 8 condition: the 16-bit operand in fp16 mode, the fp32 values otherwise"""
        return (_PRECISION == "fp16") if _LOSS_TARGET == "auto" else (_LOSS_TARGET == "operand")

    def pick(self):
        return self.xc if (self.xt is None or (self.operand_policy() and self.xc is not None)) else self.xt


def set_loss_operand(op):
    """Hand the next decoder forward of this thread the packed operand of its target x (VAE.forward does)."""
    _sink.loss_op = op


def _take_loss_operand():
    op = getattr(_sink, "loss_op", None)
    _sink.loss_op = None
    return op


def take_last_packed():
    """The packed operand the last encoder forward of this thread consumed (or wrote), once."""
    op = getattr(_sink, "last_packed", None)
    _sink.last_packed = None
    return op


def set_packed_input(op):
    """Hand the next encoder forward of this thread the already packed bf16 operand [1, N, B, Tp] of its input
    (written by sg_assemble_batch together with the batch): sg_pack_input is skipped once."""
    _sink.packed = op


def _take_packed_input():
    op = getattr(_sink, "packed", None)
    _sink.packed = None
    return op


# ------------------------------------------------------------------------------------------------
# tape primitives
# ------------------------------------------------------------------------------------------------
class Act:
    """An activation in CR layout.  data: GEMM operand [planes, C, B, Tp] (bf16 / fp32; plane pl holds
    the rows shifted by pl - planes//2, see csrc/common.cuh) or None; f32: fp32 copy [C, B, Tp] or None;
    grad: fp32 gradient buffer [C, B, Tp] filled during backward."""
    __slots__ = ("data", "f32", "grad", "C", "needs_grad", "name", "grad16", "compact")

    def __init__(self, C, data=None, f32=None, needs_grad=True, name=""):
        self.C, self.data, self.f32, self.grad, self.needs_grad, self.name = C, data, f32, None, needs_grad, name
        self.grad16 = False     # True: the only gradient contributor is one dgrad GEMM that can store 16 bits
        self.compact = False    # True: data is the compact static form [1, C, B / 8, 8] = [C][B] (static_compact())

    def center(self):
        return self.data[self.data.shape[0] // 2]

    def as_f32(self):
        if self.f32 is not None:
            return self.f32
        if self.data is not None and self.data.dtype == torch.float32:
            return self.center()
        raise RuntimeError("activation %s has no fp32 copy" % self.name)

    def residual_source(self):
        return self.f32 if self.f32 is not None else self.center()


class Ext:
    """A tensor that crosses the autograd boundary (input or output of an engine Function)."""
    __slots__ = ("tensor", "grad")

    def __init__(self, tensor):
        self.tensor, self.grad = tensor, None


class Ctx:
    def __init__(self, B, T, device, record, capture=None):
        self.B, self.T, self.Tp = B, T, tp_of(T)
        self.dev = device
        self.tape = [] if record else None
        self.op_dtype = {"bf16": torch.bfloat16, "fp16": torch.float16}.get(_PRECISION, torch.float32)
        _sink.centre_only = bool(_STATIC_CENTRE and T == 1 and self.op_dtype != torch.float32)    # see _k()
        self.pgrads = {}
        self.capture = capture
        self.sink = get_grad_sink() if record else None
        self._side_used = False
        self._side_keep = []            # tensors read by the weight-gradient stream: kept alive until the next join
        self.prepared = {}              # id(module) -> _Prep from the batched preparation
        self.prep_record = []           # (module, kind) in execution order when preparing layer by layer
        self.prep_root = None

    # -- allocation helpers ------------------------------------------------------------------
    def f32(self, *shape):
        return torch.empty(shape, dtype=torch.float32, device=self.dev)

    def op(self, *shape):
        return torch.empty(shape, dtype=self.op_dtype, device=self.dev)

    def f64(self, *shape):
        return torch.empty(shape, dtype=torch.float64, device=self.dev)

    def grad_buf(self, act: Act):
        """(buffer, accumulate_flag) for adding a gradient contribution to `act`."""
        if act.grad is None:
            act.grad = self.op(act.C, self.B, self.Tp) if act.grad16 else self.f32(act.C, self.B, self.Tp)
            return act.grad, 0
        return act.grad, 1

    def set_pgrad(self, param, grad):
        if param is None or grad is None or self.sink is not None:
            return                      # with a sink the gradient already sits in its persistent buffer
        key = id(param)
        if key in self.pgrads:
            K.axpy(self.pgrads[key], grad, 1.0, True)
        else:
            self.pgrads[key] = grad

    def vec_grad(self, param, n):
        """fp32 [n] buffer for the gradient of a bias / GroupNorm affine vector (None if there is no such
        parameter).  Persistent when a sink is installed."""
        if param is None:
            return None
        if self.sink is not None and param.requires_grad:
            return self.sink.vec_buffer(param, n)
        return self.f32(n)

    def weight_grad_buf(self, prep, *shape):
        """fp32 buffer for the wgrad GEMM output of a layer (GEMM layout)."""
        if self.sink is not None and prep.sn:
            return self.sink.weight_buffer(prep.w, shape)
        return self.f32(*shape)


    def cap(self, name, act: Act):
        if self.capture is None:
            return
        src = act.f32 if act.f32 is not None else act.center()
        out = torch.empty(self.B, act.C, self.T, dtype=torch.float32, device=self.dev)
        K.unpack_f32(src.float() if src.dtype != torch.float32 else src, out, self.T)
        self.capture[name] = out

    def run_backward(self):
        for fn in reversed(self.tape):
            fn()
        self.tape = None
        self.join_side()

    # -- weight-gradient stream -----------------------------------------------------------------------
    # The wgrad GEMM of a layer only feeds the optimiser, while the activation-gradient chain
    # (GroupNorm/activation backward -> dgrad -> next layer) is what the rest of backward waits for.  wgrad and the
    # spectral-norm gradient therefore run on a second stream: the two tensor-core GEMMs still take turns on the
    # SMs (each persistent CTA needs > 200 KB of shared memory), but the HBM-bound streaming kernels of the chain
    # co-reside with whichever GEMM is running instead of waiting behind it.
    def side(self, *tensors, flops=None):
        """Context manager: run the enclosed launches on the weight-gradient stream, after everything issued so far
        on the current stream; `tensors` (its inputs) are kept alive until the compute stream has joined it again."""
        return _SideStream(self, tensors, flops)

    def join_side(self):
        if self._side_used:
            torch.cuda.current_stream(self.dev).wait_stream(_side_stream_of(self.dev))
            self._side_used = False
            # the compute stream is now ordered after every side-stream reader: the operands may go back to the
            # caching allocator (they were allocated on the compute stream, so plain stream order covers their reuse;
            # no record_stream(), whose event-deferred frees make the allocator's behaviour timing dependent)
            self._side_keep.clear()


_SIDE_STREAMS = {}
# "1": every wgrad on the side stream; "small" (default): only the wgrads below _OVERLAP_MAX_FLOP (the ones that cannot
# fill the GPU on their own; the big GEMMs keep the machine to themselves); "0": off.  Same-box A/B at the headline
# shape, B=64, three runs each: off 1591, all 1609, small 1624 samples/s.
_OVERLAP_MODE = os.environ.get("SIMULGEN_B200_OVERLAP_WGRAD", "small")
_OVERLAP_WGRAD = _OVERLAP_MODE != "0"
_OVERLAP_MAX_FLOP = float(os.environ.get("SIMULGEN_B200_OVERLAP_MAX_GFLOP", "300")) * 1e9 if _OVERLAP_MODE == "small" else None


def _side_stream_of(dev):
    key = (dev.type, dev.index)
    st = _SIDE_STREAMS.get(key)
    if st is None:
        st = torch.cuda.Stream(device=dev)
        _SIDE_STREAMS[key] = st
    return st


def order_after_side_stream(dev):
    """Make the current stream wait for everything issued on the weight-gradient stream (no-op when the current stream
    IS that stream - it waited for the compute stream when its context was entered - or when it was never used)."""
    if dev.type != "cuda":
        return
    side = _SIDE_STREAMS.get((dev.type, dev.index))
    if side is not None:
        cur = torch.cuda.current_stream(dev)
        if cur != side:
            cur.wait_stream(side)


class _SideStream:
    def __init__(self, ctx, tensors, flops=None):
        self.ctx, self.tensors = ctx, tensors
        self.active = _OVERLAP_WGRAD and ctx.dev.type == "cuda" and \
            (_OVERLAP_MAX_FLOP is None or flops is None or flops < _OVERLAP_MAX_FLOP)
        self.cm = None

    def __enter__(self):
        if self.active:
            side = _side_stream_of(self.ctx.dev)
            side.wait_stream(torch.cuda.current_stream(self.ctx.dev))
            self.ctx._side_keep.extend(t for t in self.tensors if t is not None)
            self.cm = torch.cuda.stream(side)
            self.cm.__enter__()
            self.ctx._side_used = True
        return self

    def __exit__(self, *exc):
        if self.cm is not None:
            self.cm.__exit__(*exc)
        return False


# ------------------------------------------------------------------------------------------------
# spectral norm / weight preparation
# ------------------------------------------------------------------------------------------------
class _Prep:
    __slots__ = ("w", "u", "v", "sigma", "wg", "Cin", "Cout", "k", "Cin_p", "so", "si", "flip", "sn", "version", "mod")


_SIGMAS = {}                                # id(weight parameter) -> (weakref, persistent [1] fp32 sigma tensor)
_PREP_CACHE = weakref.WeakKeyDictionary()   # encoder / decoder module -> _PrepCache


def sigma_of(param):
    """Persistent device scalar for the spectral norm of `param`: written by every forward, read by the
    backward and by the fused optimiser step (whose launch table stores its address)."""
    key = id(param)
    ent = _SIGMAS.get(key)
    if ent is None or ent[0]() is not param or ent[1].device != param.device:
        t = torch.ones(1, dtype=torch.float32, device=param.device)
        _SIGMAS[key] = (weakref.ref(param, lambda _, key=key: _SIGMAS.pop(key, None)), t)
        return t
    return ent[1]


class _PrepCache:
    """Everything the batched spectral-norm preparation of one sub-network needs, built from the layer
    sequence recorded during the first forward: persistent operand copies of the weights, the launch table
    (kernels.SnPlan) and the _Prep objects handed to the graph."""

    def __init__(self, record, op_dtype, device):
        self.op_dtype, self.device = op_dtype, device
        self.preps = {}
        self.mods = []
        layers = []
        for mod, kind in record:
            p = _prep_shapes(mod, kind)
            p.sigma = sigma_of(p.w)
            p.wg = torch.empty(p.k, p.Cout, p.Cin_p, dtype=op_dtype, device=device) if kind != "linear" else None
            p.version = 0
            self.preps[id(mod)] = p
            self.mods.append(mod)
            layers.append(dict(w=p.w.data, u=p.u, v=p.v, sigma=p.sigma, wg=p.wg, H=p.Cout, Cin=p.Cin, k=p.k,
                               Cin_p=p.Cin_p, so=p.so, si=p.si, flip=p.flip))
        self.key = self._key()
        self.plan = K.SnPlan(layers, device, op_dtype)

    def _key(self):
        k = []
        for mod in self.mods:
            p = self.preps[id(mod)]
            w, u, v, sn = _sn_tensors(None, mod)
            k.append((w.data_ptr(), u.data_ptr() if sn else 0, v.data_ptr() if sn else 0, sn))
        return tuple(k)

    def valid(self, ctx, training):
        if self.op_dtype != ctx.op_dtype or self.device != ctx.dev:
            return False
        if any(m.training != training for m in self.mods):
            return False
        return self._key() == self.key


def prepare_all(ctx, root):
    """Batched power iteration + weight packing for every layer `root` executed in its first forward.
    Returns False when there is no (valid) cache yet: the graph then prepares layer by layer and records."""
    ctx.prep_root = root
    cache = _PREP_CACHE.get(root)
    training = root.training
    if cache is None or not cache.valid(ctx, training):
        if cache is not None:
            del _PREP_CACHE[root]
        return False
    K.sn_prepare(cache.plan, training)
    for mod in cache.mods:
        p = cache.preps[id(mod)]
        if p.sn:
            p.version = _bump_version(mod) if training else getattr(mod, "_sg_sn_version", 0)
    ctx.prepared = cache.preps
    return True


def finish_prepare(ctx):
    """After the first forward of a sub-network: turn the recorded layer sequence into a cache."""
    root = ctx.prep_root
    if root is None or ctx.prepared or not ctx.prep_record:
        return
    if len(set(id(m) for m, _ in ctx.prep_record)) != len(ctx.prep_record):
        return                              # a layer used twice in one forward: keep the per-layer path
    _PREP_CACHE[root] = _PrepCache(ctx.prep_record, ctx.op_dtype, ctx.dev)


def _prep_shapes(mod, kind):
    p = _Prep()
    w, u, v, sn = _sn_tensors(None, mod)
    if kind == "linear":
        O, In = w.shape
        p.Cin, p.Cout, p.k, p.so, p.si, p.flip, p.Cin_p = In, O, 1, In, 1, 0, In
    else:
        if kind == "convT":
            Cin, Cout, k = w.shape
            so, si, flip = k, Cout * k, 1
        else:
            Cout, Cin, k = w.shape
            so, si, flip = Cin * k, k, 0
        p.Cin, p.Cout, p.k, p.so, p.si, p.flip = Cin, Cout, k, so, si, flip
        p.Cin_p = (Cin + 7) // 8 * 8
    p.w, p.u, p.v, p.sn, p.mod = w, u, v, sn, mod
    p.wg = None
    return p


def _sn_tensors(ctx, mod):
    if hasattr(mod, "weight_orig"):
        return mod.weight_orig, mod.weight_u, mod.weight_v, True
    return mod.weight, None, None, False


def _bump_version(mod):
    v = getattr(mod, "_sg_sn_version", 0) + 1
    object.__setattr__(mod, "_sg_sn_version", v)
    return v


def prep_conv(ctx: Ctx, mod, transposed=False) -> _Prep:
    """Power iteration (training) / sigma (eval) + normalised, permuted, cast weight for the GEMMs.
    common.py:15-37 -> spectral_norm.py:62-114; ConvTranspose1d uses dim=1 (spectral_norm.py:329-333).
    Served from the batched preparation (prepare_all) when the sub-network has a cache."""
    p = ctx.prepared.get(id(mod))
    if p is not None:
        return p
    kind = "convT" if transposed else "conv"
    ctx.prep_record.append((mod, kind))
    p = _prep_shapes(mod, kind)
    if p.sn:
        p.sigma = sigma_of(p.w)
        K.sn_power_iter(p.w, p.u, p.v, p.sigma, p.Cout, p.Cin, p.k, p.so, p.si, mod.training)
        p.version = _bump_version(mod) if mod.training else getattr(mod, "_sg_sn_version", 0)
    else:
        p.sigma = torch.ones(1, dtype=torch.float32, device=ctx.dev)
        p.version = 0
    p.wg = ctx.op(p.k, p.Cout, p.Cin_p)
    K.sn_pack_weight(p.w, p.sigma, p.wg, p.Cout, p.Cin, p.Cin_p, p.k, p.so, p.si, p.flip)
    return p


def prep_linear(ctx: Ctx, mod) -> _Prep:
    p = ctx.prepared.get(id(mod))
    if p is not None:
        return p
    ctx.prep_record.append((mod, "linear"))
    p = _prep_shapes(mod, "linear")
    if p.sn:
        p.sigma = sigma_of(p.w)
        K.sn_power_iter(p.w, p.u, p.v, p.sigma, p.Cout, p.Cin, 1, p.Cin, 1, mod.training)
        p.version = _bump_version(mod) if mod.training else getattr(mod, "_sg_sn_version", 0)
    else:
        p.sigma = torch.ones(1, dtype=torch.float32, device=ctx.dev)
        p.version = 0
    return p


def _weight_grad(ctx: Ctx, mod, p: _Prep, dwg):
    """dWg (fp32, GEMM layout, gradient wrt W/sigma) -> gradient wrt weight_orig (reference layout)."""
    if p.sn and getattr(mod, "_sg_sn_version", 0) != p.version:
        raise RuntimeError("simulgen_b200: spectral-norm state of a layer advanced between forward and backward "
                           "(two training forwards before one backward are not supported)")
    if ctx.sink is not None and p.sn:
        ctx.sink.commit_weight(p, dwg)  # consumed in the GEMM layout by the fused optimiser kernel
        return
    if ctx.sink is not None:
        grad = ctx.sink.vec_buffer(p.w, p.w.numel()).view(p.w.shape)
    else:
        grad = torch.empty_like(p.w)
    if p.sn:
        u, v = p.u, p.v
    else:
        u = torch.zeros(p.Cout, dtype=torch.float32, device=ctx.dev)
        v = torch.zeros(p.Cin * p.k, dtype=torch.float32, device=ctx.dev)
    K.sn_weight_grad(dwg, p.w, u, v, p.sigma, grad, p.Cout, p.Cin, p.Cin_p, p.k, p.so, p.si, p.flip)
    ctx.set_pgrad(p.w, grad)


# ------------------------------------------------------------------------------------------------
# fused blocks
# ------------------------------------------------------------------------------------------------
# Static fields (T = 1): with "same" padding a k-tap conv only ever sees its CENTRE tap (the other taps multiply the zero
# padding), so in the 16-bit modes every conv runs as a k = 1 GEMM on the centre slice of its weight copy, activations and
# gradients carry one operand plane instead of k, and the other taps of the weight gradient are exactly zero.
# SIMULGEN_B200_STATIC_CENTRE=0 keeps the k-tap GEMMs.
_STATIC_CENTRE = os.environ.get("SIMULGEN_B200_STATIC_CENTRE", "1") != "0"


def _set_centre_only(ctx):
    _sink.centre_only = bool(_STATIC_CENTRE and ctx.T == 1 and ctx.op_dtype != torch.float32)


def _centre_only() -> bool:
    return getattr(_sink, "centre_only", False)


def _k(conv) -> int:
    """operand planes a consumer conv needs from its producer"""
    return 1 if _centre_only() else int(conv.kernel_size[0])


def conv_block(ctx: Ctx, conv, gn, a_in: Act, act, res: Act = None, res_scale=1.0, post_gelu=False,
               want_f32=False, out_op_view=None, out_planes=1, transposed=False, name="") -> Act:
    """Conv1d / ConvTranspose1d (+ GroupNorm) (+ activation) (+ residual) (+ trailing GELU).
    Covers encoder.py:29-46, common.py:78-162, decoder.py:27-33,150-166."""
    B, T, Tp = ctx.B, ctx.T, ctx.Tp
    p = prep_conv(ctx, conv, transposed)
    y16 = gn is not None and ctx.op_dtype != torch.float32 and store16() and K.conv_out16_ok(p.Cout)
    y = ctx.op(p.Cout, B, Tp) if y16 else ctx.f32(p.Cout, B, Tp)
    plain = gn is None and act == K.ACT_NONE and res is None and not post_gelu
    G = gn.num_groups if gn is not None else 0
    stats = None
    res_t = None
    centre = _centre_only() and p.k > 1
    wg = p.wg[p.k // 2:p.k // 2 + 1] if centre else p.wg                       # [1][Cout][Cin_p]: the centre tap
    a_data = a_in.data
    if centre and a_data.shape[0] > 1:
        a_data = a_data[a_data.shape[0] // 2:a_data.shape[0] // 2 + 1]
    if a_in.compact:
        # static fields, first encoder conv: the GEMM runs on the compact [N][B] operand (no padding columns), its small
        # [Cout][B] result is expanded to the padded layout the rest of the network uses
        assert p.k == 1 and gn is not None and not a_in.needs_grad
        y16 = False
        y = ctx.f32(p.Cout, B, Tp)
        yc = ctx.f32(p.Cout, B // 8, 8)
        K.conv_fprop(p.wg, a_in.data, conv.bias, yc, p.Cin)
        K.rows_expand_f32(yc, y)
        stats = ctx.f32(B, G, 2)
        K.gn_stats(y, stats, T, G)
    elif gn is not None:
        stats = ctx.f32(B, G, 2)
        K.conv_fprop_gn(wg, a_data, conv.bias, y, p.Cin, stats, T, G)         # statistics from the GEMM epilogue
    else:
        K.conv_fprop(wg, a_data, conv.bias, y, p.Cin)
    if plain:
        out = Act(p.Cout, data=None, f32=y, name=name)
    else:
        out_op = out_op_view if out_op_view is not None else ctx.op(out_planes, p.Cout, B, Tp)
        out_f32 = ctx.f32(p.Cout, B, Tp) if (want_f32 and ctx.op_dtype != torch.float32) else None
        res_t = res.residual_source() if res is not None else None
        K.gn_act_fwd(y, stats, gn.weight if gn is not None else None, gn.bias if gn is not None else None,
                     res_t, res_scale, act, post_gelu, out_op, out_f32, T, G)
        out = Act(p.Cout, data=out_op, f32=out_f32, name=name)
    if name:
        ctx.cap(name, out)

    if ctx.tape is not None:
        def bwd():
            g = out.grad
            if g is None:
                return
            out.grad = None
            dy = ctx.op(1 if centre else p.k, p.Cout, B, Tp)          # k planes: dgrad reads dy shifted by the taps
            dgamma = ctx.vec_grad(gn.weight, p.Cout) if gn is not None else None
            dbeta = ctx.vec_grad(gn.bias, p.Cout) if gn is not None else None
            dbias = ctx.vec_grad(conv.bias, p.Cout)
            dres, acc = (None, 0)
            if res is not None and res.needs_grad:
                dres, acc = ctx.grad_buf(res)
            ws = None
            if ctx.sink is not None and (dbias is None or conv.bias.requires_grad) and \
                    (gn is None or (gn.weight.requires_grad and gn.bias.requires_grad)):
                ws = ctx.sink.zeroed_scratch(2 * B * max(G, 1) + 2)     # arena buffers: zeroed once per step
            K.gn_act_bwd(y, stats, gn.weight if gn is not None else None, gn.bias if gn is not None else None,
                         res_t, res_scale, act, post_gelu, g, dy, dgamma, dbeta, dbias, dres, acc, T, G, ws)
            if gn is not None:
                ctx.set_pgrad(gn.weight, dgamma)
                ctx.set_pgrad(gn.bias, dbeta)
            ctx.set_pgrad(conv.bias, dbias)
            # dgrad first (the chain waits for it), after the previous layer's wgrad has left the SMs: the two
            # tensor-core GEMMs never compete; this layer's wgrad then runs on the side stream underneath the NEXT
            # layer's (HBM-bound) GroupNorm / activation backward
            if a_in.needs_grad:
                ctx.join_side()
                dx, acc_in = ctx.grad_buf(a_in)
                K.conv_dgrad(wg, dy, dx, p.Cin, bool(acc_in))
            if p.w.requires_grad:
                dyw = dy
                if a_in.compact:                           # [Cout][B] columns of dy against the compact input
                    dyw = ctx.op(1, p.Cout, B // 8, 8)
                    K.rows_compact16(dy[0], dyw[0])
                with ctx.side(dyw, a_data, flops=2.0 * p.Cin * p.Cout * (1 if centre else p.k) * B * T):
                    dwg = ctx.weight_grad_buf(p, p.k, p.Cout, p.Cin_p)
                    if centre:                             # T = 1: only the centre tap has a gradient
                        dwg.zero_()
                        K.conv_wgrad(dyw, a_data, dwg[p.k // 2:p.k // 2 + 1], p.Cin)
                        _weight_grad(ctx, conv, p, dwg)
                    else:
                        _wgrad(ctx, conv, p, dyw, a_data, dwg)
        ctx.tape.append(bwd)
    return out


# Data parallel: a weight gradient of more than _CHUNK_WGRAD_ELEMS elements (encoder conv0: 97 M, the LAST gradient of the
# step - nothing is left to hide its all-reduce behind) is produced in row chunks, each committed to the gradient sink as
# soon as its GEMM is enqueued: the collective of chunk i runs underneath the GEMM of chunk i + 1.
_CHUNK_WGRAD_ELEMS = int(float(os.environ.get("SIMULGEN_B200_CHUNK_WGRAD_MELEMS", "48")) * 1e6)
_CHUNK_WGRAD_PARTS = int(os.environ.get("SIMULGEN_B200_CHUNK_WGRAD_PARTS", "4"))
_CHUNK_WGRAD_ALIGN = int(os.environ.get("SIMULGEN_B200_CHUNK_WGRAD_ALIGN", "256"))      # rows per chunk: whole CTA-pair tiles


def _wgrad(ctx: Ctx, conv, p: _Prep, dy, act, dwg):
    sink = ctx.sink
    chunked = (sink is not None and sink.on_commit is not None and p.sn and p.k == 1 and dy.shape[0] == 1 and
               _CHUNK_WGRAD_PARTS > 1 and dwg.numel() >= _CHUNK_WGRAD_ELEMS and p.Cout >= 2 * _CHUNK_WGRAD_ALIGN)
    if not chunked:
        K.conv_wgrad(dy, act, dwg, p.Cin)
        _weight_grad(ctx, conv, p, dwg)
        return
    if p.sn and getattr(conv, "_sg_sn_version", 0) != p.version:
        raise RuntimeError("simulgen_b200: spectral-norm state of a layer advanced between forward and backward")
    al = _CHUNK_WGRAD_ALIGN
    rows = max(al, (p.Cout // _CHUNK_WGRAD_PARTS + al - 1) // al * al)
    for m0 in range(0, p.Cout, rows):
        m1 = min(p.Cout, m0 + rows)
        K.conv_wgrad(dy[:, m0:m1], act, dwg[:, m0:m1], p.Cin)
        sink.commit_weight(p, dwg, upto=m1 * p.Cin_p)


def cgg_seq(ctx: Ctx, seq, a_in: Act, res: Act = None, res_scale=0.1, post_gelu=False, want_f32=False,
            out_op_view=None, out_planes=1, name="") -> Act:
    """nn.Sequential of (Conv1d, GroupNorm, GELU) triples; the optional residual / trailing GELU /
    fp32 copy apply to the last triple (common.py:101-102,124-125,161-162)."""
    n = len(seq) // 3
    a = a_in
    for i in range(n):
        last = i == n - 1
        a = conv_block(ctx, seq[3 * i], seq[3 * i + 1], a, K.ACT_GELU,
                       res=res if last else None, res_scale=res_scale if (last and res is not None) else 1.0,
                       post_gelu=post_gelu if last else False, want_f32=want_f32 if last else False,
                       out_op_view=out_op_view if last else None,
                       out_planes=out_planes if last else _k(seq[3 * (i + 1)]), name=name if last else "")
        if not last and ctx.op_dtype != torch.float32 and store16() and a.C > 256 and K.conv_out16_ok(a.C):
            a.grad16 = True     # sole consumer: the next conv, whose dgrad GEMM (Cin = a.C rows, CTA-pair kernel) stores 16 bits
    return a


def head(ctx: Ctx, lin, h: Act, ext_out: bool = True):
    """Linear over the flattened [C*T] features (encoder.py:158-165).  Returns an Ext (or None when
    only the spectral-norm state has to advance: the two heads whose outputs the reference never uses)."""
    p = prep_linear(ctx, lin)
    if not ext_out:
        return None
    B, T = ctx.B, ctx.T
    O = p.Cout
    out = ctx.f32(B, O)
    hf = h.as_f32()
    K.head_fwd(hf, p.w, p.sigma, lin.bias, out, T)
    ext = Ext(out)
    if ctx.tape is not None:
        def bwd():
            g = ext.grad
            if g is None:
                return
            dwn = ctx.weight_grad_buf(p, 1, O, p.Cin).view(O, p.Cin)
            dbias = ctx.vec_grad(lin.bias, O)
            dh, acc = ctx.grad_buf(h) if h.needs_grad else (None, 0)
            K.head_bwd(hf, p.w, p.sigma, g, dwn, dbias, dh, acc, T)
            ctx.set_pgrad(lin.bias, dbias)
            if p.w.requires_grad:
                _weight_grad(ctx, lin, p, dwn.view(1, O, p.Cin))
        ctx.tape.append(bwd)
    return ext


def latent_seq(ctx: Ctx, seq, z: Ext, out_op_view=None, out_planes=1, name="") -> Act:
    """Linear(d, d*T) -> Unflatten -> Conv k5 -> GroupNorm -> GELU (decoder.py:131-148)."""
    lin, conv, gn = seq[0], seq[2], seq[3]
    B, T, Tp = ctx.B, ctx.T, ctx.Tp
    p = prep_linear(ctx, lin)
    D = p.Cin
    a = Act(D, data=ctx.op(_k(conv), D, B, Tp), name=name + ".lin")
    K.latent_fwd(z.tensor, p.w, p.sigma, lin.bias, a.data, T)
    if ctx.tape is not None:
        def bwd():
            g = a.grad
            if g is None:
                return
            a.grad = None
            dwn = ctx.weight_grad_buf(p, 1, D * T, D).view(D * T, D)
            dbias = ctx.vec_grad(lin.bias, D * T)
            dz = ctx.f32(B, D)
            K.latent_bwd(z.tensor, p.w, p.sigma, g, dwn, dbias, dz, T)
            ctx.set_pgrad(lin.bias, dbias)
            if p.w.requires_grad:
                _weight_grad(ctx, lin, p, dwn.view(1, D * T, D))
            if z.grad is None:
                z.grad = dz
            else:
                K.axpy(z.grad, dz, 1.0, True)
        ctx.tape.append(bwd)
    return conv_block(ctx, conv, gn, a, K.ACT_GELU, out_op_view=out_op_view, out_planes=out_planes, name=name)


def _freeze_latent(ctx: Ctx, dec, i, freeze_level, h, zs_f32, zs_next: Act):
    """mode == "fix" and i < freeze_level (decoder.py:202-207): the first freeze_level + 1 such calls draw z and append it
    to dec.zs; later calls reuse dec.zs[i + 1] (the reference's own indexing, kept as written).  zs_f32 = h + z as just
    computed; inference only (no caller of the reference trains through it)."""
    if ctx.tape is not None:
        raise RuntimeError("simulgen_b200: freeze_level >= 0 is an inference feature (run it under torch.no_grad())")
    B, T, C = ctx.B, ctx.T, zs_next.C
    if len(dec.zs) < freeze_level + 1:
        K.axpy(zs_f32, h, -1.0, True)                       # z = (h + z) - h, CR layout
        z = torch.empty(B, C, T, dtype=torch.float32, device=ctx.dev)
        K.unpack_f32(zs_f32, z, T)
        dec.zs.append(z)
        return
    z = dec.zs[i + 1]                                       # IndexError / shape error exactly where the reference has them
    if tuple(z.shape) != (B, C, T):
        raise RuntimeError("The size of tensor a (%s) must match the size of tensor b (%s)" % (tuple(z.shape), (B, C, T)))
    zc = ctx.f32(1, C, B, ctx.Tp)
    K.pack_input(z.to(ctx.dev, torch.float32).contiguous(), zc, T)
    K.axpy(zc[0], h, 1.0, True)                             # decoder_out + z (decoder.py:179)
    K.gn_act_fwd(zc[0], None, None, None, None, 1.0, K.ACT_NONE, False, zs_next.data, None, T, 0)


# ------------------------------------------------------------------------------------------------
# encoder / decoder graphs
# ------------------------------------------------------------------------------------------------
def encoder_graph(ctx: Ctx, enc, x):
    """encoder.py:146-167.  Returns (last Ext [B, 2*z_dim], [xs Ext ...] in the reference's reversed
    order without the deepest level)."""
    B, N, T = x.shape
    _set_centre_only(ctx)
    prepare_all(ctx, enc)
    packed = _take_packed_input()
    if isinstance(x, PackedBatch):
        packed = x.operand
        if tuple(packed.shape) != (1, N, B, ctx.Tp) or packed.dtype != ctx.op_dtype:
            raise RuntimeError("simulgen_b200: PackedBatch operand %s / %s does not fit the %s mode (expected %s)"
                               % (tuple(packed.shape), packed.dtype, _PRECISION, (1, N, B, ctx.Tp)))
    conv0 = enc.encoder_blocks[0].module_list[0]._seq[0]
    if packed is not None and tuple(packed.shape) == (1, N, B, ctx.Tp) and packed.dtype == ctx.op_dtype:
        a = Act(N, data=packed, needs_grad=False, name="x")
        _sink.last_packed = a.data if ctx.op_dtype != torch.float32 else None
    elif static_compact(B, T) and int(conv0.kernel_size[0]) == 1 and not isinstance(x, PackedBatch):
        # static fields: the compact operand [N][B] (and the transposed fp32 batch when the loss reads fp32)
        a = Act(N, data=ctx.op(1, N, B // 8, 8), needs_grad=False, name="x")
        a.compact = True
        tgt = StaticTarget(a.data.view(N, B), None)
        if not StaticTarget.operand_policy():
            tgt.xt = torch.empty(N, B, dtype=torch.float32, device=ctx.dev)
        K.pack_static(x, a.data, tgt.xt)
        _sink.last_packed = tgt
    else:
        a = Act(N, data=ctx.op(1, N, B, ctx.Tp), needs_grad=False, name="x")
        K.pack_input(x, a.data, T)
        _sink.last_packed = a.data if ctx.op_dtype != torch.float32 else None
    L = len(enc.encoder_blocks)
    xs = []
    h = None
    for i in range(L):
        res_seq = enc.encoder_residual_blocks[i].seq
        a = cgg_seq(ctx, enc.encoder_blocks[i].module_list[0]._seq, a, want_f32=True, out_planes=_k(res_seq[0]))
        nxt = _k(enc.encoder_blocks[i + 1].module_list[0]._seq[0]) if i + 1 < L else 1
        h = cgg_seq(ctx, res_seq, a, res=a, res_scale=0.1, want_f32=True, out_planes=nxt,
                    name="encoder.level%d" % i)
        # xs_linear[L-1] is evaluated by the reference but its output is dropped (encoder.py:167
        # xs[:-1]): only its power iteration is observable.  xs_linear[0] is returned (exported to
        # xs.npy by the callers) but never consumed by the decoder (decoder.py:184,190): no gradient.
        live = i < L - 1
        xs.append(head(ctx, enc.xs_linear[i], h, ext_out=live))
        a = h
    last = head(ctx, enc.last_x_linear, h)
    finish_prepare(ctx)
    return last, [e for e in xs[:-1][::-1]]


def _static_recon(ctx: Ctx, conv, gn, p, out: Act, x, handoff, lossfun: str, kls, want_xhat=False):
    """Reconstruction head of the static configuration (T = 1) on compact tensors [C][B] (decoder.py:117-121 +
    VAE_network.py:110-111; with or without a backward tape - a gradient wrt x_hat itself is not supported here).  x_hat, when wanted, is written
    transposed ([N][B], coalesced) and returned as a [B, N, 1] VIEW of that buffer.  The recon conv reads the [Cin][B] columns of the
    last decoder activation and writes y [N][B] in the operand format (1/8 of the padded bytes, no GEMM work on padding);
    GroupNorm statistics, Tanh, both losses and the reductions of the GroupNorm backward are taken in two streaming passes
    over y, the backward writes dy [N][B] in one more, and the dgrad / wgrad GEMMs run on the compact operands."""
    B, T, dev = ctx.B, ctx.T, ctx.dev
    N, G = p.Cout, gn.num_groups
    B8 = B // 8
    if isinstance(handoff, StaticTarget) and handoff.pick() is not None and tuple(handoff.pick().shape) == (N, B):
        target = handoff.pick()
    else:                                                   # decoder called on its own: transpose the target here
        xt = torch.empty(N, B, dtype=torch.float32, device=dev)
        xc = torch.empty(N, B, dtype=ctx.op_dtype, device=dev)
        K.pack_static(x, xc, xt)
        target = StaticTarget(xc, xt).pick()
    hc = ctx.op(1, p.Cin, B8, 8)
    K.rows_compact16(out.data[0], hc[0])
    yc = torch.empty(N, B8, 8, dtype=ctx.op_dtype, device=dev)
    K.conv_fprop16(p.wg, hc, conv.bias, yc, p.Cin)
    y2 = yc.view(N, B)
    stats = ctx.f32(B, G, 2)
    K.static_stats(y2, stats, G)
    loss_kind = K.LOSS_KINDS.get(lossfun, 0)
    sums = ctx.f64(2)
    ws = K.static_recon_ws(N, B, G, dev)
    xhat_t = torch.empty(N, B, dtype=torch.float32, device=dev) if want_xhat else None
    K.static_recon_fwd(y2, stats, gn.weight, gn.bias, target, sums, ws, G, loss_kind, xhat_t)
    inv_numel = 1.0 / float(B * N * T)
    both = ctx.f32(2)
    K.scale_f64_to_f32(sums, both, inv_numel)
    x_hat = xhat_t.t().unsqueeze(-1) if want_xhat else None
    res = dict(x_hat=x_hat, recon=Ext(both[0:1]), mse=Ext(both[1:2]), kls=kls)
    xhat_ext = Ext(x_hat)
    res["x_hat_ext"] = xhat_ext

    def recon_bwd(out=out):
        g_loss, g_mse = res["recon"].grad, res["mse"].grad
        if xhat_ext.grad is not None:
            raise RuntimeError("simulgen_b200: a gradient wrt x_hat itself needs the padded reconstruction head "
                               "(set SIMULGEN_B200_STATIC_COMPACT=0)")
        if g_loss is None and g_mse is None:
            return
        dyc = ctx.op(1, N, B8, 8)
        dgamma, dbeta, dbias = ctx.vec_grad(gn.weight, N), ctx.vec_grad(gn.bias, N), ctx.vec_grad(conv.bias, N)
        K.static_recon_bwd(y2, stats, gn.weight, gn.bias, target, g_loss, g_mse, inv_numel, ws, dyc.view(N, B), dgamma, dbeta,
                           dbias, G, loss_kind)
        ctx.set_pgrad(gn.weight, dgamma)
        ctx.set_pgrad(gn.bias, dbeta)
        ctx.set_pgrad(conv.bias, dbias)
        dx, acc = ctx.grad_buf(out)
        if dx.dtype != torch.float32:
            raise RuntimeError("simulgen_b200: static reconstruction head expects an fp32 gradient buffer")
        dxc = ctx.f32(p.Cin, B8, 8)
        K.conv_dgrad(p.wg, dyc, dxc, p.Cin, False)
        K.rows_expand_f32(dxc, dx, bool(acc))
        if p.w.requires_grad:
            with ctx.side(dyc, hc, flops=2.0 * p.Cin * p.Cout * p.k * B * T):
                dwg = ctx.weight_grad_buf(p, p.k, p.Cout, p.Cin_p)
                K.conv_wgrad(dyc, hc, dwg, p.Cin)
                _weight_grad(ctx, conv, p, dwg)
    if ctx.tape is not None:
        ctx.tape.append(recon_bwd)
    return res


def decoder_graph(ctx: Ctx, dec, z: Ext, xs, x, lossfun: str, mode: str, want_xhat=True):
    """decoder.py:170-216 (+ the fused reconstruction losses of VAE_network.py:110-111 when x is given).
    xs: list of Ext (xs[i] feeds level i); entries beyond the used levels are ignored, None entries skipped.
    Returns dict(x_hat, recon Ext|None, mse Ext|None, kls [Ext])."""
    B, T, Tp = ctx.B, ctx.T, ctx.Tp
    _set_centre_only(ctx)
    nb = len(dec.decoder_residual_blocks)
    kls = []
    freeze_level = _take_freeze_level()
    prepare_all(ctx, dec)
    zs = latent_seq(ctx, dec.sequence_start[0], z, out_planes=_k(dec.decoder_blocks[0].module_list[0]._seq[0]),
                    name="decoder.start")
    out = None
    std_scale = 1e-10 if mode == "fix" else 1.0
    for i in range(nb):
        lastlvl = i == nb - 1
        res_seq = dec.decoder_residual_blocks[i].seq
        up = conv_block(ctx, dec.decoder_blocks[i].module_list[0]._seq[0], None, zs, K.ACT_GELU, want_f32=True,
                        out_planes=_k(res_seq[0]), transposed=True, name="decoder.up%d" % i)
        C = up.C
        cat_buf = None
        if not lastlvl:
            # [xs_sample ; decoder_out] of decoder.py:192 is a channel concat = adjacent rows in CR layout:
            # both producers write straight into the halves of one buffer (3 planes: its consumers are k3 convs)
            cat_planes = max(_k(dec.condition_z[i][0]._seq[0]), _k(dec.condition_xz[i][0]._seq[0]))
            cat_buf = ctx.op(cat_planes, 2 * C, B, Tp)
        out = cgg_seq(ctx, res_seq, up, res=up, res_scale=0.1, want_f32=True,
                      out_op_view=cat_buf[:, C:] if cat_buf is not None else None,
                      out_planes=_k(dec.recon[0]), name="decoder.level%d" % i)
        if lastlvl:
            break
        # condition_z: ResidualBlock -> GELU -> Conv k3 (decoder.py:150-157)
        cz_seq = dec.condition_z[i]
        r = cgg_seq(ctx, cz_seq[0]._seq, out, res=out, res_scale=0.1, post_gelu=True, out_planes=_k(cz_seq[2]))
        cz = conv_block(ctx, cz_seq[2], None, r, K.ACT_NONE)
        # xs path (decoder.py:189-193)
        xs_act = latent_seq(ctx, dec.xs_sequence[i], xs[i], out_op_view=cat_buf[:, :C], name="decoder.xs%d" % i)
        cat = Act(2 * C, data=cat_buf, name="cat%d" % i)
        if ctx.tape is not None:
            def cat_bwd(cat=cat, xs_act=xs_act, out=out, C=C):
                g = cat.grad
                if g is None:
                    return
                cat.grad = None
                xs_act.grad = g[:C]
                if out.grad is None:
                    out.grad = g[C:]
                else:
                    K.axpy(out.grad, g[C:], 1.0, True)
            ctx.tape.append(cat_bwd)
        cxz_seq = dec.condition_xz[i]
        r2 = cgg_seq(ctx, cxz_seq[0]._seq, cat, res=cat, res_scale=0.1, post_gelu=True, out_planes=_k(cxz_seq[2]))
        cxz = conv_block(ctx, cxz_seq[2], None, r2, K.ACT_NONE)
        # kl_2 + reparameterisation + "decoder_out + z" of the next level (decoder.py:179,193-212)
        eps = draw_eps((B, C, T), ctx.dev)
        zs_next = Act(C, data=ctx.op(_k(dec.decoder_blocks[i + 1].module_list[0]._seq[0]), C, B, Tp),
                      name="decoder.zs%d" % i)
        kl_sum = ctx.f64(1)
        out_f32 = out.as_f32()
        frozen = mode == "fix" and i < freeze_level
        zs_f32 = ctx.f32(C, B, Tp) if frozen else None
        K.kl2_reparam_fwd(cz.f32, cxz.f32, eps, out_f32, std_scale, zs_next.data, zs_f32, kl_sum, T)
        if frozen:
            _freeze_latent(ctx, dec, i, freeze_level, out_f32, zs_f32, zs_next)
        kl_t = ctx.f32(1)
        K.scale_f64_to_f32(kl_sum, kl_t, 0.5 / B)
        kl_ext = Ext(kl_t)
        kls.append(kl_ext)
        ctx.cap("decoder.zs%d" % i, zs_next)
        if ctx.tape is not None:
            def kl_bwd(cz=cz, cxz=cxz, eps=eps, zs_next=zs_next, kl_ext=kl_ext, out=out, C=C):
                dzs = zs_next.grad
                dkl = kl_ext.grad
                if dzs is None and dkl is None:
                    return
                zs_next.grad = None
                cz.grad = ctx.f32(2 * C, B, Tp)
                cxz.grad = ctx.f32(2 * C, B, Tp)
                K.kl2_reparam_bwd(cz.f32, cxz.f32, eps, std_scale, dzs, dkl, 0.5 / B, cz.grad, cxz.grad, T)
                if dzs is not None:
                    if out.grad is None:
                        out.grad = dzs
                    else:
                        K.axpy(out.grad, dzs, 1.0, True)
            ctx.tape.append(kl_bwd)
        zs = zs_next

    # reconstruction head: Conv k1 -> GroupNorm -> Tanh (+ losses)
    conv, gn = dec.recon[0], dec.recon[1]
    p = prep_conv(ctx, conv)
    finish_prepare(ctx)
    N, G = p.Cout, gn.num_groups
    # bf16 mode: the pre-norm output of the recon conv (the largest tensor of the step, read by the forward and the
    # backward head kernels) is stored as bf16; its GroupNorm statistics are taken from the fp32 accumulators
    y_16 = ctx.op_dtype != torch.float32 and N > 128
    # loss target: the packed operand of x (same layout and dtype as y) or the fp32 tensor
    x_op = _take_loss_operand()
    if static_compact(B, T) and p.k == 1 and x is not None and not isinstance(x, PackedBatch) \
            and K.conv_out16_ok(N) and out.data.shape[0] == 1:
        return _static_recon(ctx, conv, gn, p, out, x, x_op, lossfun, kls, want_xhat and _materialize_xhat())
    if isinstance(x_op, StaticTarget):
        x_op = None
    y = torch.empty(N, B, Tp, dtype=ctx.op_dtype if y_16 else torch.float32, device=ctx.dev)
    use_op = x is not None and x_op is not None and y_16 and loss_target(T) == "operand" and \
        tuple(x_op.shape) == (1, N, B, Tp) and x_op.dtype == ctx.op_dtype
    if isinstance(x, PackedBatch):
        if not use_op:
            raise RuntimeError("simulgen_b200: a PackedBatch needs loss_target() == 'operand' (fp16 mode, T % 8 == 0)")
    xt = x_op[0] if use_op else x
    stats = ctx.f32(B, G, 2)
    K.conv_fprop_gn(p.wg, out.data, conv.bias, y, p.Cin, stats, T, G)
    want_xhat = want_xhat and (_materialize_xhat() or x is None)
    x_hat = ctx.f32(B, N, T) if want_xhat else None
    res = dict(x_hat=x_hat, recon=None, mse=None, kls=kls)
    loss_kind = K.LOSS_KINDS.get(lossfun, 0)
    sums = ctx.f64(2)
    # training: the forward also takes the row sums of the GroupNorm backward (one backward pass instead of two)
    rowsums = ctx.f32(N * B, 4) if (ctx.tape is not None and x is not None) else None
    K.recon_fwd(y, stats, gn.weight, gn.bias, xt, x_hat, sums, T, G, loss_kind, rowsums)
    inv_numel = 1.0 / float(B * N * T)
    if x is not None:
        both = ctx.f32(2)
        K.scale_f64_to_f32(sums, both, inv_numel)
        res["recon"], res["mse"] = Ext(both[0:1]), Ext(both[1:2])
    xhat_ext = Ext(x_hat)
    res["x_hat_ext"] = xhat_ext
    if ctx.tape is not None:
        def recon_bwd(out=out):
            g_loss = res["recon"].grad if res["recon"] is not None else None
            g_mse = res["mse"].grad if res["mse"] is not None else None
            g_ext = xhat_ext.grad
            if g_loss is None and g_mse is None and g_ext is None:
                return
            dy = ctx.op(1, N, B, Tp)
            dgamma, dbeta, dbias = ctx.vec_grad(gn.weight, N), ctx.vec_grad(gn.bias, N), ctx.vec_grad(conv.bias, N)
            xb = xt
            if use_op and g_ext is not None:                # a gradient wrt x_hat itself: the two-pass kernels read fp32 x
                if isinstance(x, PackedBatch):
                    raise RuntimeError("simulgen_b200: a gradient wrt x_hat needs the fp32 batch, not a PackedBatch")
                xb = x
            K.recon_bwd(y, stats, gn.weight, gn.bias, xb, g_loss, g_mse, inv_numel, g_ext, dy, dgamma, dbeta, dbias,
                        T, G, loss_kind, rowsums if xb is xt else None)
            ctx.set_pgrad(gn.weight, dgamma)
            ctx.set_pgrad(gn.bias, dbeta)
            ctx.set_pgrad(conv.bias, dbias)
            dx, acc = ctx.grad_buf(out)
            K.conv_dgrad(p.wg, dy, dx, p.Cin, bool(acc))
            if p.w.requires_grad:
                with ctx.side(dy, out.data, flops=2.0 * p.Cin * p.Cout * p.k * B * T):
                    dwg = ctx.weight_grad_buf(p, p.k, p.Cout, p.Cin_p)
                    K.conv_wgrad(dy, out.data, dwg, p.Cin)
                    _weight_grad(ctx, conv, p, dwg)
        ctx.tape.append(recon_bwd)
    return res


# ------------------------------------------------------------------------------------------------
# direct training step (no torch.autograd): what Trainer's fused path runs
# ------------------------------------------------------------------------------------------------
def train_step_direct(model, x, alpha, beta, scale=None, hooks=None):
    """Forward + backward of `alpha * recon + beta * sum(kl)` (train.py:142-153) driven straight from the engine's own
    tapes: encoder graph -> main-latent reparameterisation -> decoder graph + losses, then the three backward passes in
    reverse, seeded with d loss / d recon = alpha and d loss / d kl_i = beta (times `scale`, the loss scale: a float or a
    1-element device tensor).  The autograd Functions below do the same through torch.autograd for callers that own the
    loop (the reference's train.py); here there is no autograd graph, no engine worker thread and no per-step Python
    object churn - which also makes the whole step capturable as ONE CUDA graph (Trainer.cuda_graph).
    Needs a gradient sink (set_grad_sink): parameter gradients go to its arenas.  Returns detached device scalars
    (loss, recon, kl_sum, mse).
    hooks (optional): object with before_decoder_forward() / after_decoder_backward(), called at those points of the step
    - the data-parallel trainer starts the exchange of the decoder's gradients there (Trainer._peer_*)."""
    if get_grad_sink() is None:
        raise RuntimeError("simulgen_b200: train_step_direct needs a gradient sink (Trainer installs one)")
    enc, dec = model.encoder, model.decoder
    _check_input(x, "input batch")
    dev = x.device
    with device_guard(dev), torch.no_grad():
        if not isinstance(x, PackedBatch):
            x = x.contiguous().float()
        B, T = x.shape[0], x.shape[2]
        ectx = Ctx(B, T, dev, True)
        last, xs = encoder_graph(ectx, enc, x)
        set_loss_operand(take_last_packed())
        eps0 = draw_eps((B, model.latent_dim), dev)
        z = torch.empty(B, last.tensor.shape[1] // 2, dtype=torch.float32, device=dev)
        kl_main = torch.empty(1, dtype=torch.float32, device=dev)
        K.reparam_main_fwd(last.tensor, eps0, z, kl_main)
        if hooks is not None:
            hooks.before_decoder_forward()
        dctx = Ctx(B, dec.num_time, dev, True)
        z_ext = Ext(z)
        n_levels = len(dec.decoder_residual_blocks) - 1
        if len(xs) < n_levels:
            raise RuntimeError("Decoder.forward needs xs with at least %d entries" % n_levels)
        lossfun = model.lossfun if model.lossfun in model.loss_functions else "MSE"
        res = decoder_graph(dctx, dec, z_ext, xs, x, lossfun, "random", want_xhat=False)
        recon, mse = res["recon"].tensor, res["mse"].tensor
        kl_sum = kl_main.clone()
        for k_ext in res["kls"]:
            kl_sum += k_ext.tensor
        loss = recon * alpha + kl_sum * beta
        # ---- backward ----
        if isinstance(scale, torch.Tensor):                  # live device scalar (dynamic loss scaler): no host read
            sc = scale.reshape(1).float()
            g_recon, g_kl = sc * float(alpha), sc * float(beta)
        else:
            sc = 1.0 if scale is None else float(scale)
            g_recon = torch.full((1,), sc * float(alpha), dtype=torch.float32, device=dev)
            g_kl = torch.full((1,), sc * float(beta), dtype=torch.float32, device=dev)
        res["recon"].grad = g_recon.contiguous()
        for k_ext in res["kls"]:
            k_ext.grad = g_kl.contiguous()
        dctx.run_backward()
        if hooks is not None:
            hooks.after_decoder_backward()
        dlast = torch.empty_like(last.tensor)
        K.reparam_main_bwd(last.tensor, eps0, z_ext.grad, g_kl.contiguous(), dlast)
        last.grad = dlast
        ectx.run_backward()
    return loss.reshape(()), recon.reshape(()), kl_sum.reshape(()), mse.reshape(())


# ------------------------------------------------------------------------------------------------
# autograd boundary
# ------------------------------------------------------------------------------------------------
def _contig_f32(g):
    if g is None:
        return None
    if g.dtype != torch.float32 or not g.is_contiguous():
        g = g.to(torch.float32).contiguous()
    return g


def _check_input(x, what):
    if isinstance(x, PackedBatch):
        x = x.operand
    if not x.is_cuda:
        raise RuntimeError("simulgen_b200: %s must be a CUDA tensor - the engine has no CPU fallback" % what)


def device_guard(dev):
    """The C ABI takes a stream, not a device: every launch goes to the CURRENT device's stream.  The autograd
    Functions (and Trainer.step) therefore make the tensors' device current for their duration, so a model living on
    cuda:1 works from a process whose current device is cuda:0."""
    if getattr(dev, "type", None) == "cuda":
        return torch.cuda.device(dev)
    return contextlib.nullcontext()


class EncoderFn(torch.autograd.Function):
    @staticmethod
    def forward(fctx, enc, capture, x, *params):
        _check_input(x, "encoder input")
        fctx.set_materialize_grads(False)
        with device_guard(x.device):
            if not isinstance(x, PackedBatch):
                x = x.contiguous().float()
            record = _take_grad_mode() and any(fctx.needs_input_grad)
            ctx = Ctx(x.shape[0], x.shape[2], x.device, record, capture)
            last, xs = encoder_graph(ctx, enc, x)
        fctx.ectx, fctx.exts, fctx.params = ctx, [last] + xs, params
        return (last.tensor,) + tuple(e.tensor for e in xs)

    @staticmethod
    def backward(fctx, *grads):
        ctx = fctx.ectx
        with device_guard(ctx.dev):
            for e, g in zip(fctx.exts, grads):
                e.grad = _contig_f32(g)
            ctx.run_backward()
        out = tuple(ctx.pgrads.get(id(p)) for p in fctx.params)
        fctx.ectx = None
        return (None, None, None) + out


class ReparamMainFn(torch.autograd.Function):
    """VAE_network.py:103-105,113: clamp, exp, reparameterize, kl - one kernel each way."""

    @staticmethod
    def forward(fctx, last, eps):
        _check_input(last, "latent head output")
        fctx.set_materialize_grads(False)
        last = last.contiguous()
        B, L2 = last.shape
        z = torch.empty(B, L2 // 2, dtype=torch.float32, device=last.device)
        kl = torch.empty(1, dtype=torch.float32, device=last.device)
        with device_guard(last.device):
            K.reparam_main_fwd(last, eps, z, kl)
        fctx.save_for_backward(last, eps)
        return z, kl.view(())

    @staticmethod
    def backward(fctx, dz, dkl):
        last, eps = fctx.saved_tensors
        dlast = torch.empty_like(last)
        with device_guard(last.device):
            K.reparam_main_bwd(last, eps, _contig_f32(dz), _contig_f32(dkl.reshape(1)) if dkl is not None else None, dlast)
        return dlast, None


class DecoderFn(torch.autograd.Function):
    @staticmethod
    def forward(fctx, dec, capture, lossfun, mode, n_xs, z, *rest):
        xs_t = rest[:n_xs]
        x = rest[n_xs]
        params = rest[n_xs + 1:]
        _check_input(z, "decoder latent")
        fctx.set_materialize_grads(False)
        with device_guard(z.device):
            z = z.contiguous().float()
            record = _take_grad_mode() and any(fctx.needs_input_grad)
            ctx = Ctx(z.shape[0], dec.num_time, z.device, record, capture)
            z_ext = Ext(z)
            xs_ext = [Ext(t.contiguous().float()) for t in xs_t]
            if x is not None and not isinstance(x, PackedBatch):
                x = x.contiguous().float()
            res = decoder_graph(ctx, dec, z_ext, xs_ext, x, lossfun, mode)
        fctx.ectx, fctx.res, fctx.params, fctx.z_ext, fctx.xs_ext = ctx, res, params, z_ext, xs_ext
        fctx.has_x = x is not None
        outs = [res["x_hat"]]
        if x is not None:
            outs += [res["recon"].tensor.view(()), res["mse"].tensor.view(())]
        outs += [k.tensor.view(()) for k in res["kls"]]
        return tuple(outs)

    @staticmethod
    def backward(fctx, *grads):
        ctx, res = fctx.ectx, fctx.res
        grads = list(grads)
        with device_guard(ctx.dev):
            res["x_hat_ext"].grad = _contig_f32(grads.pop(0))
            if fctx.has_x:
                g = grads.pop(0)
                res["recon"].grad = _contig_f32(g.reshape(1)) if g is not None else None
                g = grads.pop(0)
                res["mse"].grad = _contig_f32(g.reshape(1)) if g is not None else None
            for k_ext in res["kls"]:
                g = grads.pop(0)
                k_ext.grad = _contig_f32(g.reshape(1)) if g is not None else None
            ctx.run_backward()
        pg = tuple(ctx.pgrads.get(id(p)) for p in fctx.params)
        gz = fctx.z_ext.grad
        gxs = tuple(e.grad for e in fctx.xs_ext)
        fctx.ectx = None
        fctx.res = None
        return (None, None, None, None, None, gz) + gxs + (None,) + pg

