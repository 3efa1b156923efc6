"""Sharded optimiser over peer memory (data parallel without a gradient all-reduce: sg_peer_reduce_dot + sg_opt_step with
an sg_peer context, trainer.shard_item) with W VIRTUAL ranks inside one process: every "rank" has its own gradient
arenas and flat parameter buffer, holds pointers to all of them, reduces / updates only its shard and stores the new
parameters everywhere.  After one round every rank's parameters must equal the plain multi-tensor step on the summed
gradients.  CPU: the torch models of the two kernels (host-side shard arithmetic); GPU: the real kernels - on one device
the "peer" pointers are ordinary device pointers, the NVLink case differs only in where they point."""
import pytest
import torch

import kernel_emulator as emu
from conftest import rel_l2
from simulgen_vae_b200 import kernels as RK          # make_scaler_state / read_scaler_state are plain torch helpers
from simulgen_vae_b200.trainer import shard_item

SPECS = [("conv", 40, 24, 3), ("convT", 40, 24, 3), ("linear", 16, 64, 1), ("conv", 24, 20, 1), ("conv", 300, 40, 5),
         ("vec", 10007, 0, 0), ("vec", 4096, 0, 0), ("linear", 8, 12000, 1), ("vec", 3, 0, 0), ("conv", 5, 8, 1)]


def _rnd(*shape, seed, dev, scale=1.0):
    g = torch.Generator().manual_seed(seed + sum(shape))
    return (torch.randn(*shape, generator=g) * scale).to(dev)


def _layout():
    """offsets of every item's parameter, weight-arena gradient and vector-arena gradient"""
    offs, np_, nw, nv = [], 0, 0, 0
    for kind, a, b, k in SPECS:
        if kind == "vec":
            offs.append((np_, None, nv, a, a))
            np_ += (a + 63) // 64 * 64
            nv += (a + 63) // 64 * 64
        else:
            Cout, Cin = a, b
            Cin_p = Cin if kind == "linear" else (Cin + 7) // 8 * 8
            n, ng = Cout * Cin * k, k * Cout * Cin_p
            offs.append((np_, nw, None, n, ng))
            np_ += (n + 63) // 64 * 64
            nw += (ng + 63) // 64 * 64
    return offs, np_, nw, nv


def _make_rank(dev, seed_g):
    """one (virtual) rank: flat parameters (identical on all ranks), arenas with this rank's own gradients, full items"""
    offs, np_, nw, nv = _layout()
    flat = torch.zeros(np_, device=dev)
    weights, vecs = torch.zeros(max(nw, 1), device=dev), torch.zeros(max(nv, 1), device=dev)
    items = []
    for si, ((kind, a, b, k), (po, wo, vo, n, ng)) in enumerate(zip(SPECS, offs)):
        if kind == "vec":
            p = flat[po:po + n]
            p.copy_(_rnd(n, seed=si, dev=dev))
            g = vecs[vo:vo + n]
            g.copy_(_rnd(n, seed=seed_g + si, dev=dev))
            items.append(dict(p=p, g=g, vec_arena=True))
            continue
        Cout, Cin = a, b
        Cin_p = Cin if kind == "linear" else (Cin + 7) // 8 * 8
        shape = (Cout, Cin) if kind == "linear" else ((Cin, Cout, k) if kind == "convT" else (Cout, Cin, k))
        p = flat[po:po + n].view(shape)
        p.copy_(_rnd(*shape, seed=si, dev=dev))
        g = weights[wo:wo + ng].view(k, Cout, Cin_p)
        g.copy_(_rnd(k, Cout, Cin_p, seed=seed_g + si, dev=dev))
        items.append(dict(p=p, g=g, vec_arena=False,
                          u=torch.nn.functional.normalize(_rnd(Cout, seed=400 + si, dev=dev), dim=0),
                          vv=torch.nn.functional.normalize(_rnd(Cin * k, seed=500 + si, dev=dev), dim=0),
                          sigma=torch.tensor([1.3 + 0.1 * si], device=dev), Cout=Cout, Cin=Cin, Cin_p=Cin_p, k=k,
                          flip=int(kind == "convT")))
    return dict(flat=flat, weights=weights, vecs=vecs, items=items)


def _run(K, dev, W, with_scaler, split=False):
    ranks = [_make_rank(dev, 1000 * (r + 1)) for r in range(W)]
    # reference: the plain multi-tensor step (torch model) on the sum of all ranks' gradients
    ref = _make_rank(dev, 0)
    for i, it in enumerate(ref["items"]):
        it["g"].copy_(sum(rk["items"][i]["g"].double() for rk in ranks).float())
        it.update(m=torch.zeros_like(it["p"]), v=torch.zeros_like(it["p"]))
    gref = torch.zeros(1, device=dev, dtype=torch.float64)
    sref = RK.make_scaler_state(dev, 8.0) if with_scaler else None
    emu.opt_step(emu.OptPlan(ref["items"], dev), 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 1.0 / W, gref, sref)
    # W virtual ranks
    plans, peers, gn, scalers = [], [], [], []
    for r, rk in enumerate(ranks):
        sh = []
        for it in rk["items"]:
            s = shard_item(it, r, W)
            if s is not None:
                s.update(m=torch.zeros(s["n"], device=dev), v=torch.zeros(s["n"], device=dev))
                sh.append(s)
        rk["shards"] = sh
        if split:
            # two item tables over ONE scratch buffer, as the trainer splits decoder / encoder (Trainer._build_plan)
            n_sn = sum(1 for it in sh if it.get("u") is not None)
            dots = torch.zeros(n_sn + 6, dtype=torch.float64, device=dev)
            first, second = sh[:len(sh) // 2], sh[len(sh) // 2:]
            pa = K.OptPlan(first, dev, dots=dots, dot_base=0)
            pb = K.OptPlan(second, dev, dots=dots, dot_base=sum(1 for it in first if it.get("u") is not None))
            plans.append((pa, pb))
        else:
            plans.append(K.OptPlan(sh, dev))
        peers.append(K.make_peer(r, [q["weights"] for q in ranks], [q["vecs"] for q in ranks], [q["flat"] for q in ranks]))
        gn.append(torch.zeros(1, device=dev, dtype=torch.float64))
        scalers.append(RK.make_scaler_state(dev, 8.0) if with_scaler else None)
    for r in range(W):                                                   # (barrier) every rank reduces its shard
        if split:
            K.peer_reduce_dot(plans[r][0], with_scaler, peers[r], clear_dots=True, max_blocks=7)     # grid-stride path
            K.peer_reduce_dot(plans[r][1], with_scaler, peers[r], clear_dots=False)
        else:
            K.peer_reduce_dot(plans[r], with_scaler, peers[r])
    dots_of = (lambda r: plans[r][0].dots) if split else (lambda r: plans[r].dots)
    # the step's only collective: all-reduce of the per-layer dots (+ overflow flag); shards see different subsets of
    # the spectral-norm layers only when a layer has fewer rows than ranks, so reduce by layer identity
    totals = {}
    for r in range(W):
        for it in ranks[r]["shards"]:
            if it.get("u") is not None:
                key = it["full"]["p"].data_ptr() - ranks[r]["flat"].data_ptr()
                totals[key] = totals.get(key, 0.0) + float(dots_of(r)[it["dot_index"]])
    for r in range(W):
        for it in ranks[r]["shards"]:
            if it.get("u") is not None:
                dots_of(r)[it["dot_index"]] = totals[it["full"]["p"].data_ptr() - ranks[r]["flat"].data_ptr()]
    for r in range(W):
        if split:      # scalars once (phase 2, second table), then the first table with the same scalars (phase 3)
            K.opt_step(plans[r][1], 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 1.0 / W, gn[r], scalers[r], peer=peers[r], phase=2)
            K.opt_step(plans[r][0], 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 1.0 / W, gn[r], scalers[r], peer=peers[r], phase=3,
                       max_blocks=5)
        else:
            K.opt_step(plans[r], 1e-3, 0.9, 0.999, 1e-8, 0.01, 1, 1.0 / W, gn[r], scalers[r], peer=peers[r], phase=2)
    if dev != "cpu":
        torch.cuda.synchronize()
    for r in range(W):
        assert rel_l2(ranks[r]["flat"], ref["flat"]) < 3e-6, (r, rel_l2(ranks[r]["flat"], ref["flat"]))
        assert torch.equal(ranks[r]["flat"], ranks[0]["flat"])           # replicas stay bit-identical
    total_gn = sum(float(g) for g in gn)
    assert abs(total_gn - float(gref)) / float(gref) < 1e-5
    # moments: each rank holds only its shard
    for r in range(W):
        for it in ranks[r]["shards"]:
            i = [id(x) for x in ranks[r]["items"]].index(id(it["full"]))
            lo, hi = it["rows"]
            full_m = ref["items"][i]["m"]
            want = full_m.reshape(full_m.shape[0], -1)[lo:hi].reshape(-1) if it.get("u") is not None else full_m.reshape(-1)[lo:hi]
            assert rel_l2(it["m"], want) < 2e-5
    if with_scaler:
        st = RK.read_scaler_state(scalers[0])
        assert st["step"] == 1 and st["skipped"] == 0


def test_shard_item_partitions_every_tensor():
    rk = _make_rank("cpu", 7)
    for W in (2, 3, 8):
        for it in rk["items"]:
            n = it["p"].numel()
            covered = 0
            for r in range(W):
                s = shard_item(it, r, W)
                if s is None:
                    continue
                off = (s["p"].data_ptr() - it["p"].data_ptr()) // 4
                assert off == covered, "shards must tile the tensor in order"
                covered += s["n"]
            assert covered == n


@pytest.mark.parametrize("W", [2, 3, 4, 8])
@pytest.mark.parametrize("with_scaler,split", [(False, False), (True, False), (True, True)])
def test_peer_optimizer_virtual_ranks_cpu(W, with_scaler, split):
    _run(emu, "cpu", W, with_scaler, split)


@pytest.mark.gpu
@pytest.mark.parametrize("W", [2, 3, 4, 8])
@pytest.mark.parametrize("with_scaler,split", [(False, False), (True, False), (True, True), (False, True)])
def test_peer_optimizer_virtual_ranks_gpu(W, with_scaler, split):
    from simulgen_vae_b200 import kernels as K
    _run(K, "cuda", W, with_scaler, split)
