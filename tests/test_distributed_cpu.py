"""CPU, world_size 2 over gloo: the multi-GPU protocol (shard -> partial -> all-gather -> mod-q
add) reproduces the single-process result bit for bit.  The mod-q add runs on the oracle here
(the CUDA kernel needs a GPU; the GPU suite covers it through ckks_add_many)."""
import importlib
import os
import sys

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = "seal-fyp-logistic-regression_b200"


def _worker(rank, world, port, ret):
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        from oracle import pyoracle as po
        par = importlib.import_module(PKG + ".parallel")
        log_n = 12
        primes = po.coeff_modulus_create(log_n, [50, 40, 50])
        orc = po.Oracle(log_n, primes)
        rng = np.random.default_rng(7)          # same stream on every rank
        d, L = 9, 2
        cts = np.stack([np.stack([rng.integers(0, p, size=(2, orc.n), dtype=np.uint64) for p in primes[:L]], axis=1)
                        for _ in range(d)])
        mine = par.shard_units(d, rank, world)
        assert mine == list(range(rank, d, world))
        part = cts[mine[0]]
        for u in mine[1:]:
            part = orc.add(part, cts[u])
        g = par.gather_partials(torch.from_numpy(part.view(np.int64).copy()))
        assert g.shape[0] == world
        g = g.numpy().view(np.uint64)
        total = g[0]
        for r in range(1, world):
            total = orc.add(total, g[r])
        want = cts[0]
        for u in range(1, d):
            want = orc.add(want, cts[u])
        ok = bool(np.array_equal(total, want))
        # a non-modular integer sum (what ncclSum would do) must differ somewhere
        naive = (g[0] + g[1])
        ok_naive_differs = bool(not np.array_equal(naive, want))
        ret[rank] = (ok, ok_naive_differs, len(mine))
    finally:
        dist.destroy_process_group()


def test_sharded_sum_matches_sequential_gloo():
    world = 2
    mgr = mp.Manager()
    ret = mgr.dict()
    port = 29500 + (os.getpid() % 2000)
    mp.spawn(_worker, args=(world, port, ret), nprocs=world, join=True)
    assert len(ret) == world
    for r in range(world):
        ok, naive_differs, n = ret[r]
        assert ok and naive_differs
    assert sum(ret[r][2] for r in range(world)) == 9


def test_shard_units_partition():
    par = importlib.import_module(PKG + ".parallel")
    for n in (1, 7, 64, 128):
        for G in (1, 2, 4, 8):
            parts = [par.shard_units(n, r, G) for r in range(G)]
            assert sorted(sum(parts, [])) == list(range(n))
            assert max(len(p) for p in parts) - min(len(p) for p in parts) <= 1


def test_weighted_shards_balance_naf_cost():
    """Linear_Transform diagonals are split by key-switch cost (NAF weight), not by index"""
    par = importlib.import_module(PKG + ".parallel")
    from oracle import pyoracle as po
    for l in list(range(-130, 130)) + [4095, -8192]:
        assert par.naf_weight(l) == len(po.naf(l)), l
    for d in (64, 128):
        w = [par.naf_weight(l) for l in range(d)]
        for G in (1, 2, 4, 8):
            parts = [par.shard_units_weighted(w, r, G) for r in range(G)]
            assert sorted(sum(parts, [])) == list(range(d))                     # a partition
            loads = [sum(w[u] for u in p) for p in parts]
            assert max(loads) - min(loads) <= 1, (d, G, loads)                  # l mod G: 156 vs 199 at d=128, G=2
            assert all(p == sorted(p) for p in parts)


def test_prefix_aware_shards_for_shared_rotation_plans():
    """rotations of one ciphertext share their common NAF prefixes (ckks_rotplan_keyswitches_shared): the sharding keeps
    rotations with a common prefix on one rank and balances the resulting prefix-tree sizes"""
    par = importlib.import_module(PKG + ".parallel")
    from oracle import pyoracle as po

    def tree(steps):
        nodes = set()
        for st in steps:
            t = par.naf_terms(st)
            nodes.update(tuple(t[:k]) for k in range(1, len(t) + 1))
        return len(nodes)

    for l in list(range(-130, 130)) + [4095, -8192]:
        assert par.naf_terms(l) == list(po.naf(l)), l
    assert tree(range(128)) == 169 and sum(par.naf_weight(l) for l in range(128)) == 355
    for d, want in ((64, {1: 84, 2: 43, 4: 22, 8: 12}), (128, {1: 169, 2: 85, 4: 43, 8: 22})):
        w = [par.naf_weight(l) for l in range(d)]
        for G in (1, 2, 4, 8):
            parts = [par.shard_rotations_shared(list(range(d)), r, G) for r in range(G)]
            assert sorted(sum(parts, [])) == list(range(d))
            assert all(p == sorted(p) for p in parts)
            assert max(tree(p) for p in parts) == want[G], (d, G, [tree(p) for p in parts])
            # never worse than the cost-balanced split that ignores prefixes
            assert max(tree(p) for p in parts) <= max(tree(par.shard_units_weighted(w, r, G)) for r in range(G))
    # arbitrary step lists, more ranks than rotations
    parts = [par.shard_rotations_shared([5, -3, 21], r, 4) for r in range(4)]
    assert sorted(sum(parts, [])) == [0, 1, 2]


def test_strong_scaling_units_partition_the_problem():
    """bench.py --scaling strong: the 32 (mini-batch, feature) gradient chains of one 8 x 32768 problem"""
    import importlib.util
    spec = importlib.util.spec_from_file_location("bench", os.path.join(ROOT, "bench.py"))
    bench = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(bench)
    M, C = bench.R_PER_GPU // bench.B_MINI, bench.C_FEAT
    for G in (1, 2, 4, 8, 16, 32):
        parts = [bench.strong_units(r, G) for r in range(G)]
        assert sorted(sum(parts, [])) == [(m, j) for m in range(M) for j in range(C)]
        assert len({len(p) for p in parts}) == 1
        if G <= M:
            assert all(len({m for m, _ in p}) == M // G for p in parts)        # whole mini-batches per GPU
    with pytest.raises(SystemExit):
        bench.strong_units(0, 3)
