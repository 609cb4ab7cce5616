"""Multi-GPU sharding of the hot path: independent ciphertexts are partitioned
across ranks (one process per GPU, torch.distributed), every rank computes its partial result
with no collective in the data path, and the partial ciphertexts are combined with one
all-gather followed by the mod-q add kernel (NCCL's sum is not modular).  Modular addition is
associative and commutative, so the combined ciphertext is bit-identical to the sequential
add_many of the reference (helper.h:259, logistic_regression_ckks.cpp:316)."""
import torch
import torch.distributed as dist

from .engine import Ciphertext


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_units(n_units, rank, world_size):
    """interleaved partition: unit u belongs to rank u mod G.  For the diagonals of a linear
    transform this balances the NAF weight (key switches) of the rotations far better than
    contiguous blocks."""
    return list(range(rank, n_units, world_size))


def naf_weight(steps):
    """number of key switches SEAL's rotate_vector spends on `steps` with the default power-of-two Galois
    keys: the number of non-zero digits of its non-adjacent form (util::naf)"""
    v, w = abs(int(steps)), 0
    while v:
        if v & 1:
            z = 2 - (v & 3)
            v -= z
            w += 1
        v >>= 1
    return w


def shard_units_weighted(weights, rank, world_size):
    """partition units by COST instead of index: longest-processing-time greedy (heaviest unit first, always
    onto the least-loaded rank; ties broken by rank index so every rank computes the same assignment).
    For the diagonals of a linear transform the cost of diagonal l is naf_weight(l) key switches: `l mod G`
    gives rank 0 all even diagonals (157 key switches at d = 128, G = 2) and rank 1 all odd ones (199);
    this split is 178 / 178.  Returns this rank's unit indices in increasing order."""
    order = sorted(range(len(weights)), key=lambda u: (-weights[u], u))
    load = [0] * world_size
    mine = []
    for u in order:
        r = min(range(world_size), key=lambda k: (load[k], k))
        load[r] += weights[u]
        if r == rank:
            mine.append(u)
    return sorted(mine)


def naf_terms(steps):
    """the signed power-of-two terms SEAL's rotate_vector applies for `steps`, in its order (least significant
    first; util::naf)"""
    v, i, out = abs(int(steps)), 0, []
    sign = 1 if steps >= 0 else -1
    while v:
        if v & 1:
            z = 2 - (v & 3)
            v -= z
            out.append(sign * z * (1 << i))
        v >>= 1
        i += 1
    return out


def shard_rotations_shared(steps, rank, world_size):
    """partition the rotations of ONE ciphertext for rotation plans that share common NAF prefixes
    (ckks_rotplan_keyswitches_shared): rotations are ordered by their term sequences, so that rotations with a common
    prefix are neighbours, and the order is cut into world_size contiguous runs of equal cost, where a rotation costs the
    terms it does not share with its predecessor in the run (= the key switches the run's prefix tree adds for it).
    d = 128: 169 key switches on 1 GPU, 85 / 85 on 2, 4 x 43 on 4, 8 x 22 on 8 (cost-balanced without regard to prefixes: 90 / 51 / 31).
    Returns this rank's indices into `steps`, increasing."""
    seqs = [tuple(naf_terms(st)) for st in steps]
    order = sorted(range(len(steps)), key=lambda u: (seqs[u], u))

    def lcp(a, b):
        n = 0
        while n < len(a) and n < len(b) and a[n] == b[n]:
            n += 1
        return n

    def cuts(limit):
        """greedy: fewest runs with cost <= limit each; returns the list of runs"""
        runs, cur, cost = [], [], 0
        for u in order:
            add = len(seqs[u]) - (lcp(seqs[cur[-1]], seqs[u]) if cur else 0)
            if cur and cost + add > limit:
                runs.append(cur)
                cur, cost = [], 0
                add = len(seqs[u])
            cur.append(u)
            cost += add
        if cur:
            runs.append(cur)
        return runs

    lo, hi = max([len(q) for q in seqs] + [1]), sum(len(q) for q in seqs) + 1
    while lo < hi:                       # smallest per-run cost that needs at most world_size runs
        mid = (lo + hi) // 2
        if len(cuts(mid)) <= world_size:
            hi = mid
        else:
            lo = mid + 1
    runs = cuts(lo)
    runs += [[] for _ in range(world_size - len(runs))]
    return sorted(runs[rank])


def gather_partials(t):
    """all-gather one tensor per rank -> [G, ...] on every rank (any backend)"""
    rank, G = world()
    if G == 1:
        return t.unsqueeze(0) if t.dim() == 3 else t
    flat = t.reshape((1,) + tuple(t.shape[-3:])) if t.dim() == 3 else t
    out = torch.empty((G * flat.shape[0],) + tuple(flat.shape[1:]), dtype=flat.dtype, device=flat.device)
    dist.all_gather_into_tensor(out, flat.contiguous())
    return out


def combine_partials(ev, partial):
    """partial: batch-1 Ciphertext held by every rank -> sum over ranks (mod q) on every rank"""
    rank, G = world()
    if G == 1:
        return partial
    gathered = gather_partials(partial.data)
    return ev.add_many(Ciphertext(partial.ctx, gathered, partial.limbs, partial.scale))


def sharded_linear_transform_plain(ev, ct_new_rots_fn, diags_local, d, plans_steps):
    """helper for a sharded Linear_Transform_Plain: this rank owns the diagonals `plans_steps`
    (a subset of range(d)); returns the combined ciphertext"""
    rots = ct_new_rots_fn(plans_steps)
    part = ev.multiply_plain_sum(rots, diags_local)
    return combine_partials(ev, part)
