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
