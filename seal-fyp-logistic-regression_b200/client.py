"""Client-side objects the reference builds around the Evaluator: CKKSEncoder, KeyGenerator,
Encryptor, Decryptor (reference: logistic_regression_ckks.cpp:426-441, helper.h callers).

These sit *beside* the hot path.  All ring arithmetic runs on the
GPU through the C ABI (NTT, dyadic products, divide-by-last-prime); the host only samples
randomness (numpy) and does the floating-point embedding / CRT composition.  Results are
compared with tolerance, never bit-exactly (SEAL's own randomness is unpinned).
"""
import os

import numpy as np
import torch

from .engine import Ciphertext, Evaluator, KeySet


def _bitrev(x, bits):
    x = np.asarray(x, dtype=np.uint64)
    r = np.zeros_like(x)
    for _ in range(bits):
        r = (r << np.uint64(1)) | (x & np.uint64(1))
        x = x >> np.uint64(1)
    return r


def galois_perm(log_n, g):
    """index table of the NTT-domain automorphism: out[i] = in[perm[i]]"""
    n = 1 << log_n
    i = np.arange(n, dtype=np.uint64)
    e = (np.uint64(g) * (np.uint64(2) * _bitrev(i, log_n) + np.uint64(1))) & np.uint64(2 * n - 1)
    return _bitrev((e - np.uint64(1)) >> np.uint64(1), log_n).astype(np.int64)


class CKKSEncoder:
    """SEAL CKKSEncoder: slot i <-> evaluation at zeta^(3^i), zeta = exp(2 pi i / 2N)."""

    def __init__(self, ctx):
        self.ctx = ctx
        self.ev = Evaluator(ctx)
        n = ctx.n
        self.slots = n // 2
        pos = np.empty(self.slots, dtype=np.int64)
        v = 1
        for i in range(self.slots):
            pos[i] = v
            v = v * 3 % (2 * n)
        self._k = (pos - 1) // 2                  # exponent 2k+1 = 3^i
        self._kc = (2 * n - pos - 1) // 2         # conjugate position
        self._twist = np.exp(1j * np.pi * np.arange(n) / n)   # zeta^j

    def slot_count(self):
        return self.slots

    def _to_residues(self, coeffs, limbs):
        """integer-valued float64 coefficients -> [limbs][N] uint64 residues"""
        out = np.empty((limbs, self.ctx.n), dtype=np.uint64)
        if np.abs(coeffs).max() >= 2.0 ** 62:
            ints = [int(c) for c in coeffs]
            for j in range(limbs):
                p = self.ctx.primes[j]
                out[j] = np.array([c % p for c in ints], dtype=np.uint64)
        else:
            ints = coeffs.astype(np.int64)
            for j in range(limbs):
                out[j] = np.mod(ints, np.int64(self.ctx.primes[j])).astype(np.uint64)
        return out

    def encode(self, values, scale, limbs=None):
        """encode(vector<double>) or encode(double) -> Plaintext (NTT form, on device): the
        embedding DFT, rounding, RNS reduction and NTT all run on the GPU (ckks_encode).
        A 2-D `values` array (numpy or CUDA tensor) gives one plaintext per row (a batch)."""
        if np.isscalar(values):
            return self.ev.encode_scalar(float(values), scale, limbs)
        if not torch.is_tensor(values):
            values = torch.from_numpy(np.atleast_2d(np.asarray(values, dtype=np.float64)))
        if values.dim() == 2 and values.shape[1] > self.slots:
            values = values[:, : self.slots]
        return self.ev.encode(values, scale, limbs)

    def decode(self, pt):
        """Plaintext (batch B) -> float64 [B][slots] (numpy), computed on the GPU (ckks_decode)"""
        return self.ev.decode(pt).cpu().numpy()

    def encode_host(self, values, scale, limbs=None):
        """numpy restatement of encode (host FFT + device NTT); kept as a cross-check of the device path"""
        limbs = self.ctx.top_limbs if limbs is None else limbs
        n = self.ctx.n
        if np.isscalar(values):
            ints = int(round(float(values) * scale))
            res = np.empty((1, 1, limbs, n), dtype=np.uint64)
            for j in range(limbs):
                res[0, 0, j, :] = ints % self.ctx.primes[j]    # NTT of a constant is that constant
            return self.ctx.upload(res, scale=scale)
        vals = np.atleast_2d(np.asarray(values, dtype=np.float64))
        B = vals.shape[0]
        m = min(vals.shape[1], self.slots)
        res = np.empty((B, 1, limbs, n), dtype=np.uint64)
        for b in range(B):
            v = np.zeros(n, dtype=np.complex128)
            v[self._k[:m]] = vals[b, :m]
            v[self._kc[:m]] = vals[b, :m]
            coeffs = np.real(np.fft.fft(v) / n * np.conj(self._twist))
            res[b, 0] = self._to_residues(np.round(coeffs * scale), limbs)
        pt = self.ctx.upload(res, scale=scale)
        self.ev.ntt_forward(pt.data.view(B, limbs, n))
        return pt

    def decode_host(self, pt):
        """numpy restatement of decode (device INTT + host CRT/FFT); cross-check of the device path"""
        n, L = self.ctx.n, pt.limbs
        t = pt.data[:, 0, :L, :].clone()          # the inverse NTT runs in place: keep the plaintext intact
        self.ev.ntt_inverse(t)
        res = t.cpu().numpy().view(np.uint64)
        primes = self.ctx.primes[:L]
        Q = 1
        for p in primes:
            Q *= p
        out = np.empty((pt.batch, self.slots))
        for b in range(pt.batch):
            if L == 1:
                c = res[b, 0].astype(np.int64)
                c = np.where(c > primes[0] // 2, c - primes[0], c).astype(np.float64)
            else:
                acc = np.zeros(n, dtype=object)
                for j, p in enumerate(primes):
                    Qj = Q // p
                    acc = acc + res[b, j].astype(object) * (Qj * pow(Qj % p, -1, p))
                acc = acc % Q
                c = np.array([float(x - Q) if x > Q // 2 else float(x) for x in acc])
            v = n * np.fft.ifft(c / pt.scale * self._twist)
            out[b] = np.real(v[self._k])
        return out


class _RingOps:
    """device ring helpers shared by key generation / encryption / decryption"""

    def __init__(self, ctx, seed=None):
        """seed=None (the default): a fresh 256-bit ChaCha20 key from the operating system's CSPRNG (os.urandom) --
        what real keys and encryptions must use.  An integer seed gives a reproducible stream for tests and
        benchmarks (at most 64 bits of entropy: never for real keys)."""
        self.ctx = ctx
        self.ev = Evaluator(ctx)
        self._seed = os.urandom(32) if seed is None else int(seed) & (2 ** 64 - 1)
        self._stream = 0

    def _draw(self, kind, count, limbs):
        """sampling runs on the device (ckks_sample_keyed / ckks_sample): ChaCha20 in counter mode keyed by this
        object's key, one stream id (nonce) per call"""
        self._stream += 1
        return self.ev.sample(kind, self._seed, self._stream, count, limbs)

    def ternary_ntt(self, count, limbs):
        """count ternary polynomials, the same small integers in every limb, NTT form: [count][limbs][N]"""
        return self._draw(Evaluator.TERNARY, count, limbs)

    def errors_ntt(self, count, limbs):
        """rounded normal, sigma 3.2, clipped at 6 sigma (SEAL sample_poly_normal), NTT form"""
        return self._draw(Evaluator.NORMAL, count, limbs)

    def uniform(self, count, limbs):
        return self._draw(Evaluator.UNIFORM, count, limbs)

    def polys(self, t, limbs):
        """[P][limbs][N] tensor -> size-1 Ciphertext batch over primes [0, limbs)"""
        return Ciphertext(self.ctx, t[:, None].contiguous(), limbs)

    def mul(self, a, s):
        """a [P][W][N] times s [W][N] limb-wise (multiply_plain kernel, broadcast plaintext)"""
        W = a.shape[1]
        return self.ev.multiply_plain(self.polys(a, W), self.polys(s[None], W)).data[:, 0]

    def enc_zero_sym(self, count, sk, limbs):
        """count x (-(a s + e), a) over primes [0, limbs): tensor [count][2][limbs][N]"""
        a = self.uniform(count, limbs)
        e = self.errors_ntt(count, limbs)
        ase = self.ev.add(self.polys(self.mul(a, sk[:limbs]), limbs), self.polys(e, limbs))
        self.ev.negate_inplace(ase)
        return torch.stack([ase.data[:, 0], a], dim=1).contiguous()


class KeyGenerator:
    """SEAL KeyGenerator: ternary secret, public key, relinearisation and Galois keys
    (SURVEY.md A.5), at the key level (all K primes)."""

    def __init__(self, ctx, seed=None):
        """seed=None draws the generator key from os.urandom; an int makes the keys reproducible (tests, bench)"""
        self.ctx = ctx
        self.ops = _RingOps(ctx, seed)
        self._sk = self.ops.ternary_ntt(1, ctx.K)[0]      # [K][N]

    def secret_key(self):
        return self._sk

    def public_key(self):
        return self.ops.enc_zero_sym(1, self._sk, self.ctx.K)[0]             # [2][K][N]

    def _kswitch_key(self, new_key):
        """generate_one_kswitch_key: digit i is an encryption of zero with (P mod q_i) * new_key
        added into limb i of component 0.  new_key: [K][N], NTT form."""
        ctx, K, ops = self.ctx, self.ctx.K, self.ops
        key = ops.enc_zero_sym(K - 1, self._sk, K)                           # [K-1][2][K][N]
        P = ctx.primes[K - 1]
        fac = torch.empty((K, ctx.n), dtype=torch.int64, device=ctx.device)
        for j in range(K):
            fac[j] = P % ctx.primes[j]
        scaled = ops.mul(new_key[None], fac)[0]                              # limb j times (P mod q_j)
        for i in range(K - 1):
            p = ctx.primes[i]
            s = key[i, 0, i] + scaled[i]
            key[i, 0, i] = torch.where(s >= p, s - p, s)
        return key

    def relin_keys(self):
        return self._kswitch_key(self.ops.mul(self._sk[None], self._sk)[0])

    def galois_key(self, galois_elt):
        perm = torch.from_numpy(galois_perm(self.ctx.log_n, galois_elt)).to(self.ctx.device)
        return self._kswitch_key(self._sk[:, perm].contiguous())

    def default_galois_elts(self):
        """KeyGenerator::galois_keys(): steps +-2^i and the conjugation"""
        elts = {2 * self.ctx.n - 1}
        for i in range(self.ctx.log_n - 1):
            elts.add(self.ctx.galois_elt(1 << i))
            elts.add(self.ctx.galois_elt(-(1 << i)))
        return sorted(elts)

    def keyset(self, steps=None, relin=True):
        """RelinKeys + GaloisKeys for the given rotation steps (default: SEAL's +-2^i set)"""
        ks = KeySet(self.ctx)
        if relin:
            ks.set_relin(self.relin_keys())
        elts = self.default_galois_elts() if steps is None else sorted({self.ctx.galois_elt(s) for s in steps})
        for g in elts:
            ks.set_galois(g, self.galois_key(g))
        return ks


class Encryptor:
    """SEAL Encryptor (public key): (u pk + e) one level above the target, divide-and-round by
    the extra prime (the rescale kernels), plaintext added to c0."""

    def __init__(self, ctx, public_key, seed=None):
        self.ctx, self.pk = ctx, public_key
        self.ops = _RingOps(ctx, seed)   # None: encryption randomness keyed from os.urandom

    def encrypt(self, pt):
        ctx, ops, ev = self.ctx, self.ops, self.ops.ev
        B, L = pt.batch, pt.limbs
        W = L + 1                      # primes 0..L: prime L is the next data prime or, at the top, P
        u = ops.ternary_ntt(B, W)
        parts = []
        for k in range(2):
            e = ops.errors_ntt(B, W)
            upk = ops.polys(ops.mul(u, self.pk[k, :W]), W)
            parts.append(ev.add(upk, ops.polys(e, W)).data[:, 0])
        big = Ciphertext(ctx, torch.stack(parts, dim=1).contiguous(), W)
        ct = ev.rescale_to_next(big)
        ct.scale = pt.scale
        ptv = Ciphertext(ctx, pt.data, L, pt.scale)
        return ev.add_plain_inplace(ct, ptv)


class Decryptor:
    """SEAL Decryptor: sum_k c_k s^k (dyadic products on the device); returns a Plaintext."""

    def __init__(self, ctx, secret_key):
        self.ctx, self.sk = ctx, secret_key
        self.ops = _RingOps(ctx, 0)      # the decryptor draws nothing

    def decrypt(self, ct):
        ops, ev, L = self.ops, self.ops.ev, ct.limbs
        s = self.sk[:L].contiguous()
        acc = ops.polys(ct.data[:, ct.size - 1, :L, :], L)
        for k in range(ct.size - 2, -1, -1):
            acc = ops.polys(ops.mul(acc.data[:, 0], s), L)
            ev.add(acc, ops.polys(ct.data[:, k, :L, :], L), out=acc)
        acc.scale = ct.scale
        return acc
