"""SEAL `CoeffModulus` helpers the reference calls when it builds EncryptionParameters
(CoeffModulus::Create: linear_transformation2.cpp:233, matrix_multiplication.cpp:147,
logistic_regression_ckks.cpp:421; BFVDefault: benchmark.cpp:137; MaxBitCount: README.md:176)."""

_WITNESSES = (2, 3, 5, 7, 11, 13, 17, 19, 23, 29, 31, 37)


def is_prime(n):
    """deterministic Miller-Rabin for 64-bit integers"""
    if n < 2:
        return False
    for w in _WITNESSES:
        if n % w == 0:
            return n == w
    d, r = n - 1, 0
    while d % 2 == 0:
        d //= 2
        r += 1
    for w in _WITNESSES:
        x = pow(w, d, n)
        if x in (1, n - 1):
            continue
        for _ in range(r - 1):
            x = x * x % n
            if x == n - 1:
                break
        else:
            return False
    return True


def _primes_of_size(bits, count, n):
    """the `count` largest primes below 2^bits that are 1 mod 2N, descending"""
    step = 2 * n
    v = (1 << bits) - step + 1
    found = []
    while len(found) < count and v > (1 << (bits - 1)):
        if is_prime(v):
            found.append(v)
        v -= step
    if len(found) < count:
        raise ValueError("failed to find enough qualifying primes")
    return found


def coeff_modulus_create(log_n, bit_sizes):
    """CoeffModulus::Create(N, bit_sizes): per bit size the c largest NTT primes; walking the
    request in order, each entry takes the smallest not-yet-used prime of its size"""
    n = 1 << log_n
    if any(b < 2 or b > 60 for b in bit_sizes):
        raise ValueError("bit_sizes is invalid")
    pools = {b: _primes_of_size(b, list(bit_sizes).count(b), n) for b in set(bit_sizes)}
    return [pools[b].pop() for b in bit_sizes]


_BFV_DEFAULT_128 = {
    12: [0xffffee001, 0xffffc4001, 0x1ffffe0001],
    13: [0x7fffffd8001, 0x7fffffc8001, 0xfffffffc001, 0xffffff6c001, 0xfffffebc001],
    14: [0xfffffffd8001, 0xfffffffa0001, 0xfffffff00001, 0x1fffffff68001, 0x1fffffff50001,
         0x1ffffffee8001, 0x1ffffffea0001, 0x1ffffffe88001, 0x1ffffffe48001],
    15: [0x7fffffffe90001, 0x7fffffffbf0001, 0x7fffffffbd0001, 0x7fffffffba0001, 0x7fffffffaa0001,
         0x7fffffffa50001, 0x7fffffff9f0001, 0x7fffffff7e0001, 0x7fffffff770001, 0x7fffffff380001,
         0x7fffffff330001, 0x7fffffff2d0001, 0x7fffffff170001, 0x7fffffff150001, 0x7ffffffef00001,
         0xfffffffff70001],
}


def bfv_default(log_n):
    """CoeffModulus::BFVDefault(N), 128-bit security"""
    return list(_BFV_DEFAULT_128[log_n])


def max_bit_count(log_n):
    """CoeffModulus::MaxBitCount(N), 128-bit classical security"""
    return {10: 27, 11: 54, 12: 109, 13: 218, 14: 438, 15: 881}[log_n]
