"""Reader / writer for Microsoft SEAL's binary `save` / `load` streams.

Purpose: the only route to a TRUE cross-check against SEAL -- load parameters, keys and input ciphertexts that a real
SEAL build wrote elsewhere, evaluate them on this engine and compare the result with the ciphertexts SEAL itself
produced, bit for bit (tools/seal_replay.py drives that; tools/seal_dump_vectors.cpp is the SEAL-side program).

Format (SEAL 3.4.x, the version the reference pins in README.md:6; written down from SEAL's sources
native/src/seal/{serialization,encryptionparams,smallmodulus,intarray,ciphertext,plaintext,publickey,secretkey,
kswitchkeys}.h -- SEAL is not available in this container, so everything below is "as published", and the reader
checks every redundant field it can so that a mismatch is reported rather than silently mis-parsed):

  every object = SEALHeader + body.  SEALHeader (8 bytes, little endian):
      uint16 magic = 0xA15E | uint8 zero_byte = 0 | uint8 compr_mode (0 none, 1 zlib deflate) | uint32 size
      (size = header + body as stored).  SEAL 3.5-3.7 use a 16-byte header instead:
      uint16 magic | uint8 header_size = 0x10 | uint8 major | uint8 minor | uint8 compr_mode | uint16 reserved | uint64 size
      -- recognised on input (compr_mode 2 = zstd is refused: no decoder here).
  SmallModulus         body: uint64 value
  EncryptionParameters body: uint8 scheme (1 BFV, 2 CKKS) | uint64 poly_modulus_degree | uint64 coeff_mod_count |
                             coeff_mod_count x SmallModulus object | SmallModulus object (plain_modulus)
  IntArray<uint64>     body: uint64 size | size x uint64
  Ciphertext           body: parms_id (4 x uint64) | uint8 is_ntt_form | uint64 size | uint64 poly_modulus_degree |
                             uint64 coeff_mod_count | double scale | IntArray object  (data[poly][limb][coeff])
  Plaintext            body: parms_id | double scale | IntArray object   (3.5+: parms_id | uint64 coeff_count | scale | data)
  PublicKey            body: Ciphertext object            SecretKey body: Plaintext object
  KSwitchKeys          body: parms_id | uint64 dim1 | dim1 x ( uint64 dim2 | dim2 x PublicKey object )
                       RelinKeys: index = key power - 2; GaloisKeys: index = (galois_elt - 1) / 2, absent elements
                       have dim2 = 0; each present entry holds K-1 decomposition digits = size-2 ciphertexts over
                       all K primes in NTT form (exactly the engine's key layout [digit][2][K][N]).
  parms_id             SHA3-256 (3.4.x; Blake2b from 3.5) of the uint64 array {scheme, poly_modulus_degree,
                       coeff_modulus values..., plain_modulus value}; a lower level's id hashes the chain with the last
                       primes dropped.  Used here only as a cross-check (the level is taken from coeff_mod_count).

Nothing here touches the GPU: arrays are numpy uint64 in SEAL's in-memory layout, which is also the engine's.
"""
import hashlib
import io
import struct
import zlib

import numpy as np

MAGIC = 0xA15E
SCHEME_BFV, SCHEME_CKKS = 1, 2
SEEDED_MARKER = 0xFFFFFFFFFFFFFFFF


class SealFormatError(ValueError):
    pass


# ----------------------------------------------------------------------------------------------- header
def _read_exact(f, n):
    b = f.read(n)
    if len(b) != n:
        raise SealFormatError("stream ended early (wanted %d bytes, got %d)" % (n, len(b)))
    return b


def read_object(f):
    """one SEAL object from a binary file object -> (body bytes, header dict).  Handles both header generations and
    zlib-compressed bodies."""
    h = _read_exact(f, 4)
    magic, b2, b3 = struct.unpack("<HBB", h)
    if magic != MAGIC:
        raise SealFormatError("bad magic 0x%04x (not a SEAL >= 3.4 stream)" % magic)
    if b2 == 0x00:                                   # SEAL 3.4.x: zero byte, compr_mode, uint32 size
        compr, hdr_len = b3, 8
        (size,) = struct.unpack("<I", _read_exact(f, 4))
        version = (3, 4)
    elif b2 == 0x10:                                 # SEAL 3.5+: header_size, major, minor, compr_mode, reserved, uint64 size
        rest = _read_exact(f, 12)
        minor, compr, _res, size = struct.unpack("<BBHQ", rest)
        version, hdr_len = (b3, minor), 16
    else:
        raise SealFormatError("unknown SEAL header (byte 2 = 0x%02x)" % b2)
    if size < hdr_len:
        raise SealFormatError("header size field %d smaller than the header" % size)
    body = _read_exact(f, size - hdr_len)
    if compr == 1:
        body = zlib.decompress(body)
    elif compr != 0:
        raise SealFormatError("compression mode %d is not supported (only none / zlib deflate)" % compr)
    return body, {"version": version, "compr_mode": compr, "size": size}


def write_object(f, body, version=(3, 4), compress=False):
    stored = zlib.compress(body) if compress else body
    if version < (3, 5):
        f.write(struct.pack("<HBBI", MAGIC, 0, 1 if compress else 0, 8 + len(stored)))
    else:
        f.write(struct.pack("<HBBBBHQ", MAGIC, 0x10, version[0], version[1], 1 if compress else 0, 0, 16 + len(stored)))
    f.write(stored)


def _sub(body, off):
    """nested object starting at body[off] -> (inner body, header, new offset)"""
    f = io.BytesIO(body)
    f.seek(off)
    inner, hdr = read_object(f)
    return inner, hdr, f.tell()


# ----------------------------------------------------------------------------------------------- parameters
def parms_id(scheme, n, primes, plain_modulus=0, algo="sha3_256"):
    """parms_id of a parameter set as 4 little-endian uint64 words"""
    data = struct.pack("<%dQ" % (3 + len(primes)), scheme, n, *primes, plain_modulus)
    dig = hashlib.sha3_256(data).digest() if algo == "sha3_256" else hashlib.blake2b(data, digest_size=32).digest()
    return struct.unpack("<4Q", dig)


class Params:
    def __init__(self, scheme, n, primes, plain_modulus=0, version=(3, 4)):
        self.scheme, self.n, self.primes, self.plain_modulus, self.version = scheme, int(n), [int(p) for p in primes], int(plain_modulus), version

    @property
    def log_n(self):
        return self.n.bit_length() - 1

    def id_at(self, limbs):
        """parms_id of the level with `limbs` primes (limbs = K is the key level)"""
        return parms_id(self.scheme, self.n, self.primes[:limbs], self.plain_modulus,
                        "sha3_256" if self.version < (3, 5) else "blake2b")


def load_params(f):
    body, hdr = read_object(f)
    scheme, n, cnt = struct.unpack_from("<BQQ", body, 0)
    off, primes = 17, []
    for _ in range(cnt + 1):
        inner, _, off = _sub(body, off)
        if len(inner) != 8:
            raise SealFormatError("SmallModulus body of %d bytes" % len(inner))
        primes.append(struct.unpack("<Q", inner)[0])
    if off != len(body):
        raise SealFormatError("trailing bytes after EncryptionParameters")
    return Params(scheme, n, primes[:-1], primes[-1], hdr["version"])


def save_params(f, p):
    body = io.BytesIO()
    body.write(struct.pack("<BQQ", p.scheme, p.n, len(p.primes)))
    for v in list(p.primes) + [p.plain_modulus]:
        write_object(body, struct.pack("<Q", v), p.version)
    write_object(f, body.getvalue(), p.version)


# ----------------------------------------------------------------------------------------------- arrays
def _parse_intarray(body):
    (size,) = struct.unpack_from("<Q", body, 0)
    if len(body) != 8 + 8 * size:
        raise SealFormatError("IntArray says %d words but holds %d bytes" % (size, len(body) - 8))
    return np.frombuffer(body, dtype="<u8", count=size, offset=8).copy()


def _intarray_body(a):
    a = np.ascontiguousarray(a, dtype="<u8").reshape(-1)
    return struct.pack("<Q", a.size) + a.tobytes()


class Ct:
    """a loaded ciphertext: data uint64 [size][limbs][N], scale, the stream's parms_id words"""

    def __init__(self, data, scale, pid, is_ntt=True):
        self.data, self.scale, self.parms_id, self.is_ntt = data, float(scale), tuple(pid), bool(is_ntt)

    @property
    def limbs(self):
        return self.data.shape[1]


def _parse_ct(body, params=None, what="Ciphertext"):
    pid = struct.unpack_from("<4Q", body, 0)
    is_ntt, size, n, cnt, scale = struct.unpack_from("<BQQQd", body, 32)
    inner, _, off = _sub(body, 32 + 1 + 24 + 8)
    data = _parse_intarray(inner)
    if data.size != size * cnt * n:
        if size == 2 and data.size > 0 and data.size < size * cnt * n:
            raise SealFormatError("%s is a seeded (symmetric, save-only) ciphertext: expand it with SEAL before export" % what)
        raise SealFormatError("%s: %d words for size %d x %d primes x N=%d" % (what, data.size, size, cnt, n))
    if size == 2 and data.size and int(data.reshape(size, cnt, n)[1, 0, 0]) == SEEDED_MARKER:
        raise SealFormatError("%s carries the seed marker in c1: expand it with SEAL before export" % what)
    if params is not None:
        if n != params.n:
            raise SealFormatError("%s has N=%d, parameters have N=%d" % (what, n, params.n))
        for j in range(cnt):
            if int(data.reshape(size, cnt, n)[:, j].max(initial=0)) >= params.primes[j]:
                raise SealFormatError("%s: residue >= prime %d (wrong level or wrong parameters)" % (what, j))
    return Ct(data.reshape(size, cnt, n), scale, pid, is_ntt), off


def load_ciphertext(f, params=None):
    body, _ = read_object(f)
    ct, off = _parse_ct(body, params)
    if off != len(body):
        raise SealFormatError("trailing bytes after Ciphertext")
    return ct


def _ct_body(data, scale, pid, version, is_ntt=True):
    size, cnt, n = data.shape
    out = io.BytesIO()
    out.write(struct.pack("<4Q", *pid))
    out.write(struct.pack("<BQQQd", 1 if is_ntt else 0, size, n, cnt, scale))
    write_object(out, _intarray_body(data), version)
    return out.getvalue()


def save_ciphertext(f, data, scale, params, compress=False):
    """data uint64 [size][limbs][N] at the level with data.shape[1] primes"""
    data = np.asarray(data, dtype=np.uint64)
    write_object(f, _ct_body(data, scale, params.id_at(data.shape[1]), params.version), params.version, compress)


def load_plaintext(f, params=None):
    """-> (uint64 [limbs][N], scale, parms_id).  3.4 layout: parms_id | scale | data; 3.5+: parms_id | coeff_count | scale | data"""
    body, hdr = read_object(f)
    pid = struct.unpack_from("<4Q", body, 0)
    off = 32
    if hdr["version"] >= (3, 5):
        off += 8
    (scale,) = struct.unpack_from("<d", body, off)
    inner, _, end = _sub(body, off + 8)
    data = _parse_intarray(inner)
    if params is not None and data.size % params.n == 0:
        data = data.reshape(-1, params.n)
    return data, scale, pid


def save_plaintext(f, data, scale, params):
    data = np.asarray(data, dtype=np.uint64)
    out = io.BytesIO()
    out.write(struct.pack("<4Q", *params.id_at(data.shape[0])))
    if params.version >= (3, 5):
        out.write(struct.pack("<Q", data.size))
    out.write(struct.pack("<d", scale))
    write_object(out, _intarray_body(data), params.version)
    write_object(f, out.getvalue(), params.version)


# ----------------------------------------------------------------------------------------------- key-switching keys
def load_kswitch_keys(f, params):
    """RelinKeys / GaloisKeys -> {index: uint64 [K-1][2][K][N]} in the engine's key layout.
    RelinKeys: index 0 is the key for s^2; GaloisKeys: index = (galois_elt - 1) / 2."""
    body, _ = read_object(f)
    K = len(params.primes)
    pid = struct.unpack_from("<4Q", body, 0)
    (dim1,) = struct.unpack_from("<Q", body, 32)
    off, keys = 40, {}
    for idx in range(dim1):
        (dim2,) = struct.unpack_from("<Q", body, off)
        off += 8
        if dim2 == 0:
            continue
        if dim2 != K - 1:
            raise SealFormatError("key %d has %d decomposition digits, parameters imply %d" % (idx, dim2, K - 1))
        digits = []
        for _ in range(dim2):
            pk_body, _, off = _sub(body, off)          # PublicKey object ...
            ct_body, _, end = _sub(pk_body, 0)         # ... wrapping a Ciphertext object
            ct, used = _parse_ct(ct_body, params, "key-switching key")
            if ct.data.shape != (2, K, params.n) or not ct.is_ntt:
                raise SealFormatError("key digit of shape %s" % (ct.data.shape,))
            digits.append(ct.data)
        keys[idx] = np.stack(digits)
    if off != len(body):
        raise SealFormatError("trailing bytes after KSwitchKeys")
    return keys, pid


def save_kswitch_keys(f, keys, params, dim1=None, compress=False):
    """keys: {index: uint64 [K-1][2][K][N]}; dim1 = length of the outer vector (RelinKeys: 1; GaloisKeys: N)"""
    K = len(params.primes)
    dim1 = (max(keys) + 1) if dim1 is None else dim1
    pid = params.id_at(K)
    out = io.BytesIO()
    out.write(struct.pack("<4Q", *pid))
    out.write(struct.pack("<Q", dim1))
    for idx in range(dim1):
        if idx not in keys:
            out.write(struct.pack("<Q", 0))
            continue
        k = np.asarray(keys[idx], dtype=np.uint64)
        out.write(struct.pack("<Q", k.shape[0]))
        for d in range(k.shape[0]):
            ct = io.BytesIO()
            write_object(ct, _ct_body(k[d], 1.0, pid, params.version), params.version)
            write_object(out, ct.getvalue(), params.version)       # PublicKey wrapper
    write_object(f, out.getvalue(), params.version, compress)


def galois_index(galois_elt):
    """GaloisKeys::get_index"""
    return (int(galois_elt) - 1) >> 1


def level_of(ct, params):
    """number of primes of a loaded ciphertext, cross-checked against its parms_id when the hash convention matches"""
    want = params.id_at(ct.limbs)
    return ct.limbs, (tuple(want) == tuple(ct.parms_id))
