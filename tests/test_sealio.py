"""SEAL binary stream reader / writer: byte-level layout of the 3.4.x format, round trips, the redundant
field checks, and the oracle -> file -> loader path the replay tool uses.  CPU-only (numpy)."""
import importlib
import io
import struct

import numpy as np
import pytest

PKG = "seal-fyp-logistic-regression_b200"


def _io():
    return importlib.import_module(PKG + ".sealio")


def _params(sio, version=(3, 4)):
    params = importlib.import_module(PKG + ".params")
    return sio.Params(sio.SCHEME_CKKS, 4096, params.coeff_modulus_create(12, [50, 40, 40, 50]), 0, version)


def test_header_bytes_match_seal_3_4_layout():
    sio = _io()
    buf = io.BytesIO()
    sio.write_object(buf, struct.pack("<Q", 0xffffee001))              # a SmallModulus object
    raw = buf.getvalue()
    # magic 0xA15E little endian ("5E A1"), zero byte, compr_mode none, uint32 total size = 8 + 8
    assert raw[:8] == bytes([0x5E, 0xA1, 0x00, 0x00, 0x10, 0x00, 0x00, 0x00]) and len(raw) == 16
    body, hdr = sio.read_object(io.BytesIO(raw))
    assert hdr["version"] == (3, 4) and struct.unpack("<Q", body)[0] == 0xffffee001
    # 3.6-style 16-byte header is recognised as well
    buf = io.BytesIO()
    sio.write_object(buf, b"abcdefgh", version=(3, 6), compress=True)
    raw = buf.getvalue()
    assert raw[:2] == b"\x5e\xa1" and raw[2] == 0x10 and raw[3:5] == bytes([3, 6]) and raw[5] == 1
    assert sio.read_object(io.BytesIO(raw))[0] == b"abcdefgh"


@pytest.mark.parametrize("version", [(3, 4), (3, 6)])
def test_params_ciphertext_keys_round_trip(version):
    sio = _io()
    p = _params(sio, version)
    rng = np.random.default_rng(1)
    K, n = len(p.primes), p.n
    buf = io.BytesIO()
    sio.save_params(buf, p)
    size_expected = 8 + 17 + (K + 1) * 16 if version == (3, 4) else None
    if size_expected:
        assert len(buf.getvalue()) == size_expected
    buf.seek(0)
    q = sio.load_params(buf)
    assert (q.scheme, q.n, q.primes, q.plain_modulus) == (p.scheme, p.n, p.primes, 0)
    for limbs, size, compress in ((K - 1, 2, False), (2, 3, True)):
        ct = np.stack([rng.integers(0, p.primes[j], size=(size, n), dtype=np.uint64) for j in range(limbs)], axis=1)
        buf = io.BytesIO()
        sio.save_ciphertext(buf, ct, 2.0 ** 40, p, compress=compress)
        buf.seek(0)
        got = sio.load_ciphertext(buf, p)
        assert np.array_equal(got.data, ct) and got.scale == 2.0 ** 40 and got.is_ntt
        assert sio.level_of(got, p) == (limbs, True)
    keys = {0: np.stack([np.stack([rng.integers(0, p.primes[j], size=(2, n), dtype=np.uint64) for j in range(K)], axis=1)
                         for _ in range(K - 1)])}
    buf = io.BytesIO()
    sio.save_kswitch_keys(buf, keys, p, dim1=1)
    buf.seek(0)
    got, pid = sio.load_kswitch_keys(buf, p)
    assert list(got) == [0] and np.array_equal(got[0], keys[0]) and tuple(pid) == tuple(p.id_at(K))
    gk = {sio.galois_index(3): keys[0], sio.galois_index(2 * n - 1): keys[0][::-1].copy()}
    buf = io.BytesIO()
    sio.save_kswitch_keys(buf, gk, p, dim1=n, compress=True)
    buf.seek(0)
    got, _ = sio.load_kswitch_keys(buf, p)
    assert sorted(got) == sorted(gk) and all(np.array_equal(got[i], gk[i]) for i in gk)


def test_reader_rejects_inconsistent_streams():
    sio = _io()
    p = _params(sio)
    with pytest.raises(sio.SealFormatError, match="magic"):
        sio.read_object(io.BytesIO(b"\x00" * 16))
    buf = io.BytesIO()
    ct = np.zeros((2, 3, p.n), dtype=np.uint64)
    sio.save_ciphertext(buf, ct, 1.0, p)
    raw = buf.getvalue()
    with pytest.raises(sio.SealFormatError, match="ended early"):
        sio.load_ciphertext(io.BytesIO(raw[:-9]))
    bad = ct.copy()
    bad[0, 1, 5] = p.primes[1]                        # residue not reduced modulo its prime
    buf = io.BytesIO()
    sio.save_ciphertext(buf, bad, 1.0, p)
    buf.seek(0)
    with pytest.raises(sio.SealFormatError, match="residue"):
        sio.load_ciphertext(buf, p)
    seeded = ct.copy()
    seeded[1, 0, 0] = sio.SEEDED_MARKER
    buf = io.BytesIO()
    sio.save_ciphertext(buf, seeded, 1.0, p)
    buf.seek(0)
    with pytest.raises(sio.SealFormatError, match="seed"):
        sio.load_ciphertext(buf)
    assert sio.parms_id(2, 4096, p.primes[:2]) != sio.parms_id(2, 4096, p.primes[:3])


def test_oracle_objects_survive_the_file_format(po, tmp_path):
    """what tools/seal_replay.py does, end to end on the CPU: parameters, keys and ciphertexts written in SEAL's
    format, read back, evaluated (rotate + relinearize + rescale on the oracle) -- identical to evaluating the originals"""
    sio = _io()
    log_n = 12
    primes = po.coeff_modulus_create(log_n, [50, 40, 40, 50])
    o = po.Oracle(log_n, primes)
    p = sio.Params(sio.SCHEME_CKKS, 1 << log_n, primes)
    sk = o.gen_secret(1)
    pk = o.gen_public(2, sk)
    rlk = o.gen_relin_key(3, sk)
    g = o.galois_elt(1)
    gk = o.gen_galois_key(4, sk, g)
    x = np.linspace(-1, 1, 16)
    ct = o.encrypt(5, pk, o.encode(x, 2.0 ** 40))
    d = tmp_path
    with open(d / "parms.bin", "wb") as f:
        sio.save_params(f, p)
    with open(d / "relin.bin", "wb") as f:
        sio.save_kswitch_keys(f, {0: rlk}, p, dim1=1)
    with open(d / "galois.bin", "wb") as f:
        sio.save_kswitch_keys(f, {sio.galois_index(g): gk}, p, dim1=1 << log_n, compress=True)
    with open(d / "ct.bin", "wb") as f:
        sio.save_ciphertext(f, ct, 2.0 ** 40, p)
    with open(d / "parms.bin", "rb") as f:
        p2 = sio.load_params(f)
    o2 = po.Oracle(p2.log_n, p2.primes)
    with open(d / "relin.bin", "rb") as f:
        rlk2 = sio.load_kswitch_keys(f, p2)[0][0]
    with open(d / "galois.bin", "rb") as f:
        gk2 = sio.load_kswitch_keys(f, p2)[0][sio.galois_index(g)]
    with open(d / "ct.bin", "rb") as f:
        c2 = sio.load_ciphertext(f, p2)
    want = o.rescale(o.relinearize(o.multiply(o.apply_galois(ct, g, gk), ct), rlk))
    got = o2.rescale(o2.relinearize(o2.multiply(o2.apply_galois(c2.data, g, gk2), c2.data), rlk2))
    assert np.array_equal(got, want)
