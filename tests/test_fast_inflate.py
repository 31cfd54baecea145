"""ibwa_b200/csrc/fast_inflate.h (the DEFLATE decoder of the read-ingest path) against zlib: every block type, every
compression level and strategy, inputs from one byte to megabytes, output chunks of every awkward size (the decoder is
resumable at symbol boundaries and carries 32 KB of history), truncated and damaged streams (it must say "error" or
produce zlib's bytes — never anything else, and never touch memory it does not own)."""
import ctypes
import os
import subprocess
import zlib

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
LIB = os.path.join(HERE, "harness", "libinflateharness.so")
SRC = [os.path.join(HERE, "harness", "inflate_harness.cpp"), os.path.join(ROOT, "ibwa_b200", "csrc", "fast_inflate.h")]


@pytest.fixture(scope="module")
def fi():
    if not os.path.exists(LIB) or any(os.path.getmtime(s) > os.path.getmtime(LIB) for s in SRC):
        subprocess.check_call(["g++", "-O2", "-g", "-std=c++17", "-fPIC", "-shared", "-o", LIB, SRC[0]])
    L = ctypes.CDLL(LIB)
    L.fi_inflate.restype = ctypes.c_int
    L.fi_inflate.argtypes = [ctypes.c_void_p, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_size_t,
                             ctypes.POINTER(ctypes.c_size_t), ctypes.POINTER(ctypes.c_size_t)]

    def run(raw: bytes, cap: int, chunk: int = 1 << 20):
        src = np.frombuffer(raw, dtype=np.uint8).copy() if raw else np.zeros(1, np.uint8)
        out = np.empty(cap + 64, dtype=np.uint8)
        n_out, used = ctypes.c_size_t(), ctypes.c_size_t()
        st = L.fi_inflate(src.ctypes.data, len(raw), out.ctypes.data, cap, chunk, ctypes.byref(n_out), ctypes.byref(used))
        return st, out[:n_out.value].tobytes(), used.value
    return run


def deflate(data: bytes, level=6, strategy=zlib.Z_DEFAULT_STRATEGY, wbits=-15, mem=8) -> bytes:
    c = zlib.compressobj(level, zlib.DEFLATED, wbits, mem, strategy)
    return c.compress(data) + c.flush()


def fastq_like(rng, n_reads, length):
    nt = np.frombuffer(b"ACGT", dtype=np.uint8)
    out = []
    for i in range(n_reads):
        q = np.clip(rng.normal(70, 4, size=length), 35, 74).astype(np.uint8)
        out.append(b"@r%d\n" % i + nt[rng.integers(0, 4, size=length)].tobytes() + b"\n+\n" + q.tobytes() + b"\n")
    return b"".join(out)


def corpora():
    rng = np.random.default_rng(1)
    yield "empty", b""
    yield "one_byte", b"A"
    yield "fastq", fastq_like(rng, 6000, 100)
    yield "fastq_const_quality", b"".join(b"@r%d\nACGTTGCAAC\n+\nIIIIIIIIII\n" % i for i in range(30000))
    yield "random_bytes", rng.integers(0, 256, size=300_000, dtype=np.uint8).tobytes()
    yield "zeros", bytes(700_000)
    yield "short_period", (b"abcde" * 100_000)[:400_001]
    yield "two_symbols", rng.integers(0, 2, size=200_000, dtype=np.uint8).tobytes()
    yield "long_range", rng.integers(0, 256, size=40_000, dtype=np.uint8).tobytes() * 9   # matches at distance 40 000 > 32 768: none; and near 32 K
    text = rng.integers(0, 256, size=32_768, dtype=np.uint8).tobytes()
    yield "distance_32768", text + text + text[:1000]
    yield "skewed", rng.choice(np.arange(256, dtype=np.uint8), size=500_000, p=np.r_[0.9, np.full(255, 0.1 / 255)]).tobytes()


@pytest.mark.parametrize("name,data", list(corpora()), ids=[n for n, _ in corpora()])
def test_levels_strategies_and_chunk_sizes(fi, name, data):
    for level, strategy in [(0, zlib.Z_DEFAULT_STRATEGY), (1, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_DEFAULT_STRATEGY),
                            (9, zlib.Z_DEFAULT_STRATEGY), (6, zlib.Z_FIXED), (6, zlib.Z_HUFFMAN_ONLY), (6, zlib.Z_RLE),
                            (4, zlib.Z_FILTERED)]:
        raw = deflate(data, level, strategy)
        for chunk in (1 << 20, 333, 65536 + 17):
            if chunk == 333 and len(data) > 400_000:
                continue
            st, out, used = fi(raw, len(data), chunk)
            assert st == 1, (name, level, strategy, chunk)
            assert out == data, (name, level, strategy, chunk)
            assert used == len(raw), (name, level, strategy, chunk, used, len(raw))


def test_many_flushed_blocks_and_small_windows(fi):
    """Z_SYNC_FLUSH / Z_FULL_FLUSH put empty stored blocks between Huffman blocks; small windows and memory levels
    change the block structure"""
    rng = np.random.default_rng(3)
    data = fastq_like(rng, 3000, 150)
    for wbits, mem in ((-15, 9), (-9, 1), (-12, 4)):
        c = zlib.compressobj(6, zlib.DEFLATED, wbits, mem)
        raw = b""
        for s in range(0, len(data), 7001):
            raw += c.compress(data[s:s + 7001]) + c.flush(zlib.Z_FULL_FLUSH if s % 3 == 0 else zlib.Z_SYNC_FLUSH)
        raw += c.flush()
        for chunk in (1 << 20, 4099):
            st, out, used = fi(raw, len(data), chunk)
            assert st == 1 and out == data and used == len(raw)


def test_trailing_bytes_are_left_alone(fi):
    data = b"ACGT" * 5000
    raw = deflate(data)
    st, out, used = fi(raw + b"\x12\x34\x56\x78trailer and the next gzip member", len(data))
    assert st == 1 and out == data and used == len(raw)


def test_truncated_and_damaged_streams(fi):
    """never a wrong 'done': a damaged stream either still decodes to what zlib makes of it, or is refused"""
    rng = np.random.default_rng(4)
    data = fastq_like(rng, 800, 100)
    raw = deflate(data)
    for cut in list(range(0, 40)) + [len(raw) // 3, len(raw) // 2, len(raw) - 5, len(raw) - 1]:
        st, out, _ = fi(raw[:cut], len(data))
        assert st in (2, 3), cut   # refused (3: the harness's output cap, hit by a last bogus match); what it wrote is not used
    n_same = n_refused = 0
    for k in range(400):
        bad = bytearray(raw)
        pos = int(rng.integers(0, len(raw)))
        bad[pos] ^= 1 << int(rng.integers(0, 8))
        try:
            d = zlib.decompressobj(-15)
            want = d.decompress(bytes(bad)) + d.flush()
            ok = d.eof
        except zlib.error:
            want, ok = None, False
        st, out, _ = fi(bytes(bad), 4 * len(data) + 100_000)
        if st == 1:
            assert ok and out == want, (k, pos)
            n_same += 1
        else:
            n_refused += 1
    assert n_same + n_refused == 400 and n_refused > 0 and n_same > 0
