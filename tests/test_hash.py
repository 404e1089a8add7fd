"""Pins the partition hash (csrc/common/hash.hpp) with an independent numpy restatement and known answers.
The STRUCTURE follows create_hashes (first column sets, later columns combine with (17*37+l)*37+r,
NULL leaves the running hash unchanged; partition = h % P -- sort_shuffle/writer.rs:744-747)."""
import numpy as np
import pyarrow as pa

M = (1 << 64) - 1


def mix64(x):
    x &= M
    x ^= x >> 30
    x = (x * 0xbf58476d1ce4e5b9) & M
    x ^= x >> 27
    x = (x * 0x94d049bb133111eb) & M
    x ^= x >> 31
    return x


SEED = 0x9E3779B97F4A7C15


def hash_i64(v):
    return mix64((v + SEED) & M)


def hash_bytes(b: bytes):
    h = SEED ^ ((len(b) * 0xFF51AFD7ED558CCD) & M)
    i = 0
    while i + 8 <= len(b):
        h = mix64(h ^ int.from_bytes(b[i:i + 8], "little"))
        i += 8
    if i < len(b):
        h = mix64(h ^ int.from_bytes(b[i:], "little") ^ 0x8000000000000000)
    return mix64(h)


def combine(l, r):
    return ((17 * 37 + l) * 37 + r) & M


def test_known_answers():
    assert mix64(0) == 0
    assert hash_i64(0) == mix64(SEED) == 0xe220a8397b1dcdaf
    assert hash_i64(1) == 0x910a2dec89025cc1
    assert hash_i64(2) == mix64((2 * SEED + 2 - SEED) & M)
    assert hash_bytes(b"") == mix64(SEED)
    assert hash_bytes(b"ASIA") == hash_bytes(b"ASIA") != hash_bytes(b"ASIB")


def test_oracle_matches_numpy_restatement(oracle):
    rng = np.random.default_rng(7)
    n = 500
    ints = rng.integers(-2**62, 2**62, n)
    strs = ["".join(chr(97 + int(c)) for c in rng.integers(0, 26, int(l))) for l in rng.integers(0, 20, n)]
    mask = rng.random(n) < 0.2
    b = pa.RecordBatch.from_arrays([pa.array(ints, type=pa.int64()), pa.array(strs),
                                    pa.array([None if m else int(v) for m, v in zip(mask, ints)], type=pa.int64())], names=["i", "s", "n"])
    for keys, P in (([0], 16), ([1], 7), ([0, 1], 200), ([2, 1], 5), ([1, 2, 0], 3)):
        h, pid = oracle.hash_partition_ids(b, keys, P)
        for r in range(n):
            acc, first = 0, True
            for k in keys:
                v = b.column(k)[r].as_py()
                if v is not None:
                    hv = hash_i64(v & M if v >= 0 else v + (1 << 64)) if isinstance(v, int) else hash_bytes(v.encode())
                    acc = hv if first else combine(hv, acc)
                first = False
            assert int(h[r]) == acc
            assert int(pid[r]) == acc % P


def test_int32_and_int64_hash_alike(oracle):
    a = pa.RecordBatch.from_arrays([pa.array([1, -5, 7], type=pa.int32())], names=["k"])
    b = pa.RecordBatch.from_arrays([pa.array([1, -5, 7], type=pa.int64())], names=["k"])
    assert list(oracle.hash_partition_ids(a, [0], 11)[1]) == list(oracle.hash_partition_ids(b, [0], 11)[1])
