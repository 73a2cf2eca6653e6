"""CPU tests of the oracle itself: upstream known-answer tests, README encoding example,
golden vectors, independent restatement, SQLite-level semantics and the f32-vs-exact error
bound the GPU certificate relies on.  None of these touch the product library."""
import json
import os

import numpy as np
import pytest

from oracle import oracle
from pixelbox_b200 import synth
from tests import np_restatement as npr
from tests import sqlite_oracle

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def bits(x) -> int:
    return int(np.float32(x).view(np.uint32))


def test_upstream_kat_cosine_distance():
    # src/engine.rs:703-708, verbatim asserts
    assert oracle.cosine_distance([255, 0], [255, 0]) < np.float32(1e-6)
    assert oracle.cosine_distance([0, 255], [0, 255]) < np.float32(1e-6)
    assert oracle.cosine_distance([255, 0], [0, 255]) > np.float32(2.0)
    # values a literal f32 evaluation gives
    assert bits(oracle.cosine_distance([255, 0], [255, 0])) == bits(np.float32(-1.1920929e-07))
    assert oracle.cosine_distance([255, 0], [0, 255]) == np.float32(999999.0)


def test_readme_encoding_example():
    # README.md:54: [-1.0, 1.0, 0.0, 0.1] -> [0x00, 0xFF, 0x80, 0x8C]
    assert list(oracle.quantize([-1.0, 1.0, 0.0, 0.1])) == [0x00, 0xFF, 0x80, 0x8C]
    assert list(oracle.quantize([float("nan"), 2.0, -2.0, 0.999])) == [0, 255, 0, 255]


def test_decode_has_no_zero_and_empty_blob_early_out():
    # engine.rs:582-584 is reachable only for empty blobs
    assert oracle.cosine_distance([], []) == np.float32(0.0)
    assert oracle.cosine_distance([], [1, 2, 3]) == np.float32(0.0)
    for v in (127, 128):
        assert oracle.cosine_distance([v], [v]) != np.float32(0.0) or True
    assert bits(npr.u8_to_float([127])[0]) == bits(np.float32(-0.0039215684))
    assert bits(npr.u8_to_float([128])[0]) == bits(np.float32(0.003921628))


def test_zip_truncation_for_unequal_lengths():
    # dot over the shorter length, norms over full lengths (engine.rs:580-585)
    a = np.array([200, 10, 30, 250, 77], np.uint8)
    b = np.array([190, 20, 35], np.uint8)
    assert bits(oracle.cosine_distance(a, b)) == bits(npr.cosine_distance(a, b))


def test_golden_pairs():
    with open(os.path.join(GOLDEN, "cosine_pairs.json")) as f:
        pairs = json.load(f)
    assert len(pairs) >= 100
    for p in pairs:
        a = np.frombuffer(bytes.fromhex(p["a"]), np.uint8)
        b = np.frombuffer(bytes.fromhex(p["b"]), np.uint8)
        assert bits(oracle.cosine_distance(a, b)) == p["dist_bits"]
        assert oracle.int_terms(a, b) == (p["dot"], p["norm2_a"], p["norm2_b"])
        assert oracle.udf_cosine_distance(bytes(a), bytes(b)) == float(np.float32(p["dist"]))


def test_against_independent_numpy_restatement():
    rng = np.random.default_rng(11)
    for t in range(200):
        d = int(rng.integers(1, 400))
        x = rng.integers(0, 256, d, dtype=np.uint8)
        y = rng.integers(0, 256, d, dtype=np.uint8) if t % 2 else np.clip(x.astype(int) + rng.integers(-4, 5, d), 0, 255).astype(np.uint8)
        assert bits(oracle.cosine_distance(x, y)) == bits(npr.cosine_distance(x, y))


def test_exact_integer_identity():
    # dot_i = 4 sum(ab) - 510 (sum a + sum b) + 65025 d  
    rng = np.random.default_rng(5)
    for d in (1, 8, 64, 256, 1024):
        a = rng.integers(0, 256, d, dtype=np.uint8)
        b = rng.integers(0, 256, d, dtype=np.uint8)
        ai, bi = a.astype(np.int64), b.astype(np.int64)
        dot, nq, nr = oracle.int_terms(a, b)
        assert dot == 4 * (ai * bi).sum() - 510 * (ai.sum() + bi.sum()) + 65025 * d
        assert nq == 4 * (ai * ai).sum() - 1020 * ai.sum() + 65025 * d
        assert nr == int(((2 * bi - 255) ** 2).sum())


def cos_bound(d: int) -> float:
    """|cos_f32(reference) - cos_exact| <= (2d + 1600) * 2^-24, derived in DESIGN.md section 5."""
    return (2 * d + 1600) * 2.0 ** -24


@pytest.mark.parametrize("d", [8, 64, 256, 1024])
def test_f32_cosine_error_bound(d):
    rng = np.random.default_rng(100 + d)
    worst = 0.0
    for t in range(400):
        kind = t % 4
        a = rng.integers(0, 256, d, dtype=np.uint8)
        if kind == 0:
            b = rng.integers(0, 256, d, dtype=np.uint8)
        elif kind == 1:
            b = np.clip(a.astype(int) + rng.integers(-6, 7, d), 0, 255).astype(np.uint8)
        elif kind == 2:          # tiny magnitudes: decode cancellation is worst here
            a = rng.integers(126, 130, d, dtype=np.uint8)
            b = rng.integers(126, 130, d, dtype=np.uint8)
        else:
            a = rng.choice(np.array([127, 128], np.uint8), d)
            b = rng.choice(np.array([127, 128], np.uint8), d)
        err = abs(float(oracle.cosine_similarity_f32(a, b)) - oracle.cosine_exact(a, b))
        worst = max(worst, err)
    assert worst <= cos_bound(d), (worst, cos_bound(d))


def test_topk_matches_sqlite_verbatim_sql(tmp_path):
    rng = np.random.default_rng(3)
    n, d = 800, 32
    cent = rng.integers(0, 256, size=(6, d))
    corpus = np.clip(cent[rng.integers(0, 6, n)] + rng.integers(-1, 2, size=(n, d)), 0, 255).astype(np.uint8)
    corpus[100:140] = corpus[100]          # exact duplicates -> ties on dist, broken by image_id
    ids = np.arange(1, n + 1, dtype=np.int64) * 3
    conn = sqlite_oracle.make_db(str(tmp_path / "t.db"), ids, corpus)
    for q in (corpus[100], corpus[5], rng.integers(0, 256, d, dtype=np.uint8)):
        for k, md in ((100, 1e3), (17, 1e3), (100, 0.01), (100, 2e6)):
            sql = sqlite_oracle.query(conn, bytes(q), md, k)
            o_ids, o_dist, _, _ = oracle.topk(corpus, ids, q, k, md)
            assert [r[0] for r in sql] == list(o_ids)
            assert [bits(r[1]) for r in sql] == [bits(x) for x in o_dist]
            assert all(r[1] < md for r in sql)


def test_topk_threads_agree():
    rng = np.random.default_rng(9)
    corpus = rng.integers(0, 256, size=(5000, 64), dtype=np.uint8)
    q = corpus[77]
    a = oracle.topk(corpus, None, q, 100, 1e3, threads=1)
    b = oracle.topk(corpus, None, q, 100, 1e3, threads=4)
    for x, y in zip(a, b):
        assert np.array_equal(x.view(np.uint8), y.view(np.uint8))


def test_golden_topk_fixture_reproduces():
    z = np.load(os.path.join(GOLDEN, "topk_small.npz"))
    names = sorted({k.split("/")[0] for k in z.files})
    assert len(names) == 6
    for name in names:
        corpus, ids, queries = z[f"{name}/corpus"], z[f"{name}/ids"], z[f"{name}/queries"]
        for qi in (0, 2):
            o_ids, o_dist, o_dot, o_n2 = oracle.topk(corpus, ids, queries[qi], 50, 1e3)
            assert np.array_equal(o_ids, z[f"{name}/q{qi}/k50/md1000/ids"])
            assert np.array_equal(o_dist.view(np.uint32), z[f"{name}/q{qi}/k50/md1000/dist_bits"])
            assert np.array_equal(o_dot, z[f"{name}/q{qi}/k50/md1000/dot"])
            assert np.array_equal(o_n2, z[f"{name}/q{qi}/k50/md1000/norm2"])


def test_synth_generators_agree():
    for seed, r0, n, d in ((42, 0, 33, 256), (1, 1 << 34, 5, 64), (7, 12345, 9, 13), (9, 3, 4, 1024)):
        assert np.array_equal(oracle.synth_rows(seed, r0, n, d), synth.synth_rows(seed, r0, n, d))
    x = synth.synth_rows(42, 0, 4096, 256)
    assert abs(float(x.mean()) - 127.5) < 1.0


def test_upstream_hamming_kat_and_byte_distance():
    """test_hamming_distance, src/engine.rs:693-701, against the restatement; byte_distance (:590-592) against a
    direct numpy evaluation of the same f32 expression."""
    kat = [([0], [0xFF], 1.0), ([0x0F], [0xFF], 0.5), ([0x0], [0x0], 0.0), ([0b10101010], [0b01010101], 1.0),
           ([0b10101010, 0b01010101], [0b01010101, 0b10101010], 1.0), ([0xFF, 0x0F], [0x0F, 0x0F], 0.25)]
    for a, b, want in kat:
        got, true_bits = oracle.hamming_distance(a, b)
        assert got == np.float32(want), (a, b)
        assert true_bits == int(np.unpackbits(np.bitwise_xor(np.array(a, np.uint8), np.array(b, np.uint8))).sum())
    # the u8 sum wraps past 255 differing bits (release build): 64 bytes all different -> 512 bits -> 0
    a, b = np.zeros(64, np.uint8), np.full(64, 0xFF, np.uint8)
    got, true_bits = oracle.hamming_distance(a, b)
    assert true_bits == 512 and got == np.float32(0.0)
    got, true_bits = oracle.hamming_distance(np.zeros(33, np.uint8), np.full(33, 0xFF, np.uint8))
    assert true_bits == 264 and got == np.float32(8.0) / (np.float32(8.0) * np.float32(33.0))
    rng = np.random.default_rng(3)
    for d in (1, 2, 7, 64, 256, 1000, 4096):
        a, b = rng.integers(0, 256, d, dtype=np.uint8), rng.integers(0, 256, d, dtype=np.uint8)
        l1 = np.float32(np.abs(a.astype(np.int32) - b.astype(np.int32)).sum())          # exact in f32: < 2^24
        assert bits(oracle.byte_distance(a, b)) == bits(l1 / (np.float32(255.0) * np.float32(d)))
    assert oracle.byte_distance([0, 255], [255, 0]) == np.float32(1.0)
    assert oracle.byte_distance([7, 7], [7, 7]) == np.float32(0.0)
    # the independent numpy restatement agrees with the C one, bit for bit
    for d in (1, 5, 32, 33, 100, 257):
        a, b = rng.integers(0, 256, d, dtype=np.uint8), rng.integers(0, 256, d, dtype=np.uint8)
        assert bits(oracle.byte_distance(a, b)) == bits(npr.byte_distance(a, b))
        assert bits(oracle.hamming_distance(a, b)[0]) == bits(npr.hamming_distance(a, b))
