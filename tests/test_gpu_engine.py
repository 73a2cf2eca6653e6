"""Drop-in check at the level the UI calls: the Engine mirror (GPU path) against the reference's
verbatim SQL executed by SQLite with the oracle UDF (tests/sqlite_oracle.py) on the same DB file."""
import numpy as np
import pytest

from pixelbox_b200.engine import Engine, IndexedImage
from tests import sqlite_oracle

pytestmark = pytest.mark.gpu


def _corpus(seed, n, d):
    rng = np.random.default_rng(seed)
    cent = rng.integers(0, 256, size=(30, d))
    rows = np.clip(cent[rng.integers(0, 30, n)] + rng.integers(-3, 4, size=(n, d)), 0, 255).astype(np.uint8)
    rows[10:70] = rows[10]                      # ties, ordered by image_id in SQLite
    ids = np.arange(1, n + 1, dtype=np.int64)
    return ids, rows, rng


@pytest.mark.parametrize("d", [8, 256])
def test_query_by_image_hash_matches_verbatim_sql(tmp_path, d):
    ids, rows, rng = _corpus(1, 4000, d)
    path = str(tmp_path / "pixelbox.db")
    conn = sqlite_oracle.make_db(path, ids, rows)
    eng = Engine.open(path)
    try:
        assert len(eng.corpus) == len(ids)
        for md in (1e3, 0.05):
            eng.max_distance_from_query = md
            for q in (rows[10], rows[2500], rng.integers(0, 256, d, dtype=np.uint8)):
                # right-click "Search for Similar": the query is an IndexedImage carrying a stored hash (src/ui/search.rs:82-84)
                eng.query_by_image_hash_from_image(IndexedImage(visual_hash=bytes(q)))
                got = eng.get_query_results()
                want = sqlite_oracle.query(conn, bytes(q), md, 100)       # LIMIT 100, src/engine.rs:381
                assert [g.id for g in got] == [w[0] for w in want]
                assert [g.distance_from_query for g in got] == [w[1] for w in want]   # f64 equality: same f32 widened
                for g in got[:5]:
                    assert g.visual_hash == bytes(rows[g.id - 1]) and g.filename == f"img{g.id}.png"
        # missing hash: silent return, results untouched (src/engine.rs:364-368)
        before = eng.get_query_results()
        eng.query_by_image_hash_from_image(IndexedImage(visual_hash=None))
        assert [g.id for g in eng.get_query_results()] == [g.id for g in before]
        eng.clear_query_results()
        assert eng.get_query_results() is None
    finally:
        eng.close()
        conn.close()


def test_insert_appends_only_changed_rows(tmp_path):
    path = str(tmp_path / "new.db")
    eng = Engine.new(path)
    try:
        rng = np.random.default_rng(2)
        hashes = rng.integers(0, 256, size=(300, 64), dtype=np.uint8)
        for i, h in enumerate(hashes):
            eng.insert_image_from_memory(IndexedImage(filename=f"f{i}.png", path=f"/p/f{i}.png", resolution=(4, 4),
                                                      thumbnail=b"\x00", visual_hash=bytes(h)))
        assert len(eng.corpus) == 300
        # duplicate path: INSERT OR IGNORE keeps the old row, last_insert_rowid is stale, nothing is appended (src/engine.rs:231-256)
        eng.insert_image_from_memory(IndexedImage(filename="f7.png", path="/p/f7.png", resolution=(4, 4), thumbnail=b"\x00",
                                                  visual_hash=bytes(hashes[8])))
        assert len(eng.corpus) == 300
        conn = sqlite_oracle.sqlite3.connect(path)
        sqlite_oracle.register(conn)
        assert conn.execute("SELECT COUNT(*) FROM semantic_hashes").fetchone()[0] == 300
        eng.query_by_image_hash_from_image(IndexedImage(visual_hash=bytes(hashes[8])))
        got = eng.get_query_results()
        want = sqlite_oracle.query(conn, bytes(hashes[8]), 1e3, 100)
        assert [g.id for g in got] == [w[0] for w in want] and got[0].id == 9
        assert [g.distance_from_query for g in got] == [w[1] for w in want]
        conn.close()
        # re-open: the corpus is rebuilt from the table (the SQLite file is the durable state)
        eng.close()
        eng = Engine.open(path)
        assert len(eng.corpus) == 300
    finally:
        eng.close()


def test_hash_without_image_row_is_dropped_like_the_inner_join(tmp_path):
    ids, rows, _ = _corpus(3, 500, 16)
    path = str(tmp_path / "orphan.db")
    conn = sqlite_oracle.make_db(path, ids, rows)
    conn.execute("DELETE FROM images WHERE id IN (11, 12, 13)")
    conn.commit()
    eng = Engine.open(path)
    try:
        eng.query_by_image_hash_from_image(IndexedImage(visual_hash=bytes(rows[10])))
        got = [g.id for g in eng.get_query_results()]
        assert 11 not in got and 12 not in got and 13 not in got
        want = sqlite_oracle.query(conn, bytes(rows[10]), 1e3, 100)
        # upstream applies LIMIT after the join; the mirror over-fetches to fill the list the same way
        assert got == [w[0] for w in want]
    finally:
        eng.close()
        conn.close()


def test_baseline_config0_100k_table_top50(tmp_path):
    """BASELINE configs[0]: a 100k-image semantic_hashes table (256-byte hashes) in SQLite, single query, top-50.
    Reference path = the verbatim SQL with the oracle UDF (one UDF call per scanned row, as upstream); GPU path =
    Engine.open -> device corpus -> search; identical ids and f64 distances, for a right-click query and a dropped image."""
    import time
    from pixelbox_b200 import synth
    n, d = 100_000, 256
    rows = synth.synth_rows(1, 0, n, d)
    rng = np.random.default_rng(2)
    rows[5000:5040] = rows[5000]                                # a few exact duplicates: ties by image_id
    ids = np.arange(1, n + 1, dtype=np.int64)
    path = str(tmp_path / "config0.db")
    conn = sqlite_oracle.make_db(path, ids, rows)
    eng = Engine.open(path)
    try:
        for q in (rows[12345], rows[5000], rng.integers(0, 256, d, dtype=np.uint8)):
            t0 = time.perf_counter()
            want = sqlite_oracle.query(conn, bytes(q), 1e3, 50)
            t_sql = time.perf_counter() - t0
            res = eng.corpus.search(q, 50, 1e3)[0]
            assert list(res.ids) == [w[0] for w in want]
            assert [float(x) for x in res.dist] == [w[1] for w in want]
            assert t_sql > 0
        eng.query_by_image_hash_from_image(IndexedImage(visual_hash=bytes(rows[12345])))
        got = eng.get_query_results()
        want100 = sqlite_oracle.query(conn, bytes(rows[12345]), 1e3, 100)
        assert [g.id for g in got] == [w[0] for w in want100] and got[0].id == 12346 and got[0].distance_from_query <= 1e-6
    finally:
        eng.close()
        conn.close()


def test_quantizer_matches_reference_encoder():
    """src/image_hashes/efficientnet.rs:39 on the GPU against the oracle, including the README example and the edges."""
    from oracle import oracle
    from pixelbox_b200.corpus import quantize
    assert list(quantize([-1.0, 1.0, 0.0, 0.1])) == [0x00, 0xFF, 0x80, 0x8C]            # README.md:54
    special = np.array([np.nan, np.inf, -np.inf, 2.0, -2.0, 0.999, -0.999, 0.9921875, 0.99218, -0.0, 1e-9, -1e-9,
                        127.0 / 128.0, 126.9999 / 128.0, -127.5 / 128.0, 0.0078125, 0.00781249], np.float32)
    rng = np.random.default_rng(5)
    x = np.concatenate([special, rng.uniform(-1.2, 1.2, 20000).astype(np.float32), np.tanh(rng.normal(0, 1, 20000)).astype(np.float32)])
    assert np.array_equal(quantize(x), oracle.quantize(x))


def test_byte_and_hamming_distance_pairs_match_reference():
    """byte_distance (src/engine.rs:590-592) and hamming_distance (:594-604) on the GPU against the
    oracle, bit for bit, including upstream's KATs (:693-701), odd dims (byte path) and the wrapping u8 sum."""
    from oracle import oracle
    from pixelbox_b200.corpus import byte_distance_pairs, hamming_distance_pairs
    kat = [([0], [0xFF], 1.0), ([0x0F], [0xFF], 0.5), ([0x0], [0x0], 0.0), ([0b10101010], [0b01010101], 1.0)]
    for a, b, want in kat:
        d, bits_ = hamming_distance_pairs(a, b)
        assert d[0] == np.float32(want)
    d, _ = hamming_distance_pairs([[0b10101010, 0b01010101], [0xFF, 0x0F]], [[0b01010101, 0b10101010], [0x0F, 0x0F]])
    assert list(d) == [np.float32(1.0), np.float32(0.25)]
    rng = np.random.default_rng(11)
    for dim in (1, 3, 4, 31, 64, 256, 1001, 4096):
        n = 257
        a = rng.integers(0, 256, size=(n, dim), dtype=np.uint8)
        b = rng.integers(0, 256, size=(n, dim), dtype=np.uint8)
        b[:8] = a[:8]                                       # identical pairs
        b[8:16] = 255 - a[8:16]                             # every bit differs: wraps when dim * 8 > 255
        bd, l1 = byte_distance_pairs(a, b)
        hd, hb = hamming_distance_pairs(a, b)
        for i in range(0, n, 3 if dim > 256 else 1):
            assert np.float32(bd[i]).view(np.uint32) == oracle.byte_distance(a[i], b[i]).view(np.uint32), (dim, i)
            o_h, o_bits = oracle.hamming_distance(a[i], b[i])
            assert np.float32(hd[i]).view(np.uint32) == o_h.view(np.uint32) and int(hb[i]) == o_bits, (dim, i)
        assert np.array_equal(l1, np.abs(a.astype(np.int64) - b.astype(np.int64)).sum(axis=1).astype(np.uint32))
