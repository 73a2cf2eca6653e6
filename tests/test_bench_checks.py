"""bench.py's full-size parity checks are themselves checked: on a small two-shard synthetic corpus the true answer
passes `completeness_check`, and an answer with a missing row, a swapped order or a flipped distance bit does not; the
planted-tie check wants both copies first, equal distance bits, lower image_id first.  CPU only (oracle + numpy)."""
import numpy as np

import bench
from oracle import oracle


def _truth(world, rows_per_shard, dim, k, query):
    total = world * rows_per_shard
    rows = oracle.synth_rows(bench.SEED, 0, total, dim)
    ids = np.arange(1, total + 1, dtype=np.int64)
    return oracle.topk(rows, ids, query, k, 1e3, threads=2)


def test_completeness_check_accepts_the_truth_and_rejects_wrong_answers():
    world, rps, dim, k = 2, 6000, 64, 20
    rng = np.random.default_rng(3)
    query = oracle.synth_rows(bench.SEED, 777, 1, dim)[0].copy()
    query[:8] = rng.integers(0, 256, 8)
    ids, dist, _, _ = _truth(world, rps, dim, k, query)
    ok, detail = bench.completeness_check(ids, dist, query, k, dim, world, rps, stripe_rows=rps)
    assert ok, detail
    # the best row of the second shard dropped, everything below it moved up, a worse row appended at the end:
    # the returned rows still re-rank consistently, only the stripe can notice
    second = [i for i, x in enumerate(ids) if x > rps]
    assert second, "the fixture needs a hit on the second shard"
    drop = second[0]
    more_ids, more_dist, _, _ = _truth(world, rps, dim, k + 1, query)
    wrong_ids = np.concatenate([ids[:drop], ids[drop + 1:], more_ids[k:k + 1]])
    wrong_dist = np.concatenate([dist[:drop], dist[drop + 1:], more_dist[k:k + 1]])
    ok, detail = bench.completeness_check(wrong_ids, wrong_dist, query, k, dim, world, rps, stripe_rows=rps)
    assert not ok and "missing" in detail, detail
    # two rows swapped
    sw_ids, sw_dist = ids.copy(), dist.copy()
    sw_ids[[2, 3]] = sw_ids[[3, 2]]
    sw_dist[[2, 3]] = sw_dist[[3, 2]]
    ok, detail = bench.completeness_check(sw_ids, sw_dist, query, k, dim, world, rps, stripe_rows=rps)
    assert not ok and "re-rank" in detail
    # one distance off by one unit in the last place
    bad = dist.copy()
    bad.view(np.uint32)[5] += 1
    ok, _ = bench.completeness_check(ids, bad, query, k, dim, world, rps, stripe_rows=rps)
    assert not ok


def test_planted_check():
    lo, hi = bench.PLANT_IDS
    d = np.float32(0.0)
    assert bench.planted_check([lo, hi, 5], np.array([d, d, 0.5], np.float32))
    assert not bench.planted_check([hi, lo, 5], np.array([d, d, 0.5], np.float32))                 # tie broken the wrong way
    assert not bench.planted_check([lo, 5, hi], np.array([d, 0.1, 0.1], np.float32))               # a copy is not second
    assert not bench.planted_check([lo, hi], np.array([d, np.nextafter(d, np.float32(1))], np.float32))   # distance bits differ
    assert not bench.planted_check([lo], np.array([d], np.float32))
