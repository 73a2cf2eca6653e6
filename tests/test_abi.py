"""CPU-side checks of the C-ABI boundary: the library loads, exports exactly what
include/pixelbox_b200.h declares, fails loudly without a GPU, and the host-only merge step
(pbx_merge_hits) reproduces the oracle's global order.  No compute calls need a GPU here."""
import ctypes
import os
import re

import numpy as np
import pytest

from oracle import oracle
from pixelbox_b200 import _native as nat
from pixelbox_b200 import build as pbx_build
from pixelbox_b200.corpus import Corpus, merge_hits

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module", autouse=True)
def _built():
    pbx_build.build()


def _have_gpu() -> bool:
    return nat.lib().pbx_device_count() > 0


def test_header_and_library_agree():
    with open(os.path.join(ROOT, "include", "pixelbox_b200.h")) as f:
        text = f.read()
    declared = re.findall(r"PBX_API\s+[\w\s\*]+?\b(pbx_\w+)\s*\(", text)
    assert sorted(declared) == sorted(nat.EXPORTS)
    L = ctypes.CDLL(nat.SO_PATH)
    for name in declared:
        assert hasattr(L, name), name


def test_library_exports_nothing_else():
    import subprocess
    out = subprocess.run(["nm", "-D", "--defined-only", nat.SO_PATH], capture_output=True, text=True).stdout
    pbx = sorted(line.split()[-1] for line in out.splitlines() if " T " in line and line.split()[-1].startswith("pbx_"))
    assert pbx == sorted(nat.EXPORTS)


def test_hit_record_layout():
    assert nat.HIT_DTYPE.itemsize == 24
    assert nat.HIT_DTYPE.fields["image_id"][1] == 0
    assert nat.HIT_DTYPE.fields["dist"][1] == 8
    assert nat.HIT_DTYPE.fields["dot"][1] == 12
    assert nat.HIT_DTYPE.fields["norm2"][1] == 16
    assert nat.HIT_DTYPE.fields["flags"][1] == 20
    assert b"sm_100a" in nat.lib().pbx_version()


def test_argument_errors_need_no_gpu():
    L = nat.lib()
    h = ctypes.c_void_p(0)
    assert L.pbx_corpus_create(0, 0, 0, ctypes.byref(h)) == -2            # PBX_E_DIM
    assert L.pbx_corpus_create(nat.PBX_MAX_DIM + 1, 0, 0, ctypes.byref(h)) == -2
    assert L.pbx_corpus_create(8, 0, 0, None) == -1                        # PBX_E_INVALID
    assert L.pbx_search(None, None, 1, 10, 1e3, None, None, None, None, None) == -1
    assert b"NULL" in L.pbx_last_error()
    n = ctypes.c_uint64(7)
    assert L.pbx_corpus_size(None, ctypes.byref(n)) == -1


def test_no_gpu_means_error_not_fallback():
    if _have_gpu():
        pytest.skip("a B200 is visible; the no-device path cannot be exercised")
    with pytest.raises(nat.PbxError) as ei:
        Corpus(256)
    assert ei.value.code == -5 and "no CPU fallback" in str(ei.value)
    a = np.zeros((1, 8), np.uint8)
    out = np.zeros(1, np.float32)
    assert nat.lib().pbx_cosine_distance_pairs(0, nat.ptr(a), nat.ptr(a), 1, 8, nat.ptr(out), None, None, None) == -5


def _oracle_hits(corpus, ids, q, k, md):
    o_ids, o_dist, o_dot, o_n2 = oracle.topk(corpus, ids, q, k, md)
    h = np.zeros(k, nat.HIT_DTYPE)
    h["image_id"] = np.iinfo(np.int64).max
    h["dist"] = np.inf
    c = len(o_ids)
    h["image_id"][:c], h["dist"][:c], h["dot"][:c], h["norm2"][:c] = o_ids, o_dist, o_dot, o_n2
    return h, c


@pytest.mark.parametrize("n_shards", [1, 2, 3, 8])
def test_merge_hits_equals_global_topk(n_shards):
    """Shard the rows round-robin, take each shard's oracle top-k (what a GPU shard returns),
    merge with pbx_merge_hits and compare with the oracle over the whole corpus."""
    rng = np.random.default_rng(40 + n_shards)
    n, d, k = 3000, 32, 50
    cent = rng.integers(0, 256, size=(5, d))
    corpus = np.clip(cent[rng.integers(0, 5, n)] + rng.integers(-1, 2, size=(n, d)), 0, 255).astype(np.uint8)
    corpus[500:560] = corpus[500]                      # duplicates -> equal distances across shards
    ids = rng.permutation(np.arange(1, n + 1)).astype(np.int64)
    queries = np.stack([corpus[500], corpus[7], rng.integers(0, 256, d, dtype=np.uint8)])
    for md in (1e3, 0.05):
        gathered = np.zeros((n_shards, len(queries), k), nat.HIT_DTYPE)
        counts = np.zeros((n_shards, len(queries)), np.uint32)
        for s in range(n_shards):
            for qi, q in enumerate(queries):
                gathered[s, qi], counts[s, qi] = _oracle_hits(corpus[s::n_shards], ids[s::n_shards], q, k, md)
        out, cnt = merge_hits(gathered, counts, k)
        for qi, q in enumerate(queries):
            want, c = _oracle_hits(corpus, ids, q, k, md)
            assert cnt[qi] == c
            assert np.array_equal(out[qi]["image_id"][:c], want["image_id"][:c])
            assert np.array_equal(out[qi]["dist"][:c].view(np.uint32), want["dist"][:c].view(np.uint32))
            assert np.array_equal(out[qi]["dot"][:c], want["dot"][:c])


def test_merge_hits_argument_errors():
    g = np.zeros((1, 1, 4), nat.HIT_DTYPE)
    c = np.zeros((1, 1), np.uint32)
    o = np.zeros((1, 4), nat.HIT_DTYPE)
    oc = np.zeros(1, np.uint32)
    L = nat.lib()
    assert L.pbx_merge_hits(None, nat.ptr(c), 1, 1, 4, nat.ptr(o), nat.ptr(oc)) == -1
    assert L.pbx_merge_hits(nat.ptr(g), nat.ptr(c), 1, 1, 0, nat.ptr(o), nat.ptr(oc)) == -1
    assert L.pbx_merge_hits(nat.ptr(g), nat.ptr(c), 1, 1, 4, nat.ptr(o), nat.ptr(oc)) == 0 and oc[0] == 0


def test_product_never_touches_the_oracle():
    """The oracle is test infrastructure: nothing under pixelbox_b200/ (Python or CUDA) may import, link or read it,
    and the shared library must not depend on libpbx_oracle."""
    import os
    import re
    import subprocess
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    pkg = os.path.join(root, "pixelbox_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                text = open(os.path.join(dirpath, f), encoding="utf-8", errors="replace").read()
                assert not re.search(r"^\s*(from|import)\s+oracle\b", text, re.M), f"{f} imports the oracle"
                assert "pbx_oracle" not in text, f"{f} refers to the oracle library"
    so = os.path.join(pkg, "lib", "libpixelbox_b200.so")
    needed = subprocess.run(["readelf", "-d", so], capture_output=True, text=True).stdout
    assert "pbx_oracle" not in needed
