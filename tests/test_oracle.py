"""CPU suite, part 1: the oracle itself (rbk_oracle.c) against the independent pure-Python
restatement, the golden fixtures and the S-rules of SURVEY.md §8c."""
import math

import numpy as np
import pytest

from common import load_golden


def test_cosine_bit_exact_vs_python(oracle_mod):
    from oracle import pyref
    rng = np.random.default_rng(0)
    for d in (1, 7, 384, 768, 1024, 1536):
        for _ in range(40):
            a, b = rng.standard_normal(d), rng.standard_normal(d)
            assert oracle_mod.cosine(a, b) == pyref.cosine_similarity(a.tolist(), b.tolist())


def test_cosine_length_mismatch_and_zero_vector(oracle_mod):
    with pytest.raises(ValueError, match="Vectors must have the same length"):
        oracle_mod.cosine([1.0, 2.0], [1.0])
    assert math.isnan(oracle_mod.cosine([0.0, 0.0], [1.0, 2.0]))       # S3: 0/0


def test_golden_fixtures_bit_exact(oracle_mod):
    g = load_golden()
    for c in g["cases"]:
        k_fetch = 2 * (c["top_k"] or 10)
        s, v = oracle_mod.search(c["rows_f64"], c["query_f64"], k_fetch, c["min_score"] or 0.5)
        assert s.tolist() == c["scan_slots"]
        assert v.tolist() == c["scan_scores_f64"].tolist()
        s, v = oracle_mod.find_most_similar(c["query_f64"], c["rows_f64"], c["top_k"])
        assert s.tolist() == c["fms_slots"]
        assert v.tolist() == c["fms_scores_f64"].tolist()
    for r in g["rrf"]:
        ids, sc = oracle_mod.rrf(r["fts"], r["vec"], r["top_k"])
        assert ids.tolist() == r["ids"]
        assert sc.tolist() == [float.fromhex(x) for x in r["scores"]]


def test_threshold_is_inclusive_at_exactly_half(oracle_mod):
    # q=(1,0,0,0), c=(1,1,1,1): dot=1, |q|=1, |c|=2 -> exactly 0.5, kept by `>=` (S5)
    rows = np.array([[1.0, 1, 1, 1], [1.0, 1, 1, 2]])
    s, v = oracle_mod.search(rows, [1.0, 0, 0, 0], 10, 0.5)
    assert s.tolist() == [0] and v.tolist() == [0.5]
    # and the classic trap: (1,1,0,0).(1,0,1,0) is 0.49999999999999994 in binary64, not 0.5
    s, v = oracle_mod.search(np.array([[1.0, 0, 1, 0]]), [1.0, 1, 0, 0], 10, 0.5)
    assert len(s) == 0


def test_stable_ties_follow_slot_order(oracle_mod):
    rng = np.random.default_rng(1)
    base = rng.standard_normal((5, 16))
    rows = np.concatenate([base, base, base])        # every score appears three times
    q = rng.standard_normal(16)
    s, v = oracle_mod.search(rows, q, 15, None)
    for i in range(0, 15, 3):
        assert v[i] == v[i + 1] == v[i + 2]
        assert s[i] < s[i + 1] < s[i + 2]             # S6: lower slot first


def test_live_mask_nan_rows_and_cut(oracle_mod):
    rows = np.array([[1.0, 0], [0.0, 0], [1.0, 1], [2.0, 0], [0.5, 0.1]])
    q = [1.0, 0.0]
    s, v = oracle_mod.search(rows, q, 10, 0.0001)
    assert s.tolist() == [0, 3, 4, 2]                 # zero row 1 (NaN) excluded; 0 and 3 tie -> slot order
    s, _ = oracle_mod.search(rows, q, 10, 0.0001, live=[1, 1, 1, 0, 1])
    assert s.tolist() == [0, 4, 2]
    s, _ = oracle_mod.search(rows, q, 2, 0.0001)
    assert s.tolist() == [0, 3]                       # S7 cut
    with pytest.raises(ValueError):
        oracle_mod.search(rows, [1.0, 0.0, 0.0], 2, 0.5)


def test_multithreaded_select_equals_full_stable_sort(oracle_mod):
    from runbookai_b200 import synth
    c = synth.random_corpus(3000, 40, 5)
    c[100:160] = c[40:100]                            # ties
    q = synth.random_queries(9, 40, 6).astype(np.float64)
    for ms in (None, 0.1):
        S, V, C = oracle_mod.search_batch_mt(c, q, 24, ms, n_threads=5)
        for b in range(9):
            s, v = oracle_mod.search(c, q[b], 24, ms)
            assert C[b] == len(s)
            assert S[b, :C[b]].tolist() == s.tolist() and V[b, :C[b]].tolist() == v.tolist()


def test_verify_variant_is_bit_identical_to_the_literal_loop(oracle_mod):
    """rbk_oracle_search_batch_bf16_verify (norms hoisted, 8 queries' dot chains side by side) is the checker of
    the large GPU runs: it must return exactly what the literal per-pair loop returns - slots, fp64 scores,
    counts - including ties, NaN rows, tombstones, a threshold, ragged query counts and chunked use."""
    from runbookai_b200 import synth
    for n, d, nq, k, ms in ((4000, 768, 19, 32, None), (3000, 100, 3, 8, 0.05), (900, 1536, 64, 16, None),
                            (500, 7, 9, 5, 0.2), (50, 33, 1, 112, None)):
        c = synth.random_corpus(n, d, 50 + n)
        c[5] = 0
        c[17] = c[3]
        c[n - 1] = c[3]
        q = synth.random_queries(nq, d, 60 + d).astype(np.float64)
        if nq > 2:
            q[2] = 0.0                                   # zero query: NaN everywhere
        live = np.ones(n, dtype=np.uint8)
        live[::7] = 0
        for lv in (None, live):
            a = oracle_mod.search_batch_mt(c, q, k, ms, live=lv, n_threads=3)
            b = oracle_mod.search_batch_verify(c, q, k, ms, live=lv, n_threads=5)
            assert (a[0] == b[0]).all() and np.array_equal(a[1], b[1], equal_nan=True) and (a[2] == b[2]).all()
            cc = oracle_mod.search_chunked(lambda r0, m: c[r0:r0 + m], n, q, k, ms, chunk_rows=333, live=lv,
                                           slot_base=1000)
            assert (np.where(a[0] >= 0, a[0] + 1000, -1) == cc[0]).all()
            assert np.array_equal(a[1], cc[1], equal_nan=True) and (a[2] == cc[2]).all()


def test_bf16_corpus_path_equals_f64_path(oracle_mod):
    from runbookai_b200 import synth
    c = synth.random_corpus(500, 33, 7)
    q = synth.random_queries(1, 33, 8)[0].astype(np.float64)
    a = oracle_mod.scores(c, q)
    b = oracle_mod.scores(synth.bf16_bits_to_f32(c).astype(np.float64), q)
    assert a.tolist() == b.tolist()


def test_rrf_known_answer(oracle_mod):
    # hybrid-search.ts:118-145 by hand: fts=[a,b], vec=[b,c]; k=60, w=0.4/0.6
    ids, sc = oracle_mod.rrf([0, 1], [1, 2], 10)
    a = 0.4 * (1 / 61)
    b = 0.4 * (1 / 62) + 0.6 * (1 / 61)
    c = 0.6 * (1 / 62)
    assert ids.tolist() == [1, 2, 0]
    assert sc.tolist() == [b, c, a]
