"""CPU tests of the oracle itself: golden vectors produced by the reference's own
search_one_by_one, hand-computable known answers, C-vs-numpy restatements, fp64 truth."""
import os
import tempfile

import numpy as np
import pytest

from convdr_b200 import synth
from oracle import c_oracle, flat_ip

GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "search_one_by_one.npz")


def golden_cases():
    z = np.load(GOLDEN)
    names = sorted({k.split("/")[0] for k in z.files})
    return z, names


def case_inputs(params):
    n, W, nq, topN, dup = (int(v) for v in params)
    P = synth.block(0, n, seed=7, stream=0)
    if dup:
        P[n - dup:] = P[:dup]
    Q = synth.block(0, nq, seed=7, stream=1)
    blocks = [(P[np.arange(r, n, W)], np.arange(r, n, W, dtype=np.int64)) for r in range(W)]
    return P, Q, blocks, topN


@pytest.mark.parametrize("name", golden_cases()[1])
def test_search_one_by_one_matches_reference_golden(name):
    z, _ = golden_cases()
    P, Q, blocks, topN = case_inputs(z[name + "/params"])
    with tempfile.TemporaryDirectory() as tmp:
        for b, (emb, ids) in enumerate(blocks):
            flat_ip.write_block(tmp, b, emb, ids)
        D, I = flat_ip.search_one_by_one(tmp, flat_ip.IndexFlatIP(768), Q, topN)
    assert D.dtype == z[name + "/D"].dtype and I.dtype == z[name + "/I"].dtype
    assert D.shape == z[name + "/D"].shape
    np.testing.assert_array_equal(I, z[name + "/I"])
    np.testing.assert_array_equal(D, z[name + "/D"])


def test_block_format_roundtrip_is_reference_pickle_protocol4():
    import pickle
    emb = synth.block(0, 11)
    ids = np.arange(11, dtype=np.int64) * 3 + 1
    with tempfile.TemporaryDirectory() as tmp:
        flat_ip.write_block(tmp, 2, emb, ids)
        assert sorted(os.listdir(tmp)) == ["passage__emb_p__data_obj_2.pb", "passage__embid_p__data_obj_2.pb"]
        with open(os.path.join(tmp, "passage__emb_p__data_obj_2.pb"), "rb") as f:
            raw = f.read()
        assert raw[:2] == b"\x80\x04"  # pickle protocol 4 (utils/util.py:111)
        e2, i2 = flat_ip.read_block(tmp, 2)
    np.testing.assert_array_equal(e2, emb)
    np.testing.assert_array_equal(i2, ids)
    assert e2.dtype == np.float32 and i2.dtype == np.int64


def unit(i, d=768):
    v = np.zeros(d, dtype=np.float32)
    v[i] = 1.0
    return v


def test_known_answer_identity_rows():
    P = np.eye(768, dtype=np.float32)[:300]
    Q = np.stack([unit(5) * 2 + unit(7), unit(299) + 0.5 * unit(0)]).astype(np.float32)
    D, I = flat_ip.IndexFlatIP(768).search(Q, 3) if False else flat_ip.knn_inner_product(Q, P, 3)
    assert I[0].tolist()[:2] == [5, 7] and D[0].tolist()[:2] == [2.0, 1.0]
    assert I[1].tolist()[:2] == [299, 0] and D[1].tolist()[:2] == [1.0, 0.5]
    assert D[0, 2] == 0.0 and I[0, 2] == 0  # all remaining scores tie at 0 -> lowest index


def test_known_answer_planted_scaled_copies():
    rng = np.random.default_rng(3)
    P = synth.block(0, 2000, seed=1)
    q = synth.block(0, 1, seed=1, stream=1)
    plant = {17: 3.0, 1500: 2.5, 999: 2.0, 4: 1.5}
    for row, s in plant.items():
        P[row] = q[0] * np.float32(s)
    D, I = flat_ip.knn_inner_product(q, P, 4)
    assert I[0].tolist() == [17, 1500, 999, 4]
    np.testing.assert_allclose(D[0], [3.0, 2.5, 2.0, 1.5], rtol=1e-6)
    del rng


def test_all_equal_scores_tie_rule_lowest_index():
    P = np.tile(unit(3), (50, 1))
    D, I = flat_ip.knn_inner_product(unit(3)[None], P, 7)
    assert I[0].tolist() == list(range(7))
    Dc, Ic = c_oracle.knn_ip_heap(unit(3)[None], P, 7)
    assert sorted(Ic[0].tolist()) == list(range(7))  # strict '>' heap keeps the 7 lowest indices
    np.testing.assert_array_equal(Dc, D)


def test_k_larger_than_ntotal_pads_minus_one():
    P = synth.block(0, 5)
    q = synth.block(0, 2, stream=1)
    for fn in (flat_ip.knn_inner_product, c_oracle.knn_ip_heap):
        D, I = fn(q, P, 8)
        assert (I[:, 5:] == -1).all() and (D[:, 5:] == -flat_ip.FLT_MAX).all()
        assert sorted(I[0, :5].tolist()) == [0, 1, 2, 3, 4]
        assert (np.diff(D[:, :5], axis=1) <= 0).all()


def test_add_twice_equals_add_once_and_reset():
    P = synth.block(0, 700)
    q = synth.block(0, 9, stream=1)
    a = flat_ip.IndexFlatIP(768)
    a.add(P)
    b = flat_ip.IndexFlatIP(768)
    b.add(P[:300])
    b.add(P[300:])
    assert a.ntotal == b.ntotal == 700
    Da, Ia = a.search(q, 20)
    Db, Ib = b.search(q, 20)
    np.testing.assert_array_equal(Ia, Ib)
    np.testing.assert_array_equal(Da, Db)
    b.reset()
    assert b.ntotal == 0
    D0, I0 = b.search(q, 3)
    assert (I0 == -1).all()


@pytest.mark.parametrize("nq", [3, 40])  # both FAISS branches: scalar heap (<20) and blocked sgemm (>=20)
def test_c_heap_restatement_agrees_with_numpy_restatement(nq):
    P = synth.block(100, 30000, seed=2)
    q = synth.block(0, nq, seed=2, stream=1)
    D1, I1 = c_oracle.knn_ip_heap(q, P, 50)
    D2, I2 = flat_ip.knn_inner_product(q, P, 50)
    Dt, It = flat_ip.truth_fp64(q, P, 50)
    score_of = lambda qi, ids: (q[qi].astype(np.float64) @ P[ids].astype(np.float64).T)
    r = flat_ip.compare(D1, I1, D2, I2, score_of)
    assert r["violations"] == 0
    r = flat_ip.compare(D2, I2, Dt, It, score_of)
    assert r["violations"] == 0 and r["max_rel_score_err"] < 1e-5


def test_fp32_oracle_vs_fp64_truth_config1_shape_reduced():
    # BASELINE config 1 (100k x 768, 173 queries, top-100) at 1/5 of the rows to keep the CPU suite fast
    P = c_oracle.synth_block(0, 20000)
    q = c_oracle.synth_block(0, 173, stream=1)
    D, I = flat_ip.knn_inner_product(q, P, 100)
    Dt, It = flat_ip.truth_fp64(q, P, 100)
    score_of = lambda qi, ids: (q[qi].astype(np.float64) @ P[ids].astype(np.float64).T)
    r = flat_ip.compare(D, I, Dt, It, score_of)
    assert r["violations"] == 0, r
    assert r["max_rel_score_err"] < 1e-5


def test_non_unit_norm_scale_independence():
    P = synth.block(0, 5000, norm=28.0)
    q = synth.block(0, 6, stream=1, norm=28.0)
    Pu = synth.block(0, 5000)
    qu = synth.block(0, 6, stream=1)
    _, I = flat_ip.knn_inner_product(q, P, 30)
    _, Iu = flat_ip.knn_inner_product(qu, Pu, 30)
    assert (I == Iu).mean() > 0.98  # same ranking up to fp32 rounding of the scaled rows
    np.testing.assert_allclose(np.linalg.norm(P[:10], axis=1), 28.0, rtol=1e-5)


def test_comparator_flags_real_differences_and_excuses_ties():
    D_ref = np.array([[3.0, 2.0, 1.0]])
    I_ref = np.array([[7, 8, 9]])
    assert flat_ip.compare(D_ref, I_ref, D_ref, I_ref)["exact_rows"] == 1
    # swapped ids with clearly different scores -> violation
    r = flat_ip.compare(np.array([[3.0, 2.0, 1.0]]), np.array([[8, 7, 9]]), D_ref, I_ref)
    assert r["violations"] > 0
    # a tie within 1e-5 relative may swap
    Dt = np.array([[3.0, 2.00001, 2.0]])
    r = flat_ip.compare(Dt[:, [0, 2, 1]][:, [0, 2, 1]], np.array([[7, 9, 8]]), Dt, np.array([[7, 8, 9]]),
                        score_of=lambda q, ids: np.array([{7: 3.0, 8: 2.00001, 9: 2.0}[int(i)] for i in ids]))
    assert r["violations"] == 0 and r["tie_excused"] == 2
    # score off by 1e-3 relative -> violation
    r = flat_ip.compare(np.array([[3.003, 2.0, 1.0]]), I_ref, D_ref, I_ref)
    assert r["violations"] == 1


def test_synth_host_twins_agree_bit_for_bit():
    a = c_oracle.synth_block(12345678901, 257, seed=9, stream=4, norm=3.5)
    b = synth.block(12345678901, 257, seed=9, stream=4, norm=3.5)
    np.testing.assert_array_equal(a, b)
    z = np.zeros(1, dtype=np.uint64)
    kat = [int(x[0]) for x in synth.philox4x32_10(z, z, z, z, 0, 0)]
    assert kat == [0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8]  # Random123 known-answer vector


def test_eval_rank_dedup_keeps_first_occurrences_in_rank_order():
    # reference drivers/run_convdr_inference.py:37-69: two offsets of one passage -> the better rank stays
    offset2pid = np.array([10, 11, 10, 12, 13, 11], dtype=np.int64)
    I = np.array([[0, 2, 1, 5, 3, 4], [4, 3, -1, 5, 0, 2]], dtype=np.int64)
    D = np.array([[.9, .8, .7, .6, .5, .4], [.9, .8, .7, .6, .5, .4]])
    pids, scores, counts = flat_ip.eval_rank_dedup(D, I, 5, offset2pid)
    assert pids[0].tolist() == [10, 11, 12, 0, 0] and counts[0] == 3          # only the first 5 entries are read
    assert scores[0].tolist() == [.9, .7, .5, 0.0, 0.0]
    assert pids[1].tolist() == [13, 12, 11, 10, 0] and counts[1] == 4         # -1 wraps to the last offset (pid 11)


def test_full_size_oracle_equals_materialised_truth_and_synth_twins_agree_with_mean_shift():
    """oracle_topk_synth_f64 (rows regenerated block by block, never materialised) == truth_fp64 over the
    materialised rows, for isotropic rows and for rows with a common mean; host twins bit-identical."""
    from convdr_b200 import synth
    for kw in (dict(), dict(norm=28.0, mean_shift=443)):
        P = c_oracle.synth_block(1000, 30000, seed=5, **kw)
        np.testing.assert_array_equal(P[:500], synth.block(1000, 500, seed=5, **kw))
        Q = c_oracle.synth_block(0, 6, seed=5, stream=1, **kw)
        D, I = c_oracle.topk_synth_f64(Q, 25, 1000, 30000, seed=5, **kw)
        Dt, It = flat_ip.truth_fp64(Q, P, 25)
        np.testing.assert_array_equal(I, It + 1000)
        np.testing.assert_allclose(D, Dt, rtol=1e-12)
    D, I = c_oracle.topk_synth_f64(Q, 10, 0, 4)           # fewer rows than k: padded like the index
    assert (I[:, 4:] == -1).all() and np.isneginf(D[:, 4:]).all()
