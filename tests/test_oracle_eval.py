"""CPU: pin the evaluator oracle (numpy + C) to the real reference's outputs."""
import os

import numpy as np
import pytest

from grl_b200 import synth
from oracle import eval_oracle as eo
from helpers import c_evaluate, load_c_oracle

def fixture_ids(g):
    return synth.make_eval_set(int(g["nq"]), int(g["ng_extra"]), int(g["dim"]), seed=int(g["seed"]), num_ids=25,
                               noise=float(g["noise"]), missing_query_frac=0.05)


@pytest.mark.parametrize("name", ["eval_small", "eval_rank10", "eval_ties"])
def test_eval_oracle_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    qf, gf, qp, gp, qc, gc = fixture_ids(g)
    mr = int(g["max_rank"])
    ties = int(g["quantize"]) > 0
    if not ties:
        assert np.allclose(eo.cosin_dist(qf, gf), g["d_cos"], atol=2e-6)
    # near-zero self distances are sqrt(rounding noise): compare squared distances
    assert np.allclose(eo.pairwise_distance(qf, gf) ** 2, g["d_l2"] ** 2, atol=5e-6)
    # literal port on the reference's own distance matrix: bit-identical
    cmc, mAP = eo.evaluate_literal(g["d_cos"], qp, gp, qc, gc, mr)
    assert np.array_equal(cmc, g["cmc"]) and mAP == float(g["mAP"])
    cmc2, mAP2 = eo.evaluate_literal(g["d_l2"], qp, gp, qc, gc, mr)
    assert np.array_equal(cmc2, g["cmc_l2"]) and mAP2 == float(g["mAP_l2"])
    assert eo.evaluate_seq(g["d_cos"], qp, qc, gp, gc)[0] == float(g["rank1"]) or mr != 100
    if not ties:
        # sort-free formulation (what the kernel does) and the C restatement agree with the reference
        cmc3, mAP3, ap, first = eo.evaluate_rankcount(g["d_cos"], qp, gp, qc, gc, mr)
        assert np.array_equal(cmc3, g["cmc"]) and abs(mAP3 - float(g["mAP"])) < 1e-12
        lib = load_c_oracle()
        cmc4, mAP4, nv, ap4 = c_evaluate(lib, g["d_cos"], qp, gp, qc, gc, mr)
        assert np.array_equal(cmc4, g["cmc"]) and abs(mAP4 - float(g["mAP"])) < 1e-12
        assert np.allclose(ap4, ap, atol=1e-12)


def test_rankcount_equals_stable_sort_with_ties(golden_dir):
    g = np.load(os.path.join(golden_dir, "eval_ties.npz"))
    _, _, qp, gp, qc, gc = fixture_ids(g)
    cmc_s, mAP_s = eo.evaluate_literal(g["d_cos"], qp, gp, qc, gc, 100, kind="stable")
    cmc_r, mAP_r, _, _ = eo.evaluate_rankcount(g["d_cos"], qp, gp, qc, gc, 100)
    assert np.array_equal(cmc_s, cmc_r) and abs(mAP_s - mAP_r) < 1e-12
    lib = load_c_oracle()
    cmc_c, mAP_c, _, _ = c_evaluate(lib, g["d_cos"], qp, gp, qc, gc, 100)
    assert np.array_equal(cmc_c, cmc_r) and abs(mAP_c - mAP_r) < 1e-12


def test_topk_merge_is_shard_count_invariant():
    rng = np.random.default_rng(0)
    d = rng.standard_normal((17, 400)).astype(np.float32)
    d[:, ::7] = d[:, 1:2]            # ties across shards
    v_ref, i_ref = eo.topk_stable(d, 10)
    for shards in (1, 2, 4, 8):
        w = 400 // shards
        vs, is_ = zip(*[eo.topk_stable(d[:, s * w:(s + 1) * w], 10, idx_base=s * w) for s in range(shards)])
        v, i = eo.merge_topk(list(vs), list(is_), 10)
        assert np.array_equal(v, v_ref) and np.array_equal(i, i_ref)


@pytest.mark.parametrize("metric", [0, 1])
@pytest.mark.parametrize("dup", [False, True])
def test_two_stage_search_with_skipped_rescores_is_exact(metric, dup):
    """The re-score skip rule of the fused search (grl_b200/csrc/search.cu: rescore_keys_kernel), restated in numpy: dropping
    the candidates whose coarse distance exceeds the coarse k-th by more than 2 E_max never changes the exact top-k -- on
    random data (where it drops a good part of the K' candidates) and on a gallery of near-duplicates (where the proof
    fails for many queries and the brute-force leg takes over)."""
    rng = np.random.default_rng(11 + metric)
    q = rng.standard_normal((24, 64)).astype(np.float32)
    g = rng.standard_normal((900, 64)).astype(np.float32)
    if dup:
        g[100:500] = g[7] + 1e-4 * rng.standard_normal((400, 64)).astype(np.float32)
        q[:6] = g[7] + 1e-3 * rng.standard_normal((6, 64)).astype(np.float32)
    k, kp = 10, 64
    d_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, metric), k)
    d, i, flags, n_with = eo.two_stage_topk(q, g, k, kp, metric, skip=True)
    assert np.array_equal(i, i_ref) and np.array_equal(d, d_ref)
    d2, i2, flags2, n_without = eo.two_stage_topk(q, g, k, kp, metric, skip=False)
    assert np.array_equal(i2, i_ref) and np.array_equal(flags, flags2)      # skipping never changes a proof either
    assert n_with < n_without                                               # ... and it does skip work
    if dup:
        assert flags[:6].all()


def _props_inputs(g):
    qf, gf, qp, gp, qc, gc = synth.make_eval_set(int(g["nq"]), int(g["ng_extra"]), int(g["dim"]), seed=int(g["seed"]), num_ids=25,
                                                 noise=float(g["noise"]), missing_query_frac=0.05)
    gf = gf + np.float32(1e-3) * np.random.default_rng(int(g["seed"])).standard_normal(gf.shape).astype(np.float32)
    return qf, gf, qp, gp, qc, gc


def test_evaluate_agrees_with_the_references_other_evaluators(golden_dir):
    """the reference carries two more evaluators that the live path never calls -- cmc(first_match_break=True)
    (eva_functions.py:18-78) and mean_ap (:81-115, sklearn's average_precision_score).  On a tie-free matrix they must agree
    with `evaluate`; the fixture holds all three as computed by the REAL reference, the oracle (both forms) must match them."""
    g = np.load(os.path.join(golden_dir, "eval_props.npz"))
    qf, gf, qp, gp, qc, gc = _props_inputs(g)
    topk = int(g["topk"])
    assert np.abs(g["cmc_evaluate"] - g["cmc_first_match_break"]).max() < 1e-6 and abs(float(g["mAP_evaluate"]) - float(g["mAP_sklearn"])) < 1e-12
    d = eo.cosin_dist(qf, gf)
    cmc, mAP = eo.evaluate_literal(d, qp, gp, qc, gc, topk)
    assert np.abs(cmc - g["cmc_first_match_break"]).max() < 1e-6 and abs(mAP - float(g["mAP_sklearn"])) < 1e-9
    cmc2, mAP2, _, _ = eo.evaluate_rankcount(d, qp, gp, qc, gc, topk)
    assert np.abs(cmc2 - g["cmc_first_match_break"]).max() < 1e-6 and abs(mAP2 - float(g["mAP_sklearn"])) < 1e-9
    # gallery-permutation invariance (tie-free): the metrics depend on the set of gallery rows, not on their order
    perm = np.random.default_rng(0).permutation(gf.shape[0])
    cmc3, mAP3, _, _ = eo.evaluate_rankcount(d[:, perm], qp, gp[perm], qc, gc[perm], topk)
    assert np.array_equal(cmc3, cmc2) and abs(mAP3 - mAP2) < 1e-12


def _prefilter_pass_count(row, kp):
    """numpy restatement of the cut of boot_select_sampled (grl_b200/csrc/search.cu): one value in eight of the row (thread t of
    256: columns 4t, 2048+4t+1, 4096+4t+2, 6144+4t+3) is histogrammed over 2,048 ordered bins between the sample's extremes;
    t0 = the largest sample in the bins up to the one holding sample rank ceil((kp + 4.5 sqrt(8 kp) + 16) / 8).  Returns how many
    of the row's 8,192 values pass `value <= t0` (the kernel then selects exactly among those, or falls back to its exact path
    when fewer than kp or more than 2,048 pass)."""
    b = np.ascontiguousarray(row, dtype=np.float32).view(np.uint32)
    v = np.where(b & np.uint32(0x80000000), ~b, b | np.uint32(0x80000000)).astype(np.uint32)       # order-preserving bits (common.cuh: orderable)
    t = np.arange(256)
    sv = v[np.concatenate([4 * t, 2048 + 4 * t + 1, 4096 + 4 * t + 2, 6144 + 4 * t + 3])]
    lo, hi = sv.min(), sv.max()
    if lo == hi:
        return None
    scale = np.float32(2047) / np.float32(hi - lo)
    bins = np.minimum(2047, ((sv - lo).astype(np.float32) * scale).astype(np.uint32))
    rank = (kp + int(np.float32(4.5) * np.sqrt(np.float32(kp * 8))) + 16 + 7) // 8
    edge = np.searchsorted(np.cumsum(np.bincount(bins, minlength=2048)), rank)
    return int((v <= sv[bins <= edge].max()).sum())


@pytest.mark.parametrize("kp", [256, 512, 1024])
def test_first_chunk_prefilter_cut_keeps_enough_and_not_too_many(kp):
    """The sampled cut of the first-chunk selection is a work-saving guess, never a correctness condition (rows it misjudges take
    the exact path) -- but it has to be a GOOD guess: on continuous data at least kp and at most 2,048 of the 8,192 values must
    pass it for (nearly) every row.  4,000 Gaussian and 4,000 uniform rows per list length: no row may fall outside."""
    rng = np.random.default_rng(kp)
    counts = [_prefilter_pass_count(rng.standard_normal(8192), kp) for _ in range(4000)]
    counts += [_prefilter_pass_count(rng.random(8192), kp) for _ in range(4000)]
    counts = np.array(counts)
    assert counts.min() >= kp and counts.max() <= 2048, (counts.min(), counts.max())
    assert 1.3 * kp < counts.mean() < 2.4 * kp, counts.mean()
    assert _prefilter_pass_count(np.full(8192, 0.25), kp) is None            # constant row: no cut, exact path
