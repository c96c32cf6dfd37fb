"""GPU parity: matching / evaluation kernels through the C ABI vs the oracle (tests only)."""
import ctypes
import os
import time

import numpy as np
import pytest
import torch

from grl_b200 import synth

pytestmark = pytest.mark.gpu


def _mods():
    from grl_b200 import _lib, evaluator
    return _lib, evaluator


def assert_same_ranking_up_to_ties(d_got, d_ref, tol=1e-5):
    """Rankings must be identical except where the reference distances tie within `tol` (north-star)."""
    o_got = np.argsort(d_got, axis=1, kind="stable")
    o_ref = np.argsort(d_ref, axis=1, kind="stable")
    diff = o_got != o_ref
    if diff.any():
        rows, cols = np.nonzero(diff)
        a = d_ref[rows, o_got[rows, cols]]
        b = d_ref[rows, o_ref[rows, cols]]
        assert np.all(np.abs(a - b) <= tol), "ranking differs beyond ties: max gap %g" % np.abs(a - b).max()


@pytest.mark.parametrize("nq,ng,dim", [(64, 300, 64), (200, 900, 2048), (130, 517, 6144), (33, 70, 100)])
def test_distances_match_oracle(nq, ng, dim):
    _, ev = _mods()
    from oracle import eval_oracle as eo
    rng = np.random.default_rng(nq + ng)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    g = rng.standard_normal((ng, dim)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    d = ev.cosin_dist(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda()).cpu().numpy()
    ref64 = -(q.astype(np.float64) @ g.astype(np.float64).T)
    assert np.abs(d - ref64).max() < 5e-6          # split-bf16: ~2^-17 relative per product; well inside the 1e-5 tie window
    assert_same_ranking_up_to_ties(d, eo.cosin_dist(q, g))
    l2 = ev.pairwise_distance_tensor(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda()).cpu().numpy()
    ref_l2 = eo.pairwise_distance(q, g)
    # squared distances are ~2: fp32 norms + sqrt/square round trip cost a few ulp (2.4e-7 each); compare with fp64 truth
    ref_l2_64 = eo.pairwise_distance(q.astype(np.float64), g.astype(np.float64))
    assert np.abs(l2.astype(np.float64) ** 2 - ref_l2_64 ** 2).max() < 1e-5
    assert_same_ranking_up_to_ties(l2, ref_l2, tol=2e-5)


@pytest.mark.parametrize("name", ["eval_small", "eval_rank10", "eval_ties"])
def test_evaluate_matches_reference_golden(golden_dir, name):
    _, ev = _mods()
    from oracle import eval_oracle as eo
    g = np.load(os.path.join(golden_dir, name + ".npz"))
    qf, gf, qp, gp, qc, gc = synth.make_eval_set(int(g["nq"]), int(g["ng_extra"]), int(g["dim"]), seed=int(g["seed"]),
                                                 num_ids=25, noise=float(g["noise"]), missing_query_frac=0.05)
    mr = int(g["max_rank"])
    cmc, mAP = ev.evaluate(g["d_cos"], qp, gp, qc, gc, mr)
    assert cmc.dtype == np.float32 and cmc.shape == (mr,)
    if int(g["quantize"]) == 0:
        assert np.array_equal(cmc, g["cmc"]) and abs(mAP - float(g["mAP"])) < 1e-12
        # end to end from features: distance kernel + metric kernel
        d = ev.cosin_dist(torch.from_numpy(qf).cuda(), torch.from_numpy(gf).cuda())
        cmc2, mAP2 = ev.evaluate(d, qp, gp, qc, gc, mr)
        assert np.abs(cmc2 - g["cmc"]).max() < 1e-4 and abs(mAP2 - float(g["mAP"])) < 1e-4
    else:   # exact ties: the kernel is the stable-sort answer; numpy's introsort order is implementation-defined
        cmc_s, mAP_s = eo.evaluate_literal(g["d_cos"], qp, gp, qc, gc, mr, kind="stable")
        assert np.array_equal(cmc, cmc_s) and abs(mAP - mAP_s) < 1e-12


def test_evaluate_mars_shape_vs_c_oracle():
    _, ev = _mods()
    from helpers import c_evaluate, load_c_oracle
    qf, gf, qp, gp, qc, gc = synth.make_eval_set(1980, 7350, 2048, seed=0, noise=4.0)
    d = ev.cosin_dist(torch.from_numpy(qf).cuda(), torch.from_numpy(gf).cuda())
    cmc, mAP = ev.evaluate(d, qp, gp, qc, gc)
    dn = d.cpu().numpy()
    cmc_o, mAP_o, nv, ap_o = c_evaluate(load_c_oracle(), dn, qp, gp, qc, gc)
    assert np.array_equal(cmc, cmc_o) and abs(mAP - mAP_o) < 1e-12
    assert 0.05 < mAP < 0.999          # the synthetic set is neither trivial nor random
    # against the fp32 CPU distance matrix (the reference's own arithmetic): within 1e-4
    from oracle import eval_oracle as eo
    cmc_r, mAP_r, _, _ = c_evaluate(load_c_oracle(), eo.cosin_dist(qf, gf), qp, gp, qc, gc)
    assert np.abs(cmc - cmc_r).max() < 1e-4 and abs(mAP - mAP_r) < 1e-4


def test_evaluate_many_positives_and_no_valid_query():
    _, ev = _mods()
    from oracle import eval_oracle as eo
    rng = np.random.default_rng(1)
    nq, ng = 7, 3000
    d = rng.standard_normal((nq, ng)).astype(np.float32)
    qp = np.ones(nq, np.int64); qc = np.zeros(nq, np.int64)
    gp = np.ones(ng, np.int64); gc = rng.integers(0, 3, ng)      # ~2000 positives per query: several smem rounds
    cmc, mAP = ev.evaluate(d, qp, gp, qc, gc, 50)
    cmc_o, mAP_o, _, _ = eo.evaluate_rankcount(d, qp, gp, qc, gc, 50)
    assert np.array_equal(cmc, cmc_o) and abs(mAP - mAP_o) < 1e-12
    with pytest.raises(AssertionError):
        ev.evaluate(d, qp + 5, gp, qc, gc)


def test_argsort_rows_is_stable_argsort():
    _, ev = _mods()
    rng = np.random.default_rng(2)
    for nq, ng in ((5, 9330), (3, 1), (4, 1000)):
        d = rng.standard_normal((nq, ng)).astype(np.float32)
        d[:, ::5] = np.round(d[:, ::5])        # ties
        if ng >= 3:
            d[0, :3] = [0.0, -0.0, 0.0]
        order = ev.argsort_rows(torch.from_numpy(d).cuda()).cpu().numpy()
        assert np.array_equal(order, np.argsort(d, axis=1, kind="stable"))


def test_topk_stream_and_merge_shard_invariant():
    _lib, ev = _mods()
    from oracle import eval_oracle as eo
    lib = _lib.load_library(); h = _lib.get_handle()
    rng = np.random.default_rng(3)
    nq, ng, k = 37, 20000, 100
    d = rng.standard_normal((nq, ng)).astype(np.float32)
    d[:, ::9] = d[:, 4:5]
    v_ref, i_ref = eo.topk_stable(d, k)
    dd = torch.from_numpy(d).cuda()
    st = _lib.stream_ptr()

    def run_shard(lo, hi, chunk):
        td = torch.empty((nq, k), device="cuda"); ti = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        _lib.check(h, lib.grl_topk_init(h, td.data_ptr(), ti.data_ptr(), nq, k, st), "init")
        for c0 in range(lo, hi, chunk):
            c1 = min(hi, c0 + chunk)
            tile = dd[:, c0:c1].contiguous()
            _lib.check(h, lib.grl_topk_rows(h, tile.data_ptr(), tile.stride(0), nq, c1 - c0, k, c0, td.data_ptr(),
                                            ti.data_ptr(), st), "topk")
        return td, ti

    for shards in (1, 2, 4, 8):
        w = ng // shards
        parts = [run_shard(s * w, (s + 1) * w, 1777) for s in range(shards)]
        all_d = torch.stack([p[0] for p in parts]).contiguous(); all_i = torch.stack([p[1] for p in parts]).contiguous()
        od = torch.empty((nq, k), device="cuda"); oi = torch.empty((nq, k), dtype=torch.int64, device="cuda")
        _lib.check(h, lib.grl_topk_merge(h, all_d.data_ptr(), all_i.data_ptr(), shards, nq, k, od.data_ptr(), oi.data_ptr(), st), "merge")
        assert np.array_equal(od.cpu().numpy(), v_ref) and np.array_equal(oi.cpu().numpy(), i_ref)


def _retrieval_inputs(nq, ng, dim, seed, dup_every=13):
    rng = np.random.default_rng(seed)
    q = rng.standard_normal((nq, dim)).astype(np.float32)
    g = rng.standard_normal((ng, dim)).astype(np.float32)
    q /= np.linalg.norm(q, axis=1, keepdims=True)               # unit-norm descriptors like the reference's features
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    if dup_every:
        g[::dup_every] = g[5]                                   # exact ties, also across shard boundaries
    return q, g


@pytest.mark.parametrize("metric", [0, 1])
def test_dist_topk_shard_invariant_vs_oracle(metric):
    """grl_dist_topk per shard (coarse fp16 pass + exact re-score + completeness proof, brute force where the proof fails)
    + grl_topk_merge == stable top-k of the fixed-order fp32 distance matrix, bit for bit, for 1/2/4/8 shards."""
    _, ev = _mods()
    from oracle import eval_oracle as eo
    nq, ng, dim, k = 41, 20011, 128, 100
    q, g = _retrieval_inputs(nq, ng, dim, 11)                   # ~1540 duplicated rows: queries near them fail the proof
    qd, gd = torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda()
    v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, metric), k)
    for shards in (1, 2, 4, 8):
        parts = []
        for r in range(shards):
            lo, n = ev.shard_bounds(ng, shards, r)
            parts.append(ev.retrieve_topk(qd, gd[lo:lo + n], k, idx_base=lo, metric=metric))
        od, oi = ev.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
        assert np.array_equal(oi.cpu().numpy(), i_ref), shards
        assert np.array_equal(od.cpu().numpy(), v_ref), shards
    # and against the CPU fp32 arithmetic of the reference: identical ranking except at ties within 1e-5
    ref = eo.cosin_dist(q, g) if metric == 0 else eo.pairwise_distance(q, g)
    got_d = np.take_along_axis(ref, oi.cpu().numpy(), 1)
    best = np.sort(ref, axis=1)[:, :k]
    assert np.abs(got_d - best).max() <= 2e-5


@pytest.mark.parametrize("metric,dup_every,k", [(0, 13, 100), (1, 0, 100), (0, 0, 300), (1, 13, 17)])
def test_staged_search_over_simulated_shards(metric, dup_every, k):
    """The stages sharded_retrieve runs on every rank, driven here for 1/3/8 shards on one GPU with the collectives
    replaced by stack / max / sum: coarse lists -> merge -> owned re-scores summed -> finalize -> brute force of flagged rows."""
    _, ev = _mods()
    from oracle import eval_oracle as eo
    st = ev.CudaSearchStages
    nq, ng, dim = 53, 9001, 256
    q, g = _retrieval_inputs(nq, ng, dim, 21 + k, dup_every)
    qd, gd = torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda()
    v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, metric), k)
    kp = st.kprime(k)
    n_flagged = []
    for shards in (1, 3, 8):
        spans = [ev.shard_bounds(ng, shards, r) for r in range(shards)]
        coarse = [st.coarse(qd, gd[lo:lo + n].contiguous(), kp, lo, metric) for lo, n in spans]
        gmax2 = torch.stack([c[2] for c in coarse]).max(dim=0).values
        dirty = torch.stack([c[3] for c in coarse]).max(dim=0).values
        cd, ci = st.merge(torch.stack([c[0] for c in coarse]), torch.stack([c[1] for c in coarse]))
        ed = sum(st.rescore(qd, gd[lo:lo + n].contiguous(), ci, lo, metric) for lo, n in spans)
        top_d, top_i, flags = st.finalize(qd, cd, ci, ed, gmax2, dirty, k, metric)
        rows = torch.nonzero(flags).flatten()
        n_flagged.append(int(rows.numel()))
        if rows.numel():
            ex = [st.exact(qd[rows].contiguous(), gd[lo:lo + n].contiguous(), k, lo, metric) for lo, n in spans]
            d_x, i_x = st.merge(torch.stack([e[0] for e in ex]), torch.stack([e[1] for e in ex]))
            top_d[rows] = d_x
            top_i[rows] = i_x
        assert np.array_equal(top_i.cpu().numpy(), i_ref), (shards, n_flagged)
        assert np.array_equal(top_d.cpu().numpy(), v_ref), (shards, n_flagged)
        # the coarse ranking must honour its error bound against the exact distances (what the proof relies on)
        own = ci >= 0
        c_exact = torch.where(own, ed, torch.zeros_like(ed))
        c_coarse = torch.where(own, cd, torch.zeros_like(cd))
        if metric == 1:
            c_exact = c_exact ** 2
        bound = eo.coarse_error_constant(dim) * (2.0 if metric == 1 else 1.0) * 1.01 + 1e-5
        assert float((c_exact - c_coarse).abs().max()) <= bound
    if dup_every == 0 and k <= 100:
        assert n_flagged == [0, 0, 0], n_flagged                 # generic data: every proof succeeds, no brute force
    if dup_every and k == 100:
        assert n_flagged[0] > 0                                  # duplicated rows straddling K': the brute-force leg ran


@pytest.mark.parametrize("metric", [0, 1])
def test_dist_topk_large_query_block(metric):
    """nq >= 1024 and shards >= 1024 rows: the 256 x 256-tile coarse GEMM (coarse_gemm.cuh), ragged last tiles included."""
    _, ev = _mods()
    from oracle import eval_oracle as eo
    nq, ng, dim, k = 1100, 3333, 64, 20
    q, g = _retrieval_inputs(nq, ng, dim, 31, dup_every=0)
    qd, gd = torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda()
    v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, metric), k)
    for shards in (1, 2):
        parts = []
        for r in range(shards):
            lo, n = ev.shard_bounds(ng, shards, r)
            parts.append(ev.retrieve_topk(qd, gd[lo:lo + n], k, idx_base=lo, metric=metric))
        od, oi = ev.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
        assert np.array_equal(oi.cpu().numpy(), i_ref), shards
        assert np.array_equal(od.cpu().numpy(), v_ref), shards


@pytest.mark.parametrize("nq", [8, 1100])
def test_dist_topk_adversarial_order_overflows_to_brute_force(nq):
    """A gallery sorted so that every later row is closer to query 0 than all earlier ones: after the first (stored) chunk every
    column of every chunk passes query 0's threshold, its candidate buffer overflows with no tile to rescan, the row is marked
    dirty and must come back exact from the brute-force leg.  nq = 8: gemm.cuh single-plane kernel; 1100: coarse_gemm.cuh."""
    _, ev = _mods()
    from oracle import eval_oracle as eo
    ng, dim, k = 20000, 64, 30                        # > 8192 columns: the first (stored, exactly selected) chunk does not cover the gallery
    q, g = _retrieval_inputs(nq, ng, dim, 77, dup_every=0)
    g = g[np.argsort(-(g @ q[0]), kind="stable")[::-1].copy()]            # ascending similarity to q[0] == descending distance
    qd, gd = torch.from_numpy(q).cuda(), torch.from_numpy(np.ascontiguousarray(g)).cuda()
    for metric in (0, 1):
        st = ev.CudaSearchStages
        cd, ci, gmax2, dirty = st.coarse(qd, gd, st.kprime(k), 0, metric)
        assert int(dirty[0]) == 1 and int(dirty.sum()) < max(2, nq // 10)
        d, i = ev.retrieve_topk(qd, gd, k, metric=metric)
        v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, metric), k)
        assert np.array_equal(i.cpu().numpy(), i_ref) and np.array_equal(d.cpu().numpy(), v_ref)


@pytest.mark.parametrize("dups", ["all", "cycle3", 1500, 500, 100, 0])
def test_first_chunk_selection_with_duplicate_rows(dups):
    """The first chunk's selection (list_boot_select_kernel: sampled pre-filter, exact bin search, ordered scans for equal values)
    on galleries full of EXACT ties inside a full 8192-column chunk: every gallery row identical; three rows repeated; 1500 /
    500 / 100 copies of every query's nearest row scattered among random rows (pre-filter passes too many -> exact path;
    boundary bin sorted; boundary bin ranked); and no duplicates.  Ties must come back lowest index first, like the stable sort
    of the reference (eva_functions.py:73)."""
    _, ev = _mods()
    from oracle import eval_oracle as eo
    nq, ng, dim, k = 8, 9000, 64, 30
    q, g = _retrieval_inputs(nq, ng, dim, 311, dup_every=0)
    rng = np.random.default_rng(5)
    if dups == "all":
        g[:] = g[7]
    elif dups == "cycle3":
        g[:] = g[np.arange(ng) % 3]
    elif dups:
        near = q.mean(axis=0)
        g[rng.permutation(8192)[:dups]] = near / np.linalg.norm(near)       # closer to every query than any random row
    qd, gd = torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda()
    for metric in (0, 1):
        d, i = ev.retrieve_topk(qd, gd, k, metric=metric)
        v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, metric), k)
        assert np.array_equal(i.cpu().numpy(), i_ref) and np.array_equal(d.cpu().numpy(), v_ref), (dups, metric)
    # the selection alone: a single full chunk -> the coarse list is the K' smallest (coarse distance, index) keys, ascending
    st = ev.CudaSearchStages
    cd, ci, _, dirty = st.coarse(qd, gd[:8192], st.kprime(k), 0, 0)
    cd, ci = cd.cpu().numpy(), ci.cpu().numpy()
    assert int(dirty.sum()) == 0
    assert all(len(set(r.tolist())) == r.size for r in ci) and bool((np.diff(cd, axis=1) >= 0).all())
    same = np.diff(cd, axis=1) == 0
    assert bool((np.diff(ci, axis=1)[same] > 0).all())                    # equal coarse distances: ascending index
    if dups == "all":
        assert np.array_equal(ci, np.tile(np.arange(ci.shape[1]), (nq, 1)))


@pytest.mark.parametrize("nlists,kp,rows", [(2, 256, 37), (3, 256, 5), (4, 512, 9), (8, 256, 50), (8, 1024, 3), (16, 1024, 2), (5, 2, 4), (1, 256, 3)])
def test_merge_key_lists_is_the_sorted_union_prefix(nlists, kp, rows):
    """grl_merge_key_lists (the exchange step of grl_sharded_topk: a merge tree over the shards' ascending key lists) ==
    the kp smallest of the union, ascending -- for 1..16 lists, list counts that are not powers of two, lists with empty
    (all-ones) tails and keys shared between lists."""
    import ctypes as C
    _lib, _ = _mods()
    rng = np.random.default_rng(nlists * 1000 + kp)
    keys = rng.integers(0, 2**63, size=(nlists, rows, kp), dtype=np.uint64) * np.uint64(2) + rng.integers(0, 2, size=(nlists, rows, kp), dtype=np.uint64)
    keys[:, :, : kp // 4] = keys[0:1, :, : kp // 4]                      # keys present in every list
    fill = rng.integers(0, kp + 1, size=(nlists, rows))
    for s_ in range(nlists):
        for r in range(rows):
            keys[s_, r, kp - fill[s_, r] // 3:] = np.uint64(0xFFFFFFFFFFFFFFFF)           # an empty tail
    keys.sort(axis=2)
    want = np.sort(keys.transpose(1, 0, 2).reshape(rows, nlists * kp), axis=1)[:, :kp]
    lists = torch.from_numpy(keys.view(np.int64)).cuda()
    out = torch.empty((rows, kp), dtype=torch.int64, device="cuda")
    lib = _lib.load_library()
    h = _lib.get_handle(out.device)
    _lib.check(h, lib.grl_merge_key_lists(h, lists.data_ptr(), nlists, rows, kp, out.data_ptr(), _lib.stream_ptr(out.device)), "grl_merge_key_lists")
    assert np.array_equal(out.cpu().numpy().view(np.uint64), want)


def test_exact_topk_brute_force_matches_oracle():
    _, ev = _mods()
    from oracle import eval_oracle as eo
    for metric in (0, 1):
        q, g = _retrieval_inputs(19, 3001, 72, 5 + metric)
        d, i = ev.CudaSearchStages.exact(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda(), 50, 1000, metric)
        v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, metric), 50, 1000)
        assert np.array_equal(i.cpu().numpy(), i_ref) and np.array_equal(d.cpu().numpy(), v_ref)

@pytest.mark.parametrize("nq,ng,dim,k,metric", [(1, 50, 8, 1, 0), (3, 40, 100, 64, 1), (17, 700, 72, 512, 0), (5, 3000, 30, 7, 1),
                                                (2, 5, 16, 5, 0)])
def test_dist_topk_edge_shapes(nq, ng, dim, k, metric):
    """One query, k = 1, k = 512 (K' = 1024), k > ng (lists padded with +inf / -1), feature sizes that are not multiples
    of 8 (zero-padded by the wrapper: the distance definition must not change)."""
    _, ev = _mods()
    from oracle import eval_oracle as eo
    q, g = _retrieval_inputs(nq, ng, dim, 100 + nq + k, dup_every=0)
    d, i = ev.retrieve_topk(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda(), k, metric=metric)
    pad = (-dim) % 8
    qp, gp = np.pad(q, ((0, 0), (0, pad))), np.pad(g, ((0, 0), (0, pad)))
    v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(qp, gp, metric), k)
    kk = min(k, ng)
    assert np.array_equal(i.cpu().numpy()[:, :kk], i_ref) and np.array_equal(d.cpu().numpy()[:, :kk], v_ref)
    if k > ng:
        assert bool((i[:, ng:] == -1).all()) and bool(torch.isinf(d[:, ng:]).all())


def test_retrieval_rejects_bad_arguments():
    _, ev = _mods()
    q = torch.zeros((4, 16), device="cuda")
    with pytest.raises(RuntimeError):
        ev.retrieve_topk(q, torch.zeros((10, 24), device="cuda"), 3)        # feature sizes differ
    with pytest.raises(RuntimeError):
        ev.retrieve_topk(q, torch.zeros((10, 16), device="cuda"), 513)      # k above the supported 512
    with pytest.raises(RuntimeError):
        ev.retrieve_topk(q, torch.zeros((10, 16), device="cuda"), 3, idx_base=2 ** 32)   # global index must fit 32 bits


def test_retrieval_full_size_properties_10k_x_1m():
    """BASELINE configs[4] at full size (10,000 queries x 1,000,000 gallery rows x 2048-d, top-100): too large for the CPU
    oracle, so check what the search guarantees -- lists sorted by (distance, index), indices unique and in range, and for a
    sample of queries bit-identity with the brute-force search in the same arithmetic (grl_exact_topk)."""
    _, ev = _mods()
    nq, ng, dim, k = 10000, 1000000, 2048, 100
    gen = torch.Generator(device="cuda").manual_seed(5)
    g = torch.randn((ng, dim), generator=gen, device="cuda")
    g /= g.norm(dim=1, keepdim=True)
    q = torch.nn.functional.normalize(torch.randn((nq, dim), generator=gen, device="cuda"))
    d, i = ev.retrieve_topk(q, g, k, idx_base=7)
    assert d.shape == (nq, k) and i.shape == (nq, k) and bool(torch.isfinite(d).all())
    assert bool((i >= 7).all()) and bool((i < ng + 7).all())
    assert bool((d[:, 1:] >= d[:, :-1]).all())
    ties = d[:, 1:] == d[:, :-1]
    assert bool((i[:, 1:][ties] > i[:, :-1][ties]).all())
    assert int((torch.sort(i, dim=1).values[:, 1:] == torch.sort(i, dim=1).values[:, :-1]).sum()) == 0
    rows = torch.arange(0, nq, nq // 16, device="cuda")[:16]
    d_x, i_x = ev.CudaSearchStages.exact(q[rows].contiguous(), g, k, 7, 0)
    assert torch.equal(i[rows], i_x) and torch.equal(d[rows], d_x)
    # the distances are the fixed-order fp32 inner products: within 2e-6 of an fp64 evaluation
    ref = -(q[rows].double() @ g[(i[rows] - 7).reshape(-1)].double().T).reshape(16, 16 * k)
    pick = torch.stack([ref[r, r * k:(r + 1) * k] for r in range(16)])
    assert float((d[rows].double() - pick).abs().max()) < 2e-6


@pytest.mark.parametrize("metric", [0, 1])
def test_prepared_gallery_gives_identical_results(metric):
    """A gallery converted once (grl_gallery_prepare) must give bit-identical searches to the per-search conversion, for both
    coarse kernels' shapes (nq >= 1024 and small nq), an unpadded feature size, and across simulated shards."""
    _, ev = _mods()
    for nq, ng, dim, k in ((1100, 5000, 64, 20), (37, 3001, 100, 50)):
        q, g = _retrieval_inputs(nq, ng, dim, 40 + nq, dup_every=11)
        qd, gd = torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda()
        d0, i0 = ev.retrieve_topk(qd, gd, k, metric=metric)
        pg = ev.PreparedGallery(gd)
        for _ in range(2):                                  # reusable
            d1, i1 = ev.retrieve_topk(qd, pg, k, metric=metric)
            assert torch.equal(i0, i1) and torch.equal(d0, d1)
        parts = []
        for r in range(3):
            lo, n = ev.shard_bounds(ng, 3, r)
            parts.append(ev.retrieve_topk(qd, ev.PreparedGallery(gd[lo:lo + n]), k, idx_base=lo, metric=metric))
        dm, im = ev.merge_topk(torch.stack([p[0] for p in parts]), torch.stack([p[1] for p in parts]))
        assert torch.equal(im, i0) and torch.equal(dm, d0)


@pytest.mark.parametrize("k", [100, 200, 512])
def test_sharded_topk_single_rank_stats_and_large_k(k):
    """grl_sharded_topk without a communicator (one rank): exact results, and on generic data NO query may fall back to brute
    force for any supported k -- the per-chunk candidate buffer scales with K' (k = 200 -> K' = 512, k = 512 -> K' = 1024;
    a fixed 512-entry buffer used to overflow there and sent nearly every query to the brute-force leg)."""
    _, ev = _mods()
    from oracle import eval_oracle as eo
    nq, ng, dim = 48, 40000, 64
    q, g = _retrieval_inputs(nq, ng, dim, 300 + k, dup_every=0)
    qd, gd = torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda()
    stats = torch.zeros(8, dtype=torch.int32, device="cuda")
    d, i = ev.sharded_topk(qd, gd, k, 0, stats=stats)
    v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, 0), k)
    assert np.array_equal(i.cpu().numpy(), i_ref) and np.array_equal(d.cpu().numpy(), v_ref)
    flagged, dirty, rescored, skipped = [int(v) for v in stats[:4].cpu()]
    assert flagged == 0 and dirty == 0, (flagged, dirty)
    kp = ev.CudaSearchStages.kprime(k)
    assert rescored + skipped == nq * kp and rescored >= nq * k
    if k == 100:
        assert skipped > 0                                   # the skip rule does skip work on generic data


@pytest.mark.parametrize("metric", [0, 1])
def test_sharded_topk_async_mode_matches_sync(metric):
    """max_flagged >= 0: no host synchronisation inside the call (graph-capturable); the brute-force leg runs for a fixed number
    of row slots gated on the device-side count.  With enough slots the result equals the synchronous one bit for bit; with
    none, the flagged rows are reported in stats[0] and everything else is already final."""
    _, ev = _mods()
    nq, ng, dim, k = 41, 20011, 128, 100
    q, g = _retrieval_inputs(nq, ng, dim, 11)                   # duplicated rows: some proofs fail
    qd, gd = torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda()
    s0 = torch.zeros(8, dtype=torch.int32, device="cuda")
    d0, i0 = ev.sharded_topk(qd, gd, k, 0, metric=metric, stats=s0)
    nflag = int(s0[0])
    assert nflag > 0
    s1 = torch.zeros(8, dtype=torch.int32, device="cuda")
    d1, i1 = ev.sharded_topk(qd, gd, k, 0, metric=metric, max_flagged=nq, stats=s1)
    assert int(s1[0]) == nflag and torch.equal(i1, i0) and torch.equal(d1, d0)
    s2 = torch.zeros(8, dtype=torch.int32, device="cuda")
    d2, i2 = ev.sharded_topk(qd, gd, k, 0, metric=metric, max_flagged=0, stats=s2)
    assert int(s2[0]) == nflag
    same = (i2 == i0).all(dim=1)
    assert int((~same).sum()) <= nflag                          # only flagged rows may still differ
    # the asynchronous form is capturable in a CUDA graph (the synchronous one reads a count on the host)
    out = (torch.empty_like(d0), torch.empty_like(i0))
    side = torch.cuda.Stream()
    side.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(side):
        ev.sharded_topk(qd, gd, k, 0, metric=metric, max_flagged=nq, out=out)
    torch.cuda.current_stream().wait_stream(side)
    torch.cuda.synchronize()
    graph = torch.cuda.CUDAGraph()
    with torch.cuda.graph(graph):
        ev.sharded_topk(qd, gd, k, 0, metric=metric, max_flagged=nq, out=out)
    out[0].zero_(); out[1].zero_()
    graph.replay()
    torch.cuda.synchronize()
    assert torch.equal(out[1], i0) and torch.equal(out[0], d0)


def test_evaluate_sharded_single_rank_equals_evaluate(golden_dir):
    """grl_cmc_map_sharded without a communicator == grl_cmc_map, bit for bit (the split into collect / count / reduce keeps
    the arithmetic, incl. the summation order of the average precision)."""
    _, ev = _mods()
    for seed, nq, extra in ((3, 60, 240), (8, 200, 1500)):
        qf, gf, qp, gp, qc, gc = synth.make_eval_set(nq, extra, 64, seed=seed, num_ids=25, noise=1.5, missing_query_frac=0.05)
        d = ev.cosin_dist(torch.from_numpy(qf).cuda(), torch.from_numpy(gf).cuda())
        d[:, ::7] = d[:, 3:4]                                # ties
        cmc0, map0 = ev.evaluate(d, qp, gp, qc, gc, 50)
        cmc1, map1 = ev.evaluate_sharded(d, qp, gp, qc, gc, 0, 50)
        assert np.array_equal(cmc0, cmc1) and map0 == map1


# (nq, ng, dim, k, duplicated rows): a large block (256 x 256 coarse kernel, brute-force leg), a small ragged one, and a single
# query (the second rank's query slice is EMPTY)
_NCCL_CASES = ((1101, 30011, 128, 100, 13), (37, 9001, 72, 30, 0), (1, 5000, 64, 10, 0))


def _nccl_worker(rank, world, port, out_dir):
    import os
    import torch.distributed as dist
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=torch.device("cuda", rank))
    from grl_b200 import evaluator as ev
    res = {}
    for metric, (nq, ng, dim, k, dup) in enumerate(_NCCL_CASES):
        metric = metric & 1
        q, g = _retrieval_inputs(nq, ng, dim, 500 + metric, dup)
        lo, n = ev.shard_bounds(ng, world, rank)
        qlo, qn = ev.query_slice(nq, world, rank)
        gd = torch.from_numpy(g[lo:lo + n]).cuda()
        stats = torch.zeros(8, dtype=torch.int32, device="cuda")
        # this rank contributes only its slice of the queries; the library all-gathers them
        d, i = ev.sharded_retrieve(torch.from_numpy(q[qlo:qlo + qn]).cuda(), gd, k, lo, metric=metric, nq=nq, stats=stats)
        # ... the same with replicated queries and a prepared gallery
        d2, i2 = ev.sharded_retrieve(torch.from_numpy(q).cuda(), ev.PreparedGallery(gd), k, lo, metric=metric)
        assert torch.equal(d, d2) and torch.equal(i, i2)
        tag = "%d_%d" % (nq, metric)
        res["d" + tag], res["i" + tag], res["s" + tag] = d.cpu().numpy(), i.cpu().numpy(), stats.cpu().numpy()
    # sharded CMC / mAP
    qf, gf, qp, gp, qc, gc = synth.make_eval_set(300, 2500, 64, seed=9, num_ids=25, noise=1.5, missing_query_frac=0.05)
    ngt = gf.shape[0]
    lo, n = ev.shard_bounds(ngt, world, rank)
    dl = ev.cosin_dist(torch.from_numpy(qf).cuda(), torch.from_numpy(gf[lo:lo + n]).cuda())
    cmc, mAP = ev.evaluate_sharded(dl, qp, gp[lo:lo + n], qc, gc[lo:lo + n], lo, 50)
    res["cmc"], res["mAP"] = cmc, np.float64(mAP)
    np.savez(os.path.join(out_dir, "rank%d.npz" % rank), **res)
    dist.barrier()
    ev.destroy_search_comm()
    dist.destroy_process_group()


@pytest.mark.skipif(not torch.cuda.is_available() or torch.cuda.device_count() < 2, reason="needs two GPUs (NCCL ranks)")
def test_sharded_search_and_eval_over_nccl_two_ranks(tmp_path):
    """The real thing: two processes, two GPUs, NCCL.  grl_sharded_topk (query-slice all-gather, all-to-all of the coarse
    lists, owned re-scores + reduce-scatter, sliced finalisation, result all-gather, brute-force leg) and grl_cmc_map_sharded
    give, on both ranks, bit for bit what ONE GPU computes for the whole gallery."""
    import socket
    import torch.multiprocessing as mp
    _, ev = _mods()
    from oracle import eval_oracle as eo
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.spawn(_nccl_worker, args=(2, port, str(tmp_path)), nprocs=2, join=False)
    deadline = time.time() + 150                       # a rank that dies leaves its peer inside a collective: never wait forever
    while not ctx.join(timeout=5):
        if time.time() > deadline:
            for p_ in ctx.processes:
                p_.kill()
            pytest.fail("NCCL workers did not finish within 150 s (one rank failed or the collectives are mismatched)")
    r0, r1 = np.load(str(tmp_path / "rank0.npz")), np.load(str(tmp_path / "rank1.npz"))
    for metric, (nq, ng, dim, k, dup) in enumerate(_NCCL_CASES):
        metric = metric & 1
        tag = "%d_%d" % (nq, metric)
        q, g = _retrieval_inputs(nq, ng, dim, 500 + metric, dup)
        d, i = ev.retrieve_topk(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda(), k, metric=metric)     # one GPU, whole gallery
        for r in (r0, r1):
            assert np.array_equal(r["i" + tag], i.cpu().numpy()) and np.array_equal(r["d" + tag], d.cpu().numpy())
        if nq <= 64:
            v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, metric), k)
            assert np.array_equal(r0["i" + tag], i_ref) and np.array_equal(r0["d" + tag], v_ref)
        if dup:
            assert int(r0["s" + tag][0]) > 0                    # queries without a proof were handled over NCCL too
    qf, gf, qp, gp, qc, gc = synth.make_eval_set(300, 2500, 64, seed=9, num_ids=25, noise=1.5, missing_query_frac=0.05)
    cmc, mAP = ev.evaluate(ev.cosin_dist(torch.from_numpy(qf).cuda(), torch.from_numpy(gf).cuda()), qp, gp, qc, gc, 50)
    for r in (r0, r1):
        assert np.array_equal(r["cmc"], cmc) and float(r["mAP"]) == float(mAP)


@pytest.mark.parametrize("metric", [0, 1])
def test_second_chance_proves_clustered_queries(metric):
    """A gallery with clusters of ~600 near-duplicates (the re-ID case): with K' = 256 the completeness proof fails for the
    queries inside a cluster (more than K' - k rows within the coarse error of the k-th neighbour); the second chance
    (the same protocol over those queries with K' = 1024) reaches past the cluster and proves them -- no brute force -- and the
    result stays bit-identical to the oracle.  Exact duplicates beyond K' = 1024 still end in the brute-force leg."""
    _, ev = _mods()
    from oracle import eval_oracle as eo
    rng = np.random.default_rng(71)
    nq, ng, dim, k = 40, 12000, 128, 100
    q, g = _retrieval_inputs(nq, ng, dim, 72, dup_every=0)
    for c in range(3):                                          # three clusters of 600 rows, spread over the gallery
        centre = g[17 + c]
        rows = np.arange(100 + c, 100 + c + 600 * 19, 19)
        g[rows] = centre + 2e-4 * rng.standard_normal((600, dim)).astype(np.float32)
        q[c * 5:(c + 1) * 5] = centre + 1e-3 * rng.standard_normal((5, dim)).astype(np.float32)
    g /= np.linalg.norm(g, axis=1, keepdims=True)
    q /= np.linalg.norm(q, axis=1, keepdims=True)
    stats = torch.zeros(8, dtype=torch.int32, device="cuda")
    d, i = ev.sharded_topk(torch.from_numpy(q).cuda(), torch.from_numpy(g).cuda(), k, 0, metric=metric, stats=stats)
    v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q, g, metric), k)
    assert np.array_equal(i.cpu().numpy(), i_ref) and np.array_equal(d.cpu().numpy(), v_ref)
    flagged, _, _, _, second, brute = [int(v) for v in stats[:6].cpu()]
    assert flagged >= 15 and second >= 15 and brute == flagged - second and brute <= 2, (flagged, second, brute)
    # 1,540 exact duplicates: more than the longest list holds -> those queries do need the brute force
    q2, g2 = _retrieval_inputs(nq, 20011, dim, 11)
    q2[0] = g2[5] + 1e-3 * q2[0]
    q2[0] /= np.linalg.norm(q2[0])
    d2, i2 = ev.sharded_topk(torch.from_numpy(q2).cuda(), torch.from_numpy(g2).cuda(), k, 0, metric=metric, stats=stats)
    v_ref, i_ref = eo.topk_stable(eo.exact_distance_fixed(q2, g2, metric), k)
    assert np.array_equal(i2.cpu().numpy(), i_ref) and np.array_equal(d2.cpu().numpy(), v_ref)
    assert int(stats[5]) >= 1


def test_evaluate_property_checks_vs_reference_cmc_and_mean_ap(golden_dir):
    """SURVEY.md section 4 property tests on the GPU path: `evaluate` agrees with the reference's independent evaluators
    cmc(first_match_break=True) and mean_ap (sklearn) as computed by the REAL reference (tests/golden/eval_props.npz), and is
    invariant under a permutation of the gallery (tie-free matrix)."""
    _, ev = _mods()
    g = np.load(os.path.join(golden_dir, "eval_props.npz"))
    qf, gf, qp, gp, qc, gc = synth.make_eval_set(int(g["nq"]), int(g["ng_extra"]), int(g["dim"]), seed=int(g["seed"]), num_ids=25,
                                                 noise=float(g["noise"]), missing_query_frac=0.05)
    gf = gf + np.float32(1e-3) * np.random.default_rng(int(g["seed"])).standard_normal(gf.shape).astype(np.float32)
    topk = int(g["topk"])
    d = ev.cosin_dist(torch.from_numpy(qf).cuda(), torch.from_numpy(gf).cuda())
    cmc, mAP = ev.evaluate(d, qp, gp, qc, gc, topk)
    assert np.abs(cmc - g["cmc_first_match_break"]).max() < 1e-4 and abs(mAP - float(g["mAP_sklearn"])) < 1e-4
    perm = np.random.default_rng(0).permutation(gf.shape[0])
    cmc_p, mAP_p = ev.evaluate(d[:, torch.from_numpy(perm).cuda()].contiguous(), qp, gp[perm], qc, gc[perm], topk)
    assert np.array_equal(cmc_p, cmc) and abs(mAP_p - mAP) < 1e-12
