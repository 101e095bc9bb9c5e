"""CPU check of the error budget behind the tensor-core sweep's two-level decision (DESIGN.md §4b,
pykg2vec_b200/csrc/kge_rank_tc.cuh `tc_query_finish`): a numpy emulation of level 1 — bf16x3 split of both
operands, products exact in fp32, fp32 accumulation k-step by k-step, the three norm columns of the distance
models — classified with the per-pair band  half(q,c) = a + b n_c + e n_c^2  restated here from the header.
Every pair the emulation calls certain must be decided the same way by the canonical fp32 scores of the oracle;
the pairs inside the band are the only ones level 2 has to re-evaluate, and they must be few.

This pins the MATH of the bound (split, accumulation, canonical-chain terms) without a GPU; the -m gpu tests
(test_gpu_baseline_shapes.py) check the same statement on the hardware's own accumulation."""
import numpy as np
import pytest

import gpu_util as gpu


def bf16_rn(x):
    """float32 -> nearest bfloat16 (ties to even), returned as float32"""
    u = np.ascontiguousarray(x, dtype=np.float32).view(np.uint32).astype(np.uint64)
    u = (u + 0x7FFF + ((u >> 16) & 1)) & 0xFFFF0000
    return u.astype(np.uint32).view(np.float32)


def split2(x):
    x = np.asarray(x, dtype=np.float32)
    hi = bf16_rn(x)
    return hi, bf16_rn((x - hi).astype(np.float32))


def split3(x):
    x = np.asarray(x, dtype=np.float32)
    n0 = bf16_rn(x)
    r1 = (x - n0).astype(np.float32)
    n1 = bf16_rn(r1)
    return n0, n1, bf16_rn((r1 - n1).astype(np.float32))


def tc_accumulate(A, B):
    """D[q, c] = sum_k a0 b0 + a0 b1 + a1 b0 with fp32 partial sums per 16-wide k-step (the products of bf16
    pairs are exact in fp32; one rounding per k-step and pass is far fewer than the budget assumes)."""
    a0, a1 = split2(A)
    b0, b1 = split2(B)
    K = A.shape[1]
    acc = np.zeros((A.shape[0], B.shape[0]), dtype=np.float32)
    for k in range(0, K, 16):
        for x, y in ((a0, b0), (a0, b1), (a1, b0)):
            part = x[:, k:k + 16].astype(np.float64) @ y[:, k:k + 16].astype(np.float64).T
            acc = (acc.astype(np.float64) + part).astype(np.float32)
    return acc


def band_coefficients(qvec, thr, kind, Kp, K, margin=0.0):
    """tc_query_finish (kge_rank_tc.cuh): centre, a, b, e per query — in float64, like the device code"""
    ss = (qvec.astype(np.float64) ** 2).sum(-1)
    nq = np.sqrt(ss) * (1.0 + 1e-7)
    nmma = 3 * ((Kp + 15) // 16)
    acc = nmma * 2.0 ** -21
    gamma = (K / 8.0 + 8.0) * 2.0 ** -24
    if kind == 0:
        centre = -thr.astype(np.float64)
        a = np.zeros_like(nq)
        b = nq * (2.0 ** -16 + acc + gamma)
        e = np.zeros_like(nq)
    else:
        g2 = gamma + 2.0 ** -22
        ks = 0.5 * g2
        # canonical: sum < T(thr), T = the smallest fp32 x with sqrt_rn(x) >= thr (rule 7 of DESIGN.md §3)
        t = thr.astype(np.float32)
        x = (t * t).astype(np.float32)
        for _ in range(4):
            up = np.sqrt(x).astype(np.float32) < t
            x = np.where(up, np.nextafter(x, np.float32(np.inf)), x).astype(np.float32)
        for _ in range(4):
            down = np.sqrt(np.nextafter(x, np.float32(0))).astype(np.float32) >= t
            x = np.where(down & (x > 0), np.nextafter(x, np.float32(0)), x).astype(np.float32)
        centre = 0.5 * (ss - x.astype(np.float64))
        a = 2.0 ** -50 * ss + ks * nq * nq
        b = nq * (2.0 ** -16 + acc) + 2.0 * ks * nq
        e = np.full_like(nq, 0.5 * acc + 2.0 ** -22 + ks)
    infl = 1.0 + 2.0 ** -18
    a = a * infl + 2.0 ** -23 * np.abs(centre) + 1e-30
    return centre.astype(np.float32).astype(np.float64), a, b * infl, e * infl


@pytest.mark.parametrize("name,d,heavy", [("distmult", 40, False), ("distmult", 64, True), ("transe", 48, False),
                                          ("complex", 36, False), ("complex", 40, True)])
def test_certain_pairs_agree_with_canonical_scores(name, d, heavy):
    import oracle
    N, R, Q = 900, 5, 24
    om, tabs = gpu.synthetic_case(name, N, R, d, seed=3 * d)
    if heavy:   # a quarter of the entity rows 8x heavier: the band of a pair must scale with ITS candidate
        rows = np.random.RandomState(1).choice(N, N // 4, replace=False)
        for k in range(2 if name == "complex" else 1):
            tabs[k][rows] *= 8.0
        om = oracle.Model(name, tabs, d)
    rng = np.random.RandomState(d)
    qh, qr = rng.randint(N, size=Q), rng.randint(R, size=Q)
    qt = rng.randint(N, size=Q)
    ent, rel = tabs[0], tabs[1]
    if name == "complex":   # tail direction: score = -(e_re . a + e_im . b), a = hr rr - hi ri, b = hi rr + hr ri
        er, ei, rr, ri = tabs
        kind, K = 0, 2 * d
        f = np.float32
        qa = (er[qh] * rr[qr]).astype(f) - (ei[qh] * ri[qr]).astype(f)
        qb = (ei[qh] * rr[qr]).astype(f) + (er[qh] * ri[qr]).astype(f)
        qvec = np.concatenate([qa.astype(f), qb.astype(f)], axis=1)
        cand = np.concatenate([er, ei], axis=1)
        aug_q = aug_c = None
    elif name == "distmult":
        kind, K = 0, d
        qvec = (ent[qh] * rel[qr]).astype(np.float32)     # tail direction: score = -(h o r) . t
        cand = ent
        aug_q = aug_c = None
    else:
        kind, K = 1, d
        nrm = lambda x: (x / np.maximum(np.sqrt((x.astype(np.float32) ** 2).sum(-1, keepdims=True, dtype=np.float32)), 1e-12)).astype(np.float32)
        cand = nrm(ent)
        qvec = (nrm(ent[qh]) + nrm(rel[qr])).astype(np.float32)   # |q - c|, q = h^ + r^
        half_norm = (0.5 * (cand.astype(np.float64) ** 2).sum(-1)).astype(np.float32)
        aug_c = np.stack(split3(half_norm), axis=1)               # rides against -1 in the query
        aug_q = -np.ones((Q, 3), dtype=np.float32)
    Kp = ((K + (3 if kind else 0) + 15) // 16) * 16
    A = np.zeros((Q, Kp), dtype=np.float32)
    B = np.zeros((N, Kp), dtype=np.float32)
    A[:, :K], B[:, :K] = qvec, cand
    if kind:
        A[:, K:K + 3], B[:, K:K + 3] = aug_q, aug_c
    D = tc_accumulate(A, B).astype(np.float64)
    thr = oracle.score_fwd(om, qh, qr, qt, 0).astype(np.float32)
    centre, a, b, e = band_coefficients(qvec, thr, kind, Kp, K)
    n = np.sqrt((cand.astype(np.float64) ** 2).sum(-1)) * (1.0 + 1e-7)
    half = a[:, None] + b[:, None] * n[None, :] + e[:, None] * n[None, :] ** 2
    u = D - centre[:, None]
    yes, no = u > half, u < -half
    for i in range(Q):
        s = oracle.score_fwd(om, np.full(N, qh[i]), np.full(N, qr[i]), np.arange(N), 0)
        better = s < thr[i]
        assert better[yes[i]].all(), (name, i)
        assert not better[no[i]].any(), (name, i)
    band = (~yes & ~no).sum() / Q
    assert band < 12, band   # a handful of pairs per query (ties with the target itself included)
