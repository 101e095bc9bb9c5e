"""CPU check (no GPU) of project_rows_kernel (pykg2vec_b200/csrc/kge_project.cuh) run under the host
emulation of tests/emu/: for a fixed relation r, TransE over [P_r, rel] must give TransH's / TransD's
scores BIT FOR BIT (the oracle evaluates both sides) — the identity the relation-grouped 1-vs-all
evaluation of these models rests on."""
import ctypes

import numpy as np
import pytest

import oracle

import emu_build


@pytest.fixture(scope="module")
def emu():
    return ctypes.CDLL(emu_build.models_lib())


def _tables(name, N, R, d, seed):
    rng = np.random.RandomState(seed)
    def t(rows):
        return (rng.standard_normal((rows, d)) * 0.5).astype(np.float32)
    if name == "transh":
        return [t(N), t(R), t(R)]            # ent, rel, w
    return [t(N), t(R), t(N), t(R)]          # ent, rel, ent_map, rel_map


@pytest.mark.parametrize("name,N,R,d,l1", [("transh", 70, 4, 48, False), ("transh", 33, 3, 50, True),
                                           ("transh", 40, 2, 7, False), ("transd", 70, 4, 48, True),
                                           ("transd", 45, 3, 200, False), ("transd", 21, 2, 10, False)])
def test_transe_on_projected_table_equals_projected_model(emu, name, N, R, d, l1):
    tabs = _tables(name, N, R, d, seed=N * d)
    om = oracle.Model(name, tabs, d, l1_flag=l1)
    m = om.c_struct()
    rng = np.random.RandomState(1)
    for r in range(R):
        P = np.full((N, d), np.nan, dtype=np.float32)
        rc = emu.emu_project_entities(ctypes.byref(m), ctypes.c_int64(r), P.ctypes.data_as(ctypes.c_void_p))
        assert rc == 0 and np.isfinite(P).all()
        te = oracle.Model("transe", [P, tabs[1]], d, l1_flag=l1)
        n = 64
        h, t = rng.randint(N, size=n), rng.randint(N, size=n)
        rr = np.full(n, r, dtype=np.int64)
        for grouping in (oracle.GROUP_TAIL, oracle.GROUP_HEAD):
            want = oracle.score_fwd(om, h, rr, t, grouping)
            got = oracle.score_fwd(te, h, rr, t, grouping)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (name, r, grouping)
        # and therefore identical rank counts for queries of this relation
        q = 5
        filt = (np.arange(q + 1, dtype=np.int64) * 3, rng.randint(N, size=3 * q).astype(np.int64))
        assert np.array_equal(oracle.rank_1vsall(te, h[:q], rr[:q], t[:q], filt, filt),
                              oracle.rank_1vsall(om, h[:q], rr[:q], t[:q], filt, filt))


@pytest.mark.parametrize("N,R,d,dr,l1", [(40, 3, 24, 16, False), (25, 2, 50, 50, True), (30, 2, 10, 7, False)])
def test_transe_on_projected_table_equals_transr(emu, N, R, d, dr, l1):
    """TransR: P_r = normalize(ent) . M_r and the once-normalised relation rows; TransE (width d_r) over
    them applies the reference's second normalisation — bit-identical scores in both groupings"""
    rng = np.random.RandomState(N * d + dr)
    ent = (rng.standard_normal((N, d)) * 0.5).astype(np.float32)
    rel = (rng.standard_normal((R, dr)) * 0.5).astype(np.float32)
    mats = (rng.standard_normal((R, d * dr)) * 0.3).astype(np.float32)
    om = oracle.Model("transr", [ent, rel, mats], d, rel_dim=dr, l1_flag=l1)
    m = om.c_struct()
    vec = 4 if (d % 4 == 0 and dr % 4 == 0) else 1
    rhat = np.full((R, dr), np.nan, dtype=np.float32)
    emu.emu_normalize_rows(rel.ctypes.data_as(ctypes.c_void_p), ctypes.c_int64(R), ctypes.c_int(dr), ctypes.c_int(vec),
                           rhat.ctypes.data_as(ctypes.c_void_p))
    for r in range(R):
        P = np.full((N, dr), np.nan, dtype=np.float32)
        assert emu.emu_project_entities(ctypes.byref(m), ctypes.c_int64(r), P.ctypes.data_as(ctypes.c_void_p)) == 0
        assert np.isfinite(P).all()
        te = oracle.Model("transe", [P, rhat], dr, l1_flag=l1)
        n = 48
        h, t = rng.randint(N, size=n), rng.randint(N, size=n)
        rr = np.full(n, r, dtype=np.int64)
        for grouping in (oracle.GROUP_TAIL, oracle.GROUP_HEAD):
            want = oracle.score_fwd(om, h, rr, t, grouping)
            got = oracle.score_fwd(te, h, rr, t, grouping)
            assert np.array_equal(got.view(np.uint32), want.view(np.uint32)), (r, grouping)


# ---- the Evaluator's relation-grouped path end to end on the CPU ------------------------------------
# host logic (sort by relation, CSR reordering, group slicing, scatter back) is the product's own code;
# the device calls are replaced by the emulated projection kernels and the oracle's TransE rank counts.
@pytest.mark.parametrize("name", ["transh", "transd", "transr"])
def test_grouped_evaluator_host_logic(emu, name, monkeypatch):
    import types
    import torch
    from pykg2vec_b200 import _lib, evaluator
    N, R, d, dr, Q = 90, 5, 24, 16, 140
    rng = np.random.RandomState(7)
    if name == "transr":
        tabs = [(rng.standard_normal((N, d)) * 0.5).astype(np.float32), (rng.standard_normal((R, dr)) * 0.5).astype(np.float32),
                (rng.standard_normal((R, d * dr)) * 0.3).astype(np.float32)]
        om = oracle.Model("transr", tabs, d, rel_dim=dr)
    else:
        tabs = _tables(name, N, R, d, seed=3)
        om = oracle.Model(name, tabs, d)

    class FakeDesc:   # what the grouped path reads of _lib.ModelDesc
        def __init__(self, mname, tables, dim, rel_dim=None, l1_flag=False, **kw):
            self.name, self.tables, self.dim = mname, tables, dim
            self.rel_dim = rel_dim if rel_dim is not None else dim
            self.l1_flag, self.num_ent = l1_flag, tables[0].shape[0]

    desc = FakeDesc(name, [torch.from_numpy(t) for t in tabs], d, rel_dim=om.rel_dim)
    cm = om.c_struct()

    def project_entities(dsc, r, out):
        assert emu.emu_project_entities(ctypes.byref(cm), ctypes.c_int64(r), ctypes.c_void_p(out.data_ptr())) == 0
        return out

    def normalize_rows_to(table):
        out = torch.empty_like(table)
        emu.emu_normalize_rows(ctypes.c_void_p(table.data_ptr()), ctypes.c_int64(table.shape[0]),
                               ctypes.c_int(table.shape[1]), ctypes.c_int(4 if table.shape[1] % 4 == 0 else 1),
                               ctypes.c_void_p(out.data_ptr()))
        return out

    def rank_1vsall(te, qh, qr, qt, filt_t=None, filt_h=None, counts=None, workspace=None, **kw):
        tm = oracle.Model("transe", [t.numpy() for t in te.tables], te.dim, l1_flag=te.l1_flag)
        f = [None if x is None else (x[0].numpy(), x[1].numpy()) for x in (filt_t, filt_h)]
        counts += torch.from_numpy(oracle.rank_1vsall(tm, qh.numpy(), qr.numpy(), qt.numpy(), f[0], f[1]))
        return counts

    for attr, fn in (("ModelDesc", FakeDesc), ("project_entities", project_entities),
                     ("normalize_rows_to", normalize_rows_to), ("rank_1vsall", rank_1vsall),
                     ("rank_workspace_bytes", lambda te, q: 16)):
        monkeypatch.setattr(_lib, attr, fn)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    qh, qr, qt = rng.randint(N, size=Q), rng.randint(R, size=Q), rng.randint(N, size=Q)
    tp, hp = np.zeros(Q + 1, dtype=np.int64), np.zeros(Q + 1, dtype=np.int64)
    ti, hi = [], []
    for i in range(Q):
        a = np.unique(np.r_[rng.randint(N, size=rng.randint(0, 6)), qt[i]])
        b = np.unique(rng.randint(N, size=rng.randint(0, 6)))
        ti.append(a); hi.append(b)
        tp[i + 1], hp[i + 1] = tp[i] + len(a), hp[i] + len(b)
    ft = (tp, np.concatenate(ti).astype(np.int64))
    fh = (hp, np.concatenate(hi).astype(np.int64) if hp[-1] else np.zeros(0, dtype=np.int64))
    ev = object.__new__(evaluator.Evaluator)
    ev.model = types.SimpleNamespace(model_name=name, kge_desc=lambda: desc)
    ev.config = types.SimpleNamespace(device="cpu", tot_entity=N, relation_grouped_eval=True)
    ev._filter_cache, ev._workspace = {}, None
    want = oracle.rank_1vsall(om, qh, qr, qt, ft, fh)
    assert np.array_equal(ev.rank_triples(qh, qr, qt, ft, fh), want)
    assert np.array_equal(ev.rank_triples(qh, qr, qt)[:, [0, 2]], oracle.rank_1vsall(om, qh, qr, qt)[:, [0, 2]])
    # the automatic choice: >= 32 queries per distinct relation for TransH / TransD, TransR only on request
    ev.config.relation_grouped_eval = None
    assert ev._use_relation_groups(qr) == (name != "transr" and Q >= 32 * len(np.unique(qr)))
    assert not ev._use_relation_groups(qr[:10]) and not ev._use_relation_groups(qr[:0])
