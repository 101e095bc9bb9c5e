"""CPU check (no GPU) of project_rows_kernel (pykg2vec_b200/csrc/kge_project.cuh) run under the host
emulation of tests/emu/: for a fixed relation r, TransE over [P_r, rel] must give TransH's / TransD's
scores BIT FOR BIT (the oracle evaluates both sides) — the identity the relation-grouped 1-vs-all
evaluation of these models rests on."""
import ctypes
import os
import subprocess

import numpy as np
import pytest

import oracle

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
SRC = os.path.join(HERE, "emu", "emu_project.cpp")
OUT = os.path.join(HERE, "emu", "_build", "libemu_project.so")
DEPS = [SRC, os.path.join(HERE, "emu", "cuda_runtime.h")] + \
       [os.path.join(ROOT, "pykg2vec_b200", "csrc", f) for f in ("kge_project.cuh", "kge_models.cuh", "kge_common.cuh")]


@pytest.fixture(scope="module")
def emu():
    if not os.path.exists(OUT) or any(os.path.getmtime(d) > os.path.getmtime(OUT) for d in DEPS):
        os.makedirs(os.path.dirname(OUT), exist_ok=True)
        cmd = ["g++", "-O1", "-std=c++17", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math", "-mfma",
               "-pthread", "-w", "-I", os.path.join(HERE, "emu"), "-I", os.path.join(ROOT, "pykg2vec_b200", "csrc"),
               "-o", OUT + ".tmp", SRC]
        subprocess.run(cmd, check=True)
        os.replace(OUT + ".tmp", OUT)
    return ctypes.CDLL(OUT)


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
