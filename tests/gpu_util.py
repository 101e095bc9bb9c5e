"""Helpers for the -m gpu tests: move a golden / synthetic case onto cuda:0 as a
pykg2vec_b200 ModelDesc (C-ABI table order) next to the matching oracle.Model."""
import numpy as np
import torch

import golden_util as gu


def desc_from_golden(g, device="cuda"):
    from pykg2vec_b200 import _lib
    kw = gu.model_kwargs(g)
    tabs = [torch.from_numpy(np.ascontiguousarray(t)).to(device) for t in gu.tables_of(g)]
    phase = float(np.float32(np.pi / kw["embedding_range"])) if kw["embedding_range"] else 0.0
    return _lib.ModelDesc(kw["name"], tabs, kw["dim"], rel_dim=kw["rel_dim"], l1_flag=kw["l1_flag"],
                          margin=kw["margin"], phase_scale=phase)


NUM_TABLE_SPECS = {
    # name: list of (kind, width_fn) in C-ABI order; kind 'e' entity rows, 'r' relation rows
    "transe": ["e", "r"], "transh": ["e", "r", "r"], "transd": ["e", "r", "e", "r"],
    "transr": ["e", "r", "M"], "rotate": ["e", "e", "r"], "distmult": ["e", "r"],
    "cp": ["e", "r", "e"], "complex": ["e", "e", "r", "r"], "transm": ["e", "r", "theta"],
    "analogy": ["e", "r", "e2", "e2", "r2", "r2"], "kg2e": ["e", "e+", "r", "r+"], "slm": ["e", "r", "dk", "dk"], "ntn": ["e", "r", "dk", "dk", "1k", "kdd"],
    "sme": ["e", "r", "dd", "dd", "d1", "dd", "dd", "d1"], "sme_bl": ["e", "r", "dd", "dd", "d1", "dd", "dd", "d1"], "quate": ["e"] * 4 + ["r"] * 4, "octonione": ["e"] * 8 + ["r"] * 8, "hole": ["e", "r"], "rescal": ["e", "MM"], "simple": ["e", "e", "r", "r"], "simple_ignr": ["e", "e", "r", "r"],
    "convkb": ["e", "r", "3d", "1"],
}


def synthetic_case(name, N, R, d, seed, dr=None, l1=False, margin=0.0, scale=0.5):
    """Seeded N(0, scale) tables -> (oracle.Model, kwargs for ModelDesc, numpy tables)."""
    import oracle
    rng = np.random.RandomState(seed)
    dr = d if dr is None else dr
    tabs = []
    for kind in NUM_TABLE_SPECS[name]:
        if kind == "e":
            tabs.append((rng.standard_normal((N, d)) * scale).astype(np.float32))
        elif kind == "r":
            tabs.append((rng.standard_normal((R, dr)) * scale).astype(np.float32))
        elif kind == "M":
            tabs.append((rng.standard_normal((R, d * dr)) * scale).astype(np.float32))
        elif kind in ("dk", "1k", "kdd", "dd", "d1", "3d", "1"):  # global dense parameters
            shape = {"dk": (d, dr), "1k": (1, dr), "kdd": (dr, d * d), "dd": (d, d), "d1": (d, 1), "3d": (3, d), "1": (1,)}[kind]
            tabs.append((rng.standard_normal(shape) * scale).astype(np.float32))
        elif kind in ("e+", "r+"):  # strictly positive (variances)
            rows = N if kind == "e+" else R
            tabs.append((0.05 + rng.rand(rows, d)).astype(np.float32))
        elif kind == "e2":
            tabs.append((rng.standard_normal((N, d // 2)) * scale).astype(np.float32))
        elif kind == "r2":
            tabs.append((rng.standard_normal((R, d // 2)) * scale).astype(np.float32))
        elif kind == "MM":
            tabs.append((rng.standard_normal((R, d * d)) * scale).astype(np.float32))
        elif kind == "theta":
            tabs.append((0.2 + rng.rand(R)).astype(np.float32))
    emb_range = (margin + 2.0) / d if name == "rotate" else None
    om = oracle.Model(name, tabs, d, rel_dim=dr, l1_flag=l1, margin=margin, embedding_range=emb_range)
    return om, tabs


def desc_from_oracle_model(om, device="cuda"):
    from pykg2vec_b200 import _lib
    tabs = [torch.from_numpy(t).to(device) for t in om.tables]
    return _lib.ModelDesc(om.name, tabs, om.dim, rel_dim=om.rel_dim, l1_flag=om.l1_flag, margin=om.margin,
                          phase_scale=om.phase_scale)


def bits(a):
    return np.ascontiguousarray(a, dtype=np.float32).view(np.uint32)


def random_filters_csr(rng, N, qh, qr, qt, per_query=12):
    """CSR filters containing the target plus random known entities (some duplicates allowed? no: unique)."""
    Q = len(qh)
    tp, ti, hp, hi = [0], [], [0], []
    for i in range(Q):
        ts = set(rng.randint(N, size=per_query).tolist()) | {int(qt[i])}
        hs = set(rng.randint(N, size=per_query).tolist()) | {int(qh[i])}
        ti.extend(sorted(ts)); tp.append(len(ti))
        hi.extend(sorted(hs)); hp.append(len(hi))
    return (np.asarray(tp, np.int64), np.asarray(ti, np.int64)), (np.asarray(hp, np.int64), np.asarray(hi, np.int64))
