"""Helpers shared by the tests: load a golden case (tests/golden/*.npz, produced by the
reference itself — see make_golden.py) as an oracle.Model."""
import glob
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
NON_CASES = {"losses", "settle"}
PROJ_PREFIXES = ("conve_", "tucker_")


def case_names():
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if n not in NON_CASES and not n.startswith(PROJ_PREFIXES)]


def proj_case_names():
    """ConvE cases (tests/golden/make_golden_proj.py): full state_dict + tail operands, not table lists."""
    names = sorted(os.path.splitext(os.path.basename(p))[0] for p in glob.glob(os.path.join(GOLDEN_DIR, "*.npz")))
    return [n for n in names if n.startswith("conve_")]


def proj_state(g):
    """state_dict (numpy) of the reference ConvE stored in a conve_* case"""
    return {k[3:]: g[k] for k in g if k.startswith("sd_") and not k.startswith("sd_after_")}


def load(name):
    return dict(np.load(os.path.join(GOLDEN_DIR, name + ".npz"), allow_pickle=False))


def tables_of(g):
    out, k = [], 0
    while "table%d" % k in g:
        out.append(g["table%d" % k])
        k += 1
    return out


def raw_tables_of(g):
    """ConvKB only: the reference's own parameters [ent, rel, conv_w0, conv_b0, ..., fc_w, fc_b]."""
    out, k = [], 0
    while "raw%d" % k in g:
        out.append(g["raw%d" % k])
        k += 1
    return out


def model_kwargs(g):
    """-> dict(name, dim, rel_dim, l1_flag, margin, embedding_range)"""
    name = str(g["model"])
    kw = {k[3:]: g[k] for k in g if k.startswith("kw_")}
    dim = int(kw.get("hidden_size", kw.get("ent_hidden_size", 0)))
    rel_dim = int(kw.get("rel_hidden_size", dim))
    margin = float(kw.get("margin", 0.0))
    if name in ("rescal", "sme", "sme_bl"):
        rel_dim = dim
    return dict(name=name, dim=dim, rel_dim=rel_dim, l1_flag=bool(kw.get("l1_flag", False)),
                margin=margin,
                embedding_range=((margin + 2.0) / dim) if name == "rotate" else None)


def oracle_model(g):
    import oracle
    kw = model_kwargs(g)
    return oracle.Model(kw["name"], tables_of(g), kw["dim"], rel_dim=kw["rel_dim"],
                        l1_flag=kw["l1_flag"], margin=kw["margin"],
                        embedding_range=kw["embedding_range"])


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return np.abs(a - b) / np.maximum(np.abs(b), 1e-30)
