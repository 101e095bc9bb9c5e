"""CPU oracle — TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` /
``--impl reference`` legs may import this package.  ``pykg2vec_b200`` never does.

Two pieces:

* ``kge_oracle.c`` (this module binds it with ctypes): scalar C restatement of the
  reference's score functions and of the 1-vs-all rank computation in the
  canonical arithmetic of DESIGN.md §3.  Bit-exact contract with the CUDA path.
* ``ref_port.py``: torch restatement of the reference's ATen op chains (used as
  the fp64 gradient oracle and as the timed CPU baseline).
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SRC = os.path.join(_HERE, "kge_oracle.c")
_OUT_DIR = os.path.join(_HERE, "_build")
_SO = os.path.join(_OUT_DIR, "libkge_oracle.so")

MODEL_IDS = {
    "transe": 0, "transh": 1, "transd": 2, "transr": 3, "rotate": 4, "hole": 5,
    "distmult": 6, "complex": 7, "cp": 8, "simple": 9, "transm": 10, "rescal": 11, "analogy": 12,
    "simple_ignr": 13, "quate": 14, "octonione": 15, "kg2e": 16, "slm": 17, "sme": 18, "sme_bl": 19, "ntn": 20, "convkb": 21,
}
GROUP_TAIL, GROUP_HEAD = 0, 1
MAX_TABLES = 16


class KgeModel(ctypes.Structure):
    _fields_ = [
        ("model", ctypes.c_int32), ("dim", ctypes.c_int32), ("rel_dim", ctypes.c_int32),
        ("l1_flag", ctypes.c_int32), ("margin", ctypes.c_float), ("phase_scale", ctypes.c_float),
        ("num_ent", ctypes.c_int64), ("num_rel", ctypes.c_int64),
        ("tables", ctypes.c_void_p * MAX_TABLES),
    ]


def build(force=False):
    """gcc -O2 -ffp-contract=off (no FMA contraction: fmaf() is explicit) + OpenMP."""
    if not force and os.path.exists(_SO) and os.path.getmtime(_SO) >= os.path.getmtime(_SRC):
        return _SO
    os.makedirs(_OUT_DIR, exist_ok=True)
    cmd = ["gcc", "-O2", "-std=c11", "-fPIC", "-shared", "-ffp-contract=off", "-fno-fast-math",
           "-fopenmp", "-mfma", "-o", _SO + ".tmp", _SRC, "-lm"]
    subprocess.run(cmd, check=True)
    os.replace(_SO + ".tmp", _SO)
    return _SO


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = ctypes.CDLL(build())
        _lib.kgeo_sincosf.argtypes = [ctypes.c_float, ctypes.POINTER(ctypes.c_float),
                                      ctypes.POINTER(ctypes.c_float)]
    return _lib


def _f32(a):
    a = np.ascontiguousarray(a, dtype=np.float32)
    return a


def _i64(a):
    return np.ascontiguousarray(a, dtype=np.int64)


def _ptr(a):
    return a.ctypes.data_as(ctypes.c_void_p) if a is not None else None


class Model:
    """Host-side model description: name + list of fp32 numpy tables in C-ABI order."""

    def __init__(self, name, tables, dim, rel_dim=None, l1_flag=False, margin=0.0,
                 embedding_range=None):
        self.name = name.lower()
        self.tables = [_f32(t) for t in tables]
        self.dim = int(dim)
        self.rel_dim = int(rel_dim if rel_dim is not None else dim)
        self.l1_flag = bool(l1_flag)
        self.margin = float(margin)
        # RotatE: theta = r / (embedding_range / pi)  (pairwise.py:748,776-782)
        self.phase_scale = float(np.float32(np.pi / embedding_range)) if embedding_range else 0.0
        self.num_ent = self.tables[0].shape[0]
        rel_index = {"rotate": 2, "complex": 2, "simple": 2, "simple_ignr": 2, "quate": 4,
                     "octonione": 8, "kg2e": 2}.get(self.name, 1)
        self.num_rel = self.tables[rel_index].shape[0]

    def c_struct(self):
        m = KgeModel()
        m.model = MODEL_IDS[self.name]
        m.dim, m.rel_dim, m.l1_flag = self.dim, self.rel_dim, int(self.l1_flag)
        m.margin, m.phase_scale = self.margin, self.phase_scale
        m.num_ent, m.num_rel = self.num_ent, self.num_rel
        for k, t in enumerate(self.tables):
            m.tables[k] = t.ctypes.data
        return m


def score_fwd(model, h, r, t, grouping=GROUP_TAIL):
    h, r, t = _i64(h), _i64(r), _i64(t)
    out = np.empty(h.shape[0], dtype=np.float32)
    m = model.c_struct()
    rc = lib().kgeo_score_fwd(ctypes.byref(m), int(grouping), _ptr(h), _ptr(r), _ptr(t),
                              ctypes.c_int64(h.shape[0]), _ptr(out))
    assert rc == 0
    return out


def sweep_scores(model, grouping, h, r, t, row_lo=0, row_hi=None):
    row_hi = model.num_ent if row_hi is None else row_hi
    out = np.empty(row_hi - row_lo, dtype=np.float32)
    m = model.c_struct()
    rc = lib().kgeo_sweep_scores(ctypes.byref(m), int(grouping), ctypes.c_int64(h), ctypes.c_int64(r),
                                 ctypes.c_int64(t), ctypes.c_int64(row_lo), ctypes.c_int64(row_hi),
                                 _ptr(out))
    assert rc == 0
    return out


def rank_1vsall(model, qh, qr, qt, filt_t=None, filt_h=None, row_lo=0, row_hi=None):
    """filt_* = (ptr[Q+1], idx[nnz]) CSR of known positives per query, or None."""
    qh, qr, qt = _i64(qh), _i64(qr), _i64(qt)
    Q = qh.shape[0]
    row_hi = model.num_ent if row_hi is None else row_hi
    counts = np.zeros((Q, 4), dtype=np.int32)
    ft = (None, None) if filt_t is None else (_i64(filt_t[0]), _i64(filt_t[1]))
    fh = (None, None) if filt_h is None else (_i64(filt_h[0]), _i64(filt_h[1]))
    m = model.c_struct()
    rc = lib().kgeo_rank_1vsall(ctypes.byref(m), ctypes.c_int64(row_lo), ctypes.c_int64(row_hi),
                                _ptr(qh), _ptr(qr), _ptr(qt), ctypes.c_int64(Q),
                                _ptr(ft[0]), _ptr(ft[1]), _ptr(fh[0]), _ptr(fh[1]), _ptr(counts))
    assert rc == 0
    return counts


def loss_pairwise_hinge(pos, neg, margin):
    pos, neg = _f32(pos), _f32(neg)
    out = np.zeros(1, dtype=np.float32)
    terms = np.empty(pos.shape[0], dtype=np.float32)
    lib().kgeo_loss_pairwise_hinge(_ptr(pos), _ptr(neg), ctypes.c_int64(pos.shape[0]),
                                   ctypes.c_float(margin), _ptr(out), _ptr(terms))
    return float(out[0]), terms


def loss_pointwise_logistic(preds, target):
    preds, target = _f32(preds), _f32(target)
    out = np.zeros(1, dtype=np.float32)
    lib().kgeo_loss_pointwise_logistic(_ptr(preds), _ptr(target), ctypes.c_int64(preds.shape[0]),
                                       _ptr(out))
    return float(out[0])


def loss_selfadv(pos, neg, neg_rate, alpha):
    pos, neg = _f32(pos), _f32(neg)
    out = np.zeros(1, dtype=np.float32)
    lib().kgeo_loss_selfadv(_ptr(pos), _ptr(neg), ctypes.c_int64(pos.shape[0]),
                            ctypes.c_int32(neg_rate), ctypes.c_float(alpha), _ptr(out))
    return float(out[0])


def sample_negatives(train, ph, pr, pt, neg_rate, head_prob, num_ent, seed, step, layout=0):
    """train: [n,3] int64 positives.  Returns (h, r, t[, y]) numpy int64 arrays."""
    train = _i64(train)
    th, tr, tt = _i64(train[:, 0]), _i64(train[:, 1]), _i64(train[:, 2])
    ph, pr, pt = _i64(ph), _i64(pr), _i64(pt)
    B = ph.shape[0]
    n = B * neg_rate if layout == 0 else B * (1 + neg_rate)
    oh, orr, ot, oy = (np.zeros(n, dtype=np.int64) for _ in range(4))
    hp = _f32(head_prob) if head_prob is not None else None
    rc = lib().kgeo_sample_negatives(_ptr(th), _ptr(tr), _ptr(tt), ctypes.c_int64(th.shape[0]), _ptr(ph), _ptr(pr),
                                     _ptr(pt), ctypes.c_int64(B), ctypes.c_int32(neg_rate), _ptr(hp),
                                     ctypes.c_int64(num_ent), ctypes.c_uint64(seed), ctypes.c_uint64(step),
                                     ctypes.c_int32(layout), _ptr(oh), _ptr(orr), _ptr(ot), _ptr(oy))
    assert rc == 0
    return (oh, orr, ot) if layout == 0 else (oh, orr, ot, oy)


def normalize_rows(table):
    """in-place Rescal row normalisation of a contiguous fp32 [rows, width] numpy array"""
    assert table.dtype == np.float32 and table.flags["C_CONTIGUOUS"]
    lib().kgeo_normalize_rows(_ptr(table), ctypes.c_int64(table.shape[0]), ctypes.c_int64(table.shape[1]))
    return table


def tanhf(x):
    lib().kgeo_tanhf.restype = ctypes.c_float
    lib().kgeo_tanhf.argtypes = [ctypes.c_float]
    return lib().kgeo_tanhf(ctypes.c_float(x))


def logf(x):
    lib().kgeo_logf.restype = ctypes.c_float
    lib().kgeo_logf.argtypes = [ctypes.c_float]
    return lib().kgeo_logf(ctypes.c_float(x))


def expf(x):
    lib().kgeo_expf.restype = ctypes.c_float
    lib().kgeo_expf.argtypes = [ctypes.c_float]
    return lib().kgeo_expf(ctypes.c_float(x))


def sincosf(x):
    s, c = ctypes.c_float(), ctypes.c_float()
    lib().kgeo_sincosf(ctypes.c_float(x), ctypes.byref(s), ctypes.byref(c))
    return s.value, c.value


# ---- projection-model tail (x.E^T + b -> sigmoid, BCE, rank counts) -------------------------
def proj_tail_fwd(x, ent, bias=None):
    x, ent = _f32(x), _f32(ent)
    bias = _f32(bias).reshape(-1) if bias is not None else None
    B, k = x.shape
    N = ent.shape[0]
    out = np.empty((B, N), dtype=np.float32)
    rc = lib().kgeo_proj_tail_fwd(_ptr(x), _ptr(ent), _ptr(bias), ctypes.c_int64(B), ctypes.c_int64(N),
                                  ctypes.c_int32(k), _ptr(out))
    assert rc == 0
    return out


def proj_bce(preds, labels, label_scale=1.0, label_shift=0.0, grad_scale=1.0):
    """-> (loss, grad_preds) of one direction of Criterion.multi_class_bce."""
    preds, labels = _f32(preds), _f32(labels)
    B, N = preds.shape
    loss = np.zeros(1, dtype=np.float32)
    g = np.empty((B, N), dtype=np.float32)
    rc = lib().kgeo_proj_bce(_ptr(preds), _ptr(labels), ctypes.c_int64(B), ctypes.c_int64(N),
                             ctypes.c_float(label_scale), ctypes.c_float(label_shift),
                             ctypes.c_float(grad_scale), _ptr(loss), _ptr(g))
    assert rc == 0
    return float(loss[0]), g


def proj_tail_bwd(grad_preds, preds, x, ent):
    """-> (grad_x [B,k], grad_ent [N,k], grad_bias [N]) accumulated in double."""
    grad_preds, preds, x, ent = _f32(grad_preds), _f32(preds), _f32(x), _f32(ent)
    B, k = x.shape
    N = ent.shape[0]
    gx = np.zeros((B, k), dtype=np.float32)
    ge = np.zeros((N, k), dtype=np.float32)
    gb = np.zeros(N, dtype=np.float32)
    rc = lib().kgeo_proj_tail_bwd(_ptr(grad_preds), _ptr(preds), _ptr(x), _ptr(ent), ctypes.c_int64(B),
                                  ctypes.c_int64(N), ctypes.c_int32(k), _ptr(gx), _ptr(ge), _ptr(gb))
    assert rc == 0
    return gx, ge, gb


def proj_rank(x, ent, bias, tgt, filt=None, direction=0, counts=None):
    """counts [Q,4] (+= into columns 2*direction, 2*direction+1); filt = (ptr, idx) or None."""
    x, ent = _f32(x), _f32(ent)
    bias = _f32(bias).reshape(-1) if bias is not None else None
    tgt = _i64(tgt)
    Q, k = x.shape
    if counts is None:
        counts = np.zeros((Q, 4), dtype=np.int32)
    fp, fi = (None, None) if filt is None else (_i64(filt[0]), _i64(filt[1]))
    rc = lib().kgeo_proj_rank(_ptr(x), _ptr(ent), _ptr(bias), ctypes.c_int64(Q), ctypes.c_int64(ent.shape[0]),
                              ctypes.c_int32(k), _ptr(tgt), _ptr(fp), _ptr(fi), ctypes.c_int32(direction),
                              _ptr(counts))
    assert rc == 0
    return counts


class KgeConve(ctypes.Structure):
    _fields_ = [("hidden_size", ctypes.c_int32), ("hidden_size_1", ctypes.c_int32),
                ("bn0_eps", ctypes.c_float), ("bn1_eps", ctypes.c_float)] + \
               [(n, ctypes.c_void_p) for n in ("ent", "rel", "bn0_weight", "bn0_bias", "bn0_mean", "bn0_var",
                                               "conv_weight", "conv_bias", "bn1_weight", "bn1_bias", "bn1_mean",
                                               "bn1_var", "fc_weight", "fc_bias")]


CONVE_KEYS = {"ent": "ent_embeddings.weight", "rel": "rel_embeddings.weight", "bn0_weight": "bn0.weight",
              "bn0_bias": "bn0.bias", "bn0_mean": "bn0.running_mean", "bn0_var": "bn0.running_var",
              "conv_weight": "conv2d_1.weight", "conv_bias": "conv2d_1.bias", "bn1_weight": "bn1.weight",
              "bn1_bias": "bn1.bias", "bn1_mean": "bn1.running_mean", "bn1_var": "bn1.running_var",
              "fc_weight": "fc.weight", "fc_bias": "fc.bias"}


def conve_trunk_fwd(state, hidden_size, hidden_size_1, e, r, eps=1e-5):
    """state: dict state_dict-key -> numpy array (the reference ConvE's parameters and BN buffers);
    e, r: ids (r already offset by tot_relation for the head direction).  -> x [Q, hidden_size]."""
    keep = {f: _f32(state[k]) for f, k in CONVE_KEYS.items()}
    p = KgeConve()
    p.hidden_size, p.hidden_size_1, p.bn0_eps, p.bn1_eps = hidden_size, hidden_size_1, eps, eps
    for f, a in keep.items():
        setattr(p, f, a.ctypes.data)
    e, r = _i64(e), _i64(r)
    x = np.empty((e.shape[0], hidden_size), dtype=np.float32)
    rc = lib().kgeo_conve_trunk_fwd(ctypes.byref(p), _ptr(e), _ptr(r), ctypes.c_int64(e.shape[0]), _ptr(x))
    assert rc == 0
    return x
