"""ctypes binding of libkge_b200.so (the C-ABI declared in include/kge_b200.h).

There is no fallback: if the library is missing or a call fails this module raises.
torch is used only for device memory, streams and tensor metadata.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "_build", "libkge_b200.so")
MAX_TABLES = 16
ABI_VERSION = 5
ENOTSUP = -2   # include/kge_b200.h KGE_ENOTSUP

MODEL_IDS = {
    "transe": 0, "transh": 1, "transd": 2, "transr": 3, "rotate": 4, "hole": 5,
    "distmult": 6, "complex": 7, "cp": 8, "simple": 9, "transm": 10, "rescal": 11, "analogy": 12,
    "simple_ignr": 13, "quate": 14, "octonione": 15, "kg2e": 16, "slm": 17, "sme": 18, "sme_bl": 19, "ntn": 20, "convkb": 21,
}
GROUP_TAIL, GROUP_HEAD = 0, 1
RANK_FORCE_GATHER, RANK_TAIL_ONLY, RANK_HEAD_ONLY = 1, 2, 4
RANK_SINGLE_STREAM, RANK_NO_TC, RANK_PROFILE = 8, 16, 32

# every symbol include/kge_b200.h declares (tests check they are all exported)
EXPORTS = [
    "kge_abi_version", "kge_version", "kge_last_error", "kge_launch_count",
    "kge_score_fwd", "kge_score_bwd", "kge_normalize_rows",
    "kge_loss_pairwise_hinge", "kge_loss_pointwise_logistic", "kge_loss_selfadv", "kge_reg_fwd_bwd",
    "kge_train_pairwise_hinge_sgd", "kge_train_pointwise_logistic", "kge_train_pairwise_selfadv", "kge_optim_apply_rows", "kge_optim_apply_dense",
    "kge_rank_workspace_bytes", "kge_rank_1vsall", "kge_rank_tc_probe", "kge_rank_last_sweep_ms", "kge_rank_last_sweep_directions", "kge_debug_set_tc_trace",
    "kge_tripleset_capacity", "kge_tripleset_build", "kge_sample_negatives",
    "kge_proj_tail_fwd", "kge_proj_tail_bwd", "kge_proj_bce", "kge_proj_rank_workspace_bytes", "kge_proj_rank", "kge_proj_labels",
    "kge_conve_trunk_workspace_bytes", "kge_conve_trunk_fwd", "kge_project_entities", "kge_normalize_rows_to",
]


class KgeModel(ctypes.Structure):
    _fields_ = [
        ("model", ctypes.c_int32), ("dim", ctypes.c_int32), ("rel_dim", ctypes.c_int32),
        ("l1_flag", ctypes.c_int32), ("margin", ctypes.c_float), ("phase_scale", ctypes.c_float),
        ("num_ent", ctypes.c_int64), ("num_rel", ctypes.c_int64),
        ("tables", ctypes.c_void_p * MAX_TABLES),
    ]


class KgeError(RuntimeError):
    pass


class KgeNotSupported(KgeError):
    """KGE_ENOTSUP: the entry point has no kernel for this model / shape (callers with an alternative path
    catch exactly this; every other error propagates)."""


_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built: the product
    path never falls back to a CPU or PyTorch implementation."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise KgeError("libkge_b200.so not found at %s — run `python -m pykg2vec_b200.build` "
                       "(or __graft_entry__.build())" % LIB_PATH)
    L = ctypes.CDLL(LIB_PATH)
    L.kge_version.restype = ctypes.c_char_p
    L.kge_last_error.restype = ctypes.c_char_p
    L.kge_launch_count.restype = ctypes.c_int64
    L.kge_rank_workspace_bytes.restype = ctypes.c_int64
    L.kge_tripleset_capacity.restype = ctypes.c_int64
    L.kge_proj_rank_workspace_bytes.restype = ctypes.c_int64
    L.kge_conve_trunk_workspace_bytes.restype = ctypes.c_int64
    if L.kge_abi_version() != ABI_VERSION:
        raise KgeError("libkge_b200.so ABI %d != binding ABI %d" % (L.kge_abi_version(), ABI_VERSION))
    _lib = L
    return L


def check(rc, what):
    if rc != 0:
        raise (KgeNotSupported if rc == ENOTSUP else KgeError)("%s failed (%d): %s" % (what, rc, lib().kge_last_error().decode()))


def _ptr(t):
    if t is None:
        return ctypes.c_void_p(0)
    return ctypes.c_void_p(t.data_ptr())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


def _dev_i64(t, name):
    if t.dtype != torch.int64 or not t.is_cuda or not t.is_contiguous():
        raise KgeError("%s must be a contiguous CUDA int64 tensor" % name)
    return t


def _dev_f32(t, name):
    if t.dtype != torch.float32 or not t.is_cuda or not t.is_contiguous():
        raise KgeError("%s must be a contiguous CUDA float32 tensor" % name)
    return t


class ModelDesc:
    """Host description of a model: C-ABI name + device tables in C-ABI order."""

    def __init__(self, name, tables, dim, rel_dim=None, l1_flag=False, margin=0.0, phase_scale=0.0,
                 num_ent=None, num_rel=None):
        self.name = name.lower()
        if self.name not in MODEL_IDS:
            raise KgeError("unknown model %r" % name)
        self.tables = [_dev_f32(t, "table") for t in tables]
        self.dim = int(dim)
        self.rel_dim = int(rel_dim if rel_dim is not None else dim)
        self.l1_flag = bool(l1_flag)
        self.margin = float(margin)
        self.phase_scale = float(phase_scale)
        rel_index = {"rotate": 2, "complex": 2, "simple": 2, "simple_ignr": 2, "quate": 4,
                     "octonione": 8, "kg2e": 2}.get(self.name, 1)
        self.num_ent = int(num_ent if num_ent is not None else self.tables[0].shape[0])
        self.num_rel = int(num_rel if num_rel is not None else self.tables[rel_index].shape[0])

    def c_struct(self, tables=None):
        m = KgeModel()
        m.model = MODEL_IDS[self.name]
        m.dim, m.rel_dim, m.l1_flag = self.dim, self.rel_dim, int(self.l1_flag)
        m.margin, m.phase_scale = self.margin, self.phase_scale
        m.num_ent, m.num_rel = self.num_ent, self.num_rel
        for k, t in enumerate(tables if tables is not None else self.tables):
            m.tables[k] = t.data_ptr()
        return m


def _table_ptr_array(tensors):
    arr = (ctypes.c_void_p * MAX_TABLES)()
    for k, t in enumerate(tensors):
        arr[k] = t.data_ptr() if t is not None else None
    return arr


def score_fwd(desc, h, r, t, grouping=GROUP_TAIL, out=None):
    h, r, t = _dev_i64(h, "h"), _dev_i64(r, "r"), _dev_i64(t, "t")
    n = h.numel()
    if r.numel() != n or t.numel() != n:
        raise KgeError("h, r, t must have equal length")
    if out is None:
        out = torch.empty(n, dtype=torch.float32, device=h.device)
    m = desc.c_struct()
    check(lib().kge_score_fwd(ctypes.byref(m), int(grouping), _ptr(h), _ptr(r), _ptr(t),
                              ctypes.c_int64(n), _ptr(out), _stream()), "kge_score_fwd")
    return out


def score_bwd(desc, h, r, t, grad_scores, grad_tables):
    m = desc.c_struct()
    arr = _table_ptr_array(grad_tables)
    check(lib().kge_score_bwd(ctypes.byref(m), _ptr(h), _ptr(r), _ptr(t), ctypes.c_int64(h.numel()),
                              _ptr(_dev_f32(grad_scores, "grad_scores")), arr, _stream()), "kge_score_bwd")


def normalize_rows(table):
    """Rescal.get_normalized_data (pairwise.py:862-865): rows / ||row||_2, IN PLACE."""
    t = _dev_f32(table, "table")
    check(lib().kge_normalize_rows(_ptr(t), ctypes.c_int64(t.shape[0]), ctypes.c_int64(t.shape[1]), _stream()),
          "kge_normalize_rows")
    return table


def loss_pairwise_hinge(pos, neg, margin, want_grad=True):
    pos, neg = _dev_f32(pos, "pos"), _dev_f32(neg, "neg")
    loss = torch.empty(1, dtype=torch.float32, device=pos.device)
    gp = torch.empty_like(pos) if want_grad else None
    gn = torch.empty_like(neg) if want_grad else None
    check(lib().kge_loss_pairwise_hinge(_ptr(pos), _ptr(neg), ctypes.c_int64(pos.numel()),
                                        ctypes.c_float(margin), _ptr(loss), _ptr(gp), _ptr(gn), _stream()),
          "kge_loss_pairwise_hinge")
    return loss, gp, gn


def loss_pointwise_logistic(preds, target, want_grad=True):
    preds, target = _dev_f32(preds, "preds"), _dev_f32(target, "target")
    loss = torch.empty(1, dtype=torch.float32, device=preds.device)
    g = torch.empty_like(preds) if want_grad else None
    check(lib().kge_loss_pointwise_logistic(_ptr(preds), _ptr(target), ctypes.c_int64(preds.numel()),
                                            _ptr(loss), _ptr(g), _stream()), "kge_loss_pointwise_logistic")
    return loss, g


def loss_selfadv(pos, neg, neg_rate, alpha, want_grad=True):
    pos, neg = _dev_f32(pos, "pos"), _dev_f32(neg, "neg")
    if neg.numel() != pos.numel() * neg_rate:
        raise KgeError("neg must hold neg_rate scores per positive")
    loss = torch.empty(1, dtype=torch.float32, device=pos.device)
    gp = torch.empty_like(pos) if want_grad else None
    gn = torch.empty_like(neg) if want_grad else None
    check(lib().kge_loss_selfadv(_ptr(pos), _ptr(neg), ctypes.c_int64(pos.numel()), ctypes.c_int32(neg_rate),
                                 ctypes.c_float(alpha), _ptr(loss), _ptr(gp), _ptr(gn), _stream()),
          "kge_loss_selfadv")
    return loss, gp, gn


def reg_fwd_bwd(desc, reg_type, lmbda, h, r, t, grad_scale=0.0, grad_tables=None):
    out = torch.empty(1, dtype=torch.float32, device=h.device)
    m = desc.c_struct()
    arr = _table_ptr_array(grad_tables) if grad_tables is not None else None
    check(lib().kge_reg_fwd_bwd(ctypes.byref(m), int(reg_type), ctypes.c_float(lmbda), _ptr(h), _ptr(r),
                                _ptr(t), ctypes.c_int64(h.numel()), _ptr(out), ctypes.c_float(grad_scale),
                                arr, _stream()), "kge_reg_fwd_bwd")
    return out


def train_pairwise_hinge_sgd(desc, grad_scratch, ph, pr, pt, nh, nr, nt, margin, lr, loss_out=None):
    """In-place fused step on desc.tables; grad_scratch: zero-filled dense buffers shaped
    like the tables (left zero-filled on return)."""
    if loss_out is None:
        loss_out = torch.empty(1, dtype=torch.float32, device=ph.device)
    m = desc.c_struct()
    rw = _table_ptr_array(desc.tables)
    gs = _table_ptr_array(grad_scratch)
    check(lib().kge_train_pairwise_hinge_sgd(ctypes.byref(m), rw, gs, _ptr(ph), _ptr(pr), _ptr(pt), _ptr(nh),
                                             _ptr(nr), _ptr(nt), ctypes.c_int64(ph.numel()),
                                             ctypes.c_float(margin), ctypes.c_float(lr), _ptr(loss_out),
                                             _stream()), "kge_train_pairwise_hinge_sgd")
    return loss_out


def train_pointwise_logistic(desc, grad_scratch, h, r, t, y, loss_out=None):
    """forward + Criterion.pointwise_logistic + backward of a pointwise batch in one kernel: returns the loss
    [1]; row gradients are accumulated into grad_scratch (dense buffers shaped like the tables)."""
    if loss_out is None:
        loss_out = torch.empty(1, dtype=torch.float32, device=h.device)
    m = desc.c_struct()
    gs = _table_ptr_array(grad_scratch)
    check(lib().kge_train_pointwise_logistic(ctypes.byref(m), gs, _ptr(_dev_i64(h, "h")), _ptr(_dev_i64(r, "r")),
                                             _ptr(_dev_i64(t, "t")), _ptr(_dev_i64(y, "y")), ctypes.c_int64(h.numel()),
                                             _ptr(loss_out), _stream()), "kge_train_pointwise_logistic")
    return loss_out


def train_pairwise_selfadv(desc, grad_scratch, ph, pr, pt, nh, nr, nt, neg_rate, alpha, loss_out=None):
    """RotatE: forward of the positives and their negatives + the self-adversarial loss + backward in one kernel:
    returns the loss [1]; row gradients are accumulated into grad_scratch.  KgeError (KGE_ENOTSUP) when neg_rate
    does not fit a CTA's shared memory — the caller then takes the unfused path."""
    if loss_out is None:
        loss_out = torch.empty(1, dtype=torch.float32, device=ph.device)
    if nh.numel() != ph.numel() * int(neg_rate):
        raise KgeError("self-adversarial loss: %d negatives for %d positives x neg_rate %d"
                       % (nh.numel(), ph.numel(), int(neg_rate)))
    m = desc.c_struct()
    gs = _table_ptr_array(grad_scratch)
    check(lib().kge_train_pairwise_selfadv(
        ctypes.byref(m), gs, _ptr(_dev_i64(ph, "ph")), _ptr(_dev_i64(pr, "pr")), _ptr(_dev_i64(pt, "pt")),
        _ptr(_dev_i64(nh, "nh")), _ptr(_dev_i64(nr, "nr")), _ptr(_dev_i64(nt, "nt")), ctypes.c_int64(ph.numel()),
        ctypes.c_int32(int(neg_rate)), ctypes.c_float(float(alpha)), _ptr(loss_out), _stream()),
        "kge_train_pairwise_selfadv")
    return loss_out


def optim_apply_rows(desc, grad_scratch, state, optimizer, h, r, t, lr, eps=1e-10):
    """optimizer: 0 SGD, 1 Adagrad.  Consumes (and re-zeroes) grad_scratch for the touched rows."""
    m = desc.c_struct()
    rw = _table_ptr_array(desc.tables)
    gs = _table_ptr_array(grad_scratch)
    stt = _table_ptr_array(state) if state is not None else None
    check(lib().kge_optim_apply_rows(ctypes.byref(m), rw, gs, stt, ctypes.c_int(optimizer), _ptr(h), _ptr(r),
                                     _ptr(t), ctypes.c_int64(h.numel()), ctypes.c_float(lr),
                                     ctypes.c_float(eps), _stream()), "kge_optim_apply_rows")


OPT_SGD, OPT_ADAGRAD, OPT_ADAM = 0, 1, 2


def optim_apply_dense(w, grad, optimizer, lr, state1=None, state2=None, eps=None, betas=(0.9, 0.999), step=1):
    """Dense optimizer step on ONE parameter tensor; consumes (and re-zeroes) its dense gradient buffer.
    optimizer: OPT_SGD / OPT_ADAGRAD (state1) / OPT_ADAM (state1 = exp_avg, state2 = exp_avg_sq, step >= 1)."""
    if eps is None:
        eps = 1e-8 if optimizer == OPT_ADAM else 1e-10
    w, grad = _dev_f32(w, "w"), _dev_f32(grad, "grad")
    if grad.numel() != w.numel():
        raise KgeError("gradient buffer must be shaped like the parameter")
    check(lib().kge_optim_apply_dense(_ptr(w), _ptr(grad), _ptr(state1), _ptr(state2), ctypes.c_int64(w.numel()),
                                      ctypes.c_int(optimizer), ctypes.c_float(lr), ctypes.c_float(eps),
                                      ctypes.c_float(betas[0]), ctypes.c_float(betas[1]), ctypes.c_int64(step),
                                      _stream()), "kge_optim_apply_dense")


def rank_workspace_bytes(desc, Q):
    m = desc.c_struct()
    return int(lib().kge_rank_workspace_bytes(ctypes.byref(m), ctypes.c_int64(Q)))


def rank_1vsall(desc, qh, qr, qt, filt_t=None, filt_h=None, counts=None, row_lo=0, row_hi=None,
                query_desc=None, tgt_h=None, tgt_t=None, flags=0, workspace=None):
    """Accumulate 0-based (tail raw, tail filtered, head raw, head filtered) rank counts of the
    queries over candidate rows [row_lo, row_hi) into counts [Q,4] int32.
    filt_* = (ptr[Q+1] int64 cuda, idx[nnz] int64 cuda) or None."""
    qh, qr, qt = _dev_i64(qh, "qh"), _dev_i64(qr, "qr"), _dev_i64(qt, "qt")
    Q = qh.numel()
    if row_hi is None:
        row_hi = row_lo + desc.num_ent
    if counts is None:
        counts = torch.zeros((Q, 4), dtype=torch.int32, device=qh.device)
    nbytes = rank_workspace_bytes(desc, Q)
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=qh.device)
    m = desc.c_struct()
    mq = query_desc.c_struct() if query_desc is not None else None
    ft_ptr, ft_idx = filt_t if filt_t is not None else (None, None)
    fh_ptr, fh_idx = filt_h if filt_h is not None else (None, None)
    check(lib().kge_rank_1vsall(
        ctypes.byref(m), ctypes.byref(mq) if mq is not None else None,
        ctypes.c_int64(row_lo), ctypes.c_int64(row_hi), _ptr(qh), _ptr(qr), _ptr(qt),
        _ptr(tgt_h), _ptr(tgt_t), ctypes.c_int64(Q),
        _ptr(ft_ptr), _ptr(ft_idx), ctypes.c_int64(ft_idx.numel() if ft_idx is not None else 0),
        _ptr(fh_ptr), _ptr(fh_idx), ctypes.c_int64(fh_idx.numel() if fh_idx is not None else 0),
        _ptr(counts), _ptr(workspace), ctypes.c_int64(workspace.numel()), ctypes.c_int(flags), _stream()),
        "kge_rank_1vsall")
    return counts


def rank_last_sweep_directions():
    """how many directions the launch timed by rank_last_sweep_ms(0) swept (2: the tensor-core sweep of a full
    rank call covers tail and head in one launch; 0: nothing profiled)."""
    return int(lib().kge_rank_last_sweep_directions())


def rank_last_sweep_ms(direction):
    """device time (ms) of the main sweep kernel of `direction` in the last rank_1vsall(flags |= RANK_PROFILE)."""
    ms = ctypes.c_float(0.0)
    check(lib().kge_rank_last_sweep_ms(ctypes.c_int(direction), ctypes.byref(ms)), "kge_rank_last_sweep_ms")
    return float(ms.value)


def rank_tc_probe(desc, qh, qr, qt, direction, want_dots=True, query_desc=None, row_lo=0, row_hi=None):
    """Level 1 (tensor cores) of the two-level exact sweep for ONE direction (include/kge_b200.h):
    -> (dots [Q, nc] raw accumulators or None, band, counts [Q,4]); band = (coef [Q,4] = centre, a, b, e per
    query, cn [nc] = norm bound per candidate): the pair (q, c) is certainly better when
    dots - centre > a + b cn + e cn^2, certainly not when dots - centre < -(a + b cn + e cn^2)."""
    qh, qr, qt = _dev_i64(qh, "qh"), _dev_i64(qr, "qr"), _dev_i64(qt, "qt")
    Q = qh.numel()
    if row_hi is None:
        row_hi = row_lo + desc.num_ent
    nc = row_hi - row_lo
    dots = torch.empty((Q, nc), dtype=torch.float32, device=qh.device) if want_dots else None
    tau = torch.empty(Q * 4 + nc, dtype=torch.float32, device=qh.device)
    counts = torch.zeros((Q, 4), dtype=torch.int32, device=qh.device)
    ws = torch.empty(max(rank_workspace_bytes(desc, Q), 16), dtype=torch.uint8, device=qh.device)
    m = desc.c_struct()
    mq = query_desc.c_struct() if query_desc is not None else None
    check(lib().kge_rank_tc_probe(
        ctypes.byref(m), ctypes.byref(mq) if mq is not None else None, ctypes.c_int64(row_lo), ctypes.c_int64(row_hi),
        _ptr(qh), _ptr(qr), _ptr(qt), ctypes.c_int64(Q), ctypes.c_int(direction), _ptr(dots), _ptr(tau),
        _ptr(counts), _ptr(ws), ctypes.c_int64(ws.numel()), _stream()), "kge_rank_tc_probe")
    return dots, (tau[:Q * 4].view(Q, 4), tau[Q * 4:]), counts


def project_entities(desc, r, out=None):
    """TransH / TransD: the projected row of every entity for relation r, [num_ent, dim]
    (include/kge_b200.h kge_project_entities) — TransE over [out, rel] then equals the model."""
    if out is None:
        width = desc.rel_dim if desc.name == "transr" else desc.dim
        out = torch.empty((desc.num_ent, width), dtype=torch.float32, device=desc.tables[0].device)
    m = desc.c_struct()
    check(lib().kge_project_entities(ctypes.byref(m), ctypes.c_int64(int(r)), _ptr(_dev_f32(out, "out")), _stream()),
          "kge_project_entities")
    return out


def normalize_rows_to(table, out=None):
    """F.normalize(table, dim=-1) in the canonical arithmetic into a new tensor (TransR's first
    normalisation of the relation rows, pairwise.py:430-432)."""
    t = _dev_f32(table, "table")
    if out is None:
        out = torch.empty_like(t)
    check(lib().kge_normalize_rows_to(_ptr(t), ctypes.c_int64(t.shape[0]), ctypes.c_int64(t.shape[1]),
                                      _ptr(_dev_f32(out, "out")), _stream()), "kge_normalize_rows_to")
    return out


def tripleset_build(h, r, t, num_ent, num_rel):
    """Hash set of the positive triples on the device -> uint64-as-int64 tensor [capacity]."""
    n = h.numel()
    cap = int(lib().kge_tripleset_capacity(ctypes.c_int64(n)))
    slots = torch.empty(cap, dtype=torch.int64, device=h.device)
    check(lib().kge_tripleset_build(_ptr(_dev_i64(h, "h")), _ptr(_dev_i64(r, "r")), _ptr(_dev_i64(t, "t")),
                                    ctypes.c_int64(n), _ptr(slots), ctypes.c_int64(cap), ctypes.c_int64(num_ent),
                                    ctypes.c_int64(num_rel), _stream()), "kge_tripleset_build")
    return slots


def sample_negatives(slots, ph, pr, pt, neg_rate, head_prob, num_ent, seed, step, layout=0, out=None):
    """layout 0 -> (nh, nr, nt) each [B*neg_rate]; layout 1 -> (h, r, t, y) each [B*(1+neg_rate)]."""
    B = ph.numel()
    n = B * neg_rate if layout == 0 else B * (1 + neg_rate)
    if out is None:
        out = torch.empty((4, n), dtype=torch.int64, device=ph.device)
    check(lib().kge_sample_negatives(_ptr(slots), ctypes.c_int64(slots.numel()), _ptr(_dev_i64(ph, "ph")),
                                     _ptr(_dev_i64(pr, "pr")), _ptr(_dev_i64(pt, "pt")), ctypes.c_int64(B),
                                     ctypes.c_int32(neg_rate), _ptr(head_prob), ctypes.c_int64(num_ent),
                                     ctypes.c_uint64(seed), ctypes.c_uint64(step), ctypes.c_int32(layout),
                                     _ptr(out[0]), _ptr(out[1]), _ptr(out[2]), _ptr(out[3]), _stream()),
          "kge_sample_negatives")
    return (out[0], out[1], out[2]) if layout == 0 else (out[0], out[1], out[2], out[3])


# ---- projection-model tail (include/kge_b200.h: kge_proj_*) ------------------------------------
def _bias_row(bias, N):
    if bias is None:
        return None
    b = _dev_f32(bias, "bias")
    if b.numel() != N:
        raise KgeError("bias must hold one value per entity (%d), got %d" % (N, b.numel()))
    return b


def proj_tail_fwd(x, ent, bias=None, out=None):
    """preds [B,N] = sigmoid(x . ent^T + bias)   (projection.py:100-102)."""
    x, ent = _dev_f32(x, "x"), _dev_f32(ent, "ent")
    if x.dim() != 2 or ent.dim() != 2 or x.shape[1] != ent.shape[1]:
        raise KgeError("x must be [B,k] and ent [N,k]")
    B, k = x.shape
    N = ent.shape[0]
    bias = _bias_row(bias, N)
    if out is None:
        out = torch.empty((B, N), dtype=torch.float32, device=x.device)
    check(lib().kge_proj_tail_fwd(_ptr(x), _ptr(ent), _ptr(bias), ctypes.c_int64(B), ctypes.c_int64(N),
                                  ctypes.c_int32(k), _ptr(out), _stream()), "kge_proj_tail_fwd")
    return out


def proj_tail_bwd(grad_preds, preds, x, ent, grad_x=None, grad_ent=None, grad_bias=None):
    """Accumulates into the given (zero-filled or partially filled) gradient buffers."""
    B, k = x.shape
    N = ent.shape[0]
    check(lib().kge_proj_tail_bwd(_ptr(_dev_f32(grad_preds, "grad_preds")), _ptr(_dev_f32(preds, "preds")),
                                  _ptr(_dev_f32(x, "x")), _ptr(_dev_f32(ent, "ent")), ctypes.c_int64(B),
                                  ctypes.c_int64(N), ctypes.c_int32(k), _ptr(grad_x), _ptr(grad_ent),
                                  _ptr(grad_bias), _stream()), "kge_proj_tail_bwd")


def proj_bce(preds, labels, label_scale=1.0, label_shift=0.0, grad_scale=1.0, want_grad=True):
    """One direction of Criterion.multi_class_bce -> (loss [1], grad_preds [B,N] or None)."""
    preds, labels = _dev_f32(preds, "preds"), _dev_f32(labels, "labels")
    if preds.dim() != 2 or preds.shape != labels.shape:
        raise KgeError("preds and labels must both be [B,N]")
    B, N = preds.shape
    loss = torch.empty(1, dtype=torch.float32, device=preds.device)
    g = torch.empty_like(preds) if want_grad else None
    check(lib().kge_proj_bce(_ptr(preds), _ptr(labels), ctypes.c_int64(B), ctypes.c_int64(N),
                             ctypes.c_float(label_scale), ctypes.c_float(label_shift), ctypes.c_float(grad_scale),
                             _ptr(loss), _ptr(g), _stream()), "kge_proj_bce")
    return loss, g


def proj_labels(rows, ptr, idx, B, N, out=None):
    """Dense [B,N] fp32 label rows (1.0 at the known positives) from a device CSR; rows = CSR row of
    each batch element (None: row b)."""
    ptr, idx = _dev_i64(ptr, "ptr"), _dev_i64(idx, "idx")
    if rows is not None and _dev_i64(rows, "rows").numel() != B:
        raise KgeError("rows must hold one CSR row id per batch element")
    if out is None:
        out = torch.empty((B, N), dtype=torch.float32, device=ptr.device)
    check(lib().kge_proj_labels(_ptr(rows), _ptr(ptr), _ptr(idx), ctypes.c_int64(B), ctypes.c_int64(N),
                                _ptr(_dev_f32(out, "labels")), _stream()), "kge_proj_labels")
    return out


def proj_rank(x, ent, bias, tgt, filt=None, direction=0, counts=None, workspace=None):
    """counts [Q,4] int32 += rank counts of direction 0 (tail: columns 0,1) or 1 (head: 2,3)."""
    x, ent, tgt = _dev_f32(x, "x"), _dev_f32(ent, "ent"), _dev_i64(tgt, "tgt")
    Q, k = x.shape
    N = ent.shape[0]
    bias = _bias_row(bias, N)
    if tgt.numel() != Q:
        raise KgeError("one target id per query row")
    if counts is None:
        counts = torch.zeros((Q, 4), dtype=torch.int32, device=x.device)
    nbytes = int(lib().kge_proj_rank_workspace_bytes(ctypes.c_int64(Q)))
    if workspace is None or workspace.numel() < nbytes:
        workspace = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=x.device)
    fp, fi = filt if filt is not None else (None, None)
    check(lib().kge_proj_rank(_ptr(x), _ptr(ent), _ptr(bias), ctypes.c_int64(Q), ctypes.c_int64(N),
                              ctypes.c_int32(k), _ptr(tgt), _ptr(fp), _ptr(fi),
                              ctypes.c_int64(fi.numel() if fi is not None else 0), ctypes.c_int32(direction),
                              _ptr(counts), _ptr(workspace), ctypes.c_int64(workspace.numel()), _stream()),
          "kge_proj_rank")
    return counts


class KgeConve(ctypes.Structure):
    """kge_conve_t of include/kge_b200.h."""
    _fields_ = [("hidden_size", ctypes.c_int32), ("hidden_size_1", ctypes.c_int32),
                ("bn0_eps", ctypes.c_float), ("bn1_eps", ctypes.c_float)] + \
               [(n, ctypes.c_void_p) for n in ("ent", "rel", "bn0_weight", "bn0_bias", "bn0_mean", "bn0_var",
                                               "conv_weight", "conv_bias", "bn1_weight", "bn1_bias", "bn1_mean",
                                               "bn1_var", "fc_weight", "fc_bias")]


def conve_trunk_fwd(model, e, r, out=None):
    """ConvE inference trunk (projection.py:104-112, :86-99, eval mode): x [Q,k] for entity ids e and
    relation ids r (already offset by tot_relation for the head direction).  `model` is a ConvE with
    the reference's sub-module names."""
    e, r = _dev_i64(e, "e"), _dev_i64(r, "r")
    Q = e.numel()
    if r.numel() != Q:
        raise KgeError("e and r must have equal length")
    tensors = {"ent": model.ent_embeddings.weight, "rel": model.rel_embeddings.weight,
               "bn0_weight": model.bn0.weight, "bn0_bias": model.bn0.bias,
               "bn0_mean": model.bn0.running_mean, "bn0_var": model.bn0.running_var,
               "conv_weight": model.conv2d_1.weight, "conv_bias": model.conv2d_1.bias,
               "bn1_weight": model.bn1.weight, "bn1_bias": model.bn1.bias,
               "bn1_mean": model.bn1.running_mean, "bn1_var": model.bn1.running_var,
               "fc_weight": model.fc.weight, "fc_bias": model.fc.bias}
    p = KgeConve()
    p.hidden_size, p.hidden_size_1 = int(model.hidden_size), int(model.hidden_size_1)
    p.bn0_eps, p.bn1_eps = float(model.bn0.eps), float(model.bn1.eps)
    for name, t in tensors.items():
        setattr(p, name, _dev_f32(t.detach(), name).data_ptr())
    nbytes = int(lib().kge_conve_trunk_workspace_bytes(ctypes.byref(p), ctypes.c_int64(Q)))
    ws = torch.empty(max(nbytes, 16), dtype=torch.uint8, device=e.device)
    if out is None:
        out = torch.empty((Q, p.hidden_size), dtype=torch.float32, device=e.device)
    check(lib().kge_conve_trunk_fwd(ctypes.byref(p), _ptr(e), _ptr(r), ctypes.c_int64(Q), _ptr(out), _ptr(ws),
                                    ctypes.c_int64(ws.numel()), _stream()), "kge_conve_trunk_fwd")
    return out


def launch_count():
    return int(lib().kge_launch_count())
