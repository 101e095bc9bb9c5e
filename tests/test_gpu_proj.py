"""-m gpu tests of the projection-model tail and ConvE, through the C-ABI:
kernels vs the CPU oracle (bit-exact where the arithmetic is canonical), vs a plain torch fp32
restatement of the same ops on the device, and the ConvE mirror vs the golden vectors the
reference itself produced (tests/golden/make_golden_proj.py)."""
import types

import numpy as np
import pytest
import torch

import golden_util as gu
import gpu_util as gpu
import oracle

pytestmark = pytest.mark.gpu
CASES = gu.proj_case_names()


def _L():
    from pykg2vec_b200 import _lib
    return _lib


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _case(B, N, k, seed, bias=True):
    rng = np.random.RandomState(seed)
    x = np.maximum(rng.standard_normal((B, k)) * 0.7, 0).astype(np.float32)
    ent = (rng.standard_normal((N, k)) * 0.5).astype(np.float32)
    b = (rng.standard_normal(N) * 0.3).astype(np.float32) if bias else None
    return x, ent, b


SHAPES = [(70, 131, 48, True), (5, 64, 50, True), (64, 65, 7, False), (1, 200, 16, True), (33, 1, 100, False),
          (128, 14541, 200, True), (257, 1000, 33, True)]


TILES = [None, "0", "1", "2"]   # KGE_PROJ_TILE: library heuristic, 64x64, 64x128, 128x128 CTA tiles


@pytest.mark.parametrize("tile", TILES, ids=lambda v: "tile-%s" % v)
@pytest.mark.parametrize("B,N,k,bias", SHAPES, ids=lambda v: str(v))
def test_tail_forward_bit_exact_vs_oracle(B, N, k, bias, tile, monkeypatch):
    if tile is not None:
        monkeypatch.setenv("KGE_PROJ_TILE", tile)
    L = _L()
    x, ent, b = _case(B, N, k, seed=B * 1000 + N, bias=bias)
    got = L.proj_tail_fwd(_cuda(x), _cuda(ent), _cuda(b) if bias else None).cpu().numpy()
    want = oracle.proj_tail_fwd(x, ent, b)
    assert np.array_equal(gpu.bits(got), gpu.bits(want))


def test_tail_forward_unaligned_operand_takes_scalar_loads():
    L = _L()
    B, N, k = 9, 77, 48
    x, ent, b = _case(B, N, k, seed=5)
    buf = torch.zeros(B * k + 1, dtype=torch.float32, device="cuda")
    xo = buf[1:].view(B, k)
    xo.copy_(_cuda(x))
    assert xo.data_ptr() % 16 != 0 and xo.is_contiguous()
    got = L.proj_tail_fwd(xo, _cuda(ent), _cuda(b)).cpu().numpy()
    assert np.array_equal(got, oracle.proj_tail_fwd(x, ent, b))


def test_tail_forward_vs_torch_fp32():
    """the same op in plain torch on the device: <= 1e-4 relative (north_star tolerance)"""
    L = _L()
    x, ent, b = _case(96, 5000, 200, seed=11)
    xd, ed, bd = _cuda(x), _cuda(ent), _cuda(b)
    got = L.proj_tail_fwd(xd, ed, bd)
    want = torch.sigmoid(torch.matmul(xd.double(), ed.double().T) + bd.double())
    assert ((got.double() - want).abs() / want).max().item() < 1e-5


@pytest.mark.parametrize("tile", TILES, ids=lambda v: "tile-%s" % v)
@pytest.mark.parametrize("B,N,k,bias", SHAPES[:4] + [(300, 14541, 200, True)], ids=lambda v: str(v))
def test_rank_counts_exact_vs_oracle(B, N, k, bias, tile, monkeypatch):
    if tile is not None:
        monkeypatch.setenv("KGE_PROJ_TILE", tile)
    L = _L()
    x, ent, b = _case(B, N, k, seed=B * 77 + N, bias=bias)
    rng = np.random.RandomState(B + N)
    tgt = rng.randint(N, size=B).astype(np.int64)
    dummy = np.zeros(B, dtype=np.int64)
    (ptr, idx), _ = gpu.random_filters_csr(rng, N, dummy, dummy, tgt, per_query=min(12, N))
    xd, ed, bd = _cuda(x), _cuda(ent), (_cuda(b) if bias else None)
    counts = torch.zeros((B, 4), dtype=torch.int32, device="cuda")
    want = np.zeros((B, 4), dtype=np.int32)
    for direction in (0, 1):
        L.proj_rank(xd, ed, bd, _cuda(tgt), (_cuda(ptr), _cuda(idx)), direction, counts)
        oracle.proj_rank(x, ent, b, tgt, (ptr, idx), direction, want)
    got = counts.cpu().numpy()
    assert np.array_equal(got, want)
    # the counts are what counting over the forward() matrix gives (same bits on both paths)
    preds = L.proj_tail_fwd(xd, ed, bd)
    raw = (preds > preds.gather(1, _cuda(tgt)[:, None])).sum(1).cpu().numpy()
    assert np.array_equal(got[:, 0], raw) and np.array_equal(got[:, 2], raw)
    assert (got[:, 1] <= got[:, 0]).all() and (got[:, 1] >= 0).all()


def test_rank_without_filters_and_accumulation():
    L = _L()
    x, ent, b = _case(40, 300, 64, seed=3)
    tgt = np.arange(40, dtype=np.int64)
    counts = torch.full((40, 4), 5, dtype=torch.int32, device="cuda")     # accumulate semantics
    L.proj_rank(_cuda(x), _cuda(ent), _cuda(b), _cuda(tgt), None, 0, counts)
    want = oracle.proj_rank(x, ent, b, tgt, None, 0)
    got = counts.cpu().numpy()
    assert np.array_equal(got[:, 0] - 5, want[:, 0]) and np.array_equal(got[:, 1] - 5, want[:, 0])
    assert (got[:, 2:] == 5).all()


@pytest.mark.parametrize("B,N", [(7, 131), (64, 3000), (128, 14541)])
def test_bce_value_and_gradient(B, N):
    L = _L()
    rng = np.random.RandomState(B * N)
    preds = (1.0 / (1.0 + np.exp(-rng.standard_normal((B, N)) * 3))).astype(np.float32)
    labels = (rng.rand(B, N) < 0.05).astype(np.float32)
    scale, shift = np.float32(1.0 - 0.1), np.float32(1.0 / N)
    loss, g = L.proj_bce(_cuda(preds), _cuda(labels), float(scale), float(shift), 1.0)
    want_loss, want_g = oracle.proj_bce(preds, labels, scale, shift, 1.0)
    assert abs(loss.item() - want_loss) <= 5e-6 * abs(want_loss)
    assert np.array_equal(gpu.bits(g.cpu().numpy()), gpu.bits(want_g))
    # and against torch's own BCEWithLogits on the device (what the reference calls)
    pt = _cuda(preds).requires_grad_()
    ref = torch.mean(torch.nn.BCEWithLogitsLoss()(pt, _cuda(labels) * float(scale) + float(shift)))
    ref.backward()
    assert abs(loss.item() - ref.item()) <= 5e-6 * abs(ref.item())
    assert (g - pt.grad).abs().max().item() <= 2e-6 * pt.grad.abs().max().item()


@pytest.mark.parametrize("B,N,k", [(70, 131, 48), (5, 300, 50), (64, 65, 7), (128, 14541, 200)])
def test_tail_backward_vs_oracle_and_torch(B, N, k):
    L = _L()
    x, ent, b = _case(B, N, k, seed=B + 3 * N)
    rng = np.random.RandomState(k)
    gp = (rng.standard_normal((B, N)) * 0.1).astype(np.float32)
    xd, ed, bd, gpd = _cuda(x), _cuda(ent), _cuda(b), _cuda(gp)
    preds = L.proj_tail_fwd(xd, ed, bd)
    gx = torch.zeros_like(xd)
    ge = torch.full_like(ed, 0.25)                         # accumulate semantics: pre-filled
    gb = torch.full((N,), -0.5, dtype=torch.float32, device="cuda")
    L.proj_tail_bwd(gpd, preds, xd, ed, gx, ge, gb)
    # torch fp64 restatement of the same ops on the device
    x64, e64, b64 = xd.double().requires_grad_(), ed.double().requires_grad_(), bd.double().requires_grad_()
    p64 = torch.sigmoid(x64 @ e64.T + b64)
    (p64 * gpd.double()).sum().backward()
    for got, want in ((gx, x64.grad), (ge - 0.25, e64.grad), (gb + 0.5, b64.grad)):
        err = (got.double() - want).abs().max().item()
        assert err <= 2e-5 * max(want.abs().max().item(), 1e-3), err
    if B * N * k < 5_000_000:   # the double-precision C oracle as well (small cases)
        wx, we, wb = oracle.proj_tail_bwd(gp, preds.cpu().numpy(), x, ent)
        for got, want in ((gx, wx), (ge - 0.25, we), (gb + 0.5, wb)):
            assert np.abs(got.cpu().numpy() - want).max() <= 5e-6 * max(1.0, np.abs(want).max())


def test_tail_autograd_function_matches_torch_ops():
    from pykg2vec_b200.criterion import Criterion
    from pykg2vec_b200.functional import ProjTailFunction
    B, N, k = 24, 211, 100
    x, ent, b = _case(B, N, k, seed=21)
    rng = np.random.RandomState(4)
    lab_t = (rng.rand(B, N) < 0.03).astype(np.float32)
    lab_h = (rng.rand(B, N) < 0.03).astype(np.float32)
    leaves = [_cuda(x).requires_grad_(), _cuda(ent).requires_grad_(), _cuda(b).view(1, N).requires_grad_()]
    ref = [t.detach().clone().requires_grad_() for t in leaves]
    p1 = ProjTailFunction.apply(leaves[0], leaves[1], leaves[2])
    p2 = ProjTailFunction.apply(leaves[0] * 0.5, leaves[1], leaves[2])
    loss = Criterion.multi_class_bce(p1, p2, _cuda(lab_h), _cuda(lab_t), 0.1, N)
    loss.backward()
    q1 = torch.sigmoid(ref[0] @ ref[1].T + ref[2])
    q2 = torch.sigmoid((ref[0] * 0.5) @ ref[1].T + ref[2])
    bce = torch.nn.BCEWithLogitsLoss()
    want = torch.mean(bce(q1, _cuda(lab_h) * (1.0 - 0.1) + 1.0 / N)) + torch.mean(bce(q2, _cuda(lab_t) * (1.0 - 0.1) + 1.0 / N))
    want.backward()
    assert abs(loss.item() - want.item()) <= 5e-6 * abs(want.item())
    for a, w in zip(leaves, ref):
        assert (a.grad - w.grad).abs().max().item() <= 2e-4 * w.grad.abs().max().item()


# ---- ConvE: trunk kernel and the mirror vs the reference's golden vectors -------------------------
def _mirror(g, train=False):
    from pykg2vec_b200 import import_model
    m = import_model("conve")(tot_entity=int(g["N"]), tot_relation=int(g["R"]), hidden_size=int(g["hidden_size"]),
                              hidden_size_1=int(g["hidden_size_1"]), lmbda=0.1, input_dropout=0.0,
                              feature_map_dropout=0.0, hidden_dropout=0.0)
    m.load_state_dict({k_: torch.from_numpy(np.asarray(v)) for k_, v in gu.proj_state(g).items()}, strict=True)
    m.cuda()
    m.train(train)
    return m


@pytest.mark.parametrize("name", CASES)
def test_conve_trunk_kernel_bit_exact_vs_oracle(name):
    L = _L()
    g = gu.load(name)
    m = _mirror(g)
    R = int(g["R"])
    st = gu.proj_state(g)
    for e, r in ((g["h"], g["r"]), (g["t"], g["r"] + R)):
        got = L.conve_trunk_fwd(m, _cuda(e), _cuda(r)).cpu().numpy()
        want = oracle.conve_trunk_fwd(st, int(g["hidden_size"]), int(g["hidden_size_1"]), e, r)
        assert np.array_equal(gpu.bits(got), gpu.bits(want))


@pytest.mark.parametrize("name", CASES)
def test_conve_eval_forward_matches_reference(name):
    """model.forward(e, r, direction) in eval mode vs the reference's output: <= 1e-4 relative"""
    g = gu.load(name)
    m = _mirror(g)
    h, r, t = _cuda(g["h"]), _cuda(g["r"]), _cuda(g["t"])
    with torch.no_grad():
        pt = m(h, r, direction="tail").cpu().numpy()
        ph = m(t, r, direction="head").cpu().numpy()
        xt = m.proj_query(h, r, "tail").cpu().numpy()
    assert gu.rel_err(pt, g["preds_tail"]).max() < 1e-4 and gu.rel_err(ph, g["preds_head"]).max() < 1e-4
    assert np.abs(xt - g["x_tail"]).max() < 1e-5
    # grad-enabled eval forward goes through the torch layers: same predictions
    pt2 = m(h, r, direction="tail").detach().cpu().numpy()
    assert gu.rel_err(pt2, g["preds_tail"]).max() < 1e-4
    # single-query API of the reference (predict_tail_rank -> ids sorted by descending prediction)
    with torch.no_grad():
        rank = m.predict_tail_rank(h[:1], r[:1], topk=int(g["N"])).cpu().numpy()[0]
    assert rank[-1] == int(np.argmax(pt[0])) and rank[0] == int(np.argmin(pt[0]))   # walked from the end


@pytest.mark.parametrize("name", CASES)
def test_conve_evaluator_ranks_match_reference(name):
    from pykg2vec_b200.evaluator import Evaluator, build_filter_csr
    g = gu.load(name)
    m = _mirror(g)
    Q = g["ranks"].shape[0]
    ev = object.__new__(Evaluator)
    ev.model, ev.config = m, types.SimpleNamespace(device="cuda", tot_entity=int(g["N"]))
    ev._filter_cache, ev._workspace = {}, None
    ft = (g["filt_t_ptr"], g["filt_t_idx"])
    fh = (g["filt_h_ptr"], g["filt_h_idx"])
    with torch.no_grad():
        got = ev.rank_triples(g["h"][:Q], g["r"][:Q], g["t"][:Q], ft, fh)
        raw = ev.rank_triples(g["h"][:Q], g["r"][:Q], g["t"][:Q])
    assert np.array_equal(got, g["ranks"])
    assert np.array_equal(raw[:, 0], g["ranks"][:, 0]) and np.array_equal(raw[:, 1], g["ranks"][:, 0])
    assert ev.last_d2h_bytes == Q * 16 and ev.last_h2d_bytes > 0


@pytest.mark.parametrize("name", CASES)
def test_conve_training_step_matches_reference_autograd(name):
    """train_step_projection (trainer.py:159-174): loss and the gradient of every parameter vs the
    reference's own backward; BatchNorm running statistics after the two forwards."""
    from pykg2vec_b200.trainer import Trainer
    g = gu.load(name)
    m = _mirror(g, train=True)
    cfg = types.SimpleNamespace(device="cuda", label_smoothing=float(g["label_smoothing"]), tot_entity=int(g["N"]))
    tr = object.__new__(Trainer)
    tr.model, tr.config = m, cfg
    m.zero_grad()
    loss = tr.train_step_projection(_cuda(g["h"]), _cuda(g["r"]), _cuda(g["t"]), _cuda(g["tr_labels_tail"]),
                                    _cuda(g["tr_labels_head"]))
    loss.backward()
    assert abs(loss.item() - float(g["tr_loss"])) <= 1e-5 * abs(float(g["tr_loss"]))
    for key, p in m.named_parameters():
        got = p.grad.detach().cpu().numpy()
        if "grad_" + key in g:
            want = g["grad_" + key]
        else:
            got = got.reshape(-1)[::37]
            want = g["gradsample_" + key]
        # bn0's scale and shift are (mathematically) invisible behind bn1's batch normalisation: their
        # true gradient is 0 and what autograd returns is rounding noise of order 1e-7 — hence the floor
        # (likewise conv / fc biases in front of a training-mode BatchNorm): relative bound + a floor
        assert np.abs(got - want).max() <= 5e-4 * np.abs(want).max() + 1e-7, key
    for key, v in m.state_dict().items():
        if "running" in key:
            assert np.abs(v.cpu().numpy() - g["sd_after_" + key]).max() < 1e-5, key


def test_conve_trainer_batches_reduce_the_loss():
    """Trainer.train_batch with the PROJECTION_BASED strategy: host ids + dense labels in, adam steps"""
    from pykg2vec_b200 import import_model
    from pykg2vec_b200.trainer import Trainer
    N, R, k, B = 150, 4, 48, 32
    torch.manual_seed(0)
    m = import_model("conve")(tot_entity=N, tot_relation=R, hidden_size=k, hidden_size_1=8, lmbda=0.1,
                              input_dropout=0.0, feature_map_dropout=0.0, hidden_dropout=0.0)
    kgraph = types.SimpleNamespace(read_cache_data=lambda key: {})
    cfg = types.SimpleNamespace(device="cuda", optimizer="adam", learning_rate=0.01, label_smoothing=0.1,
                                tot_entity=N, tot_relation=R, knowledge_graph=kgraph)
    tr = Trainer(m, cfg)
    tr.build_model()
    rng = np.random.RandomState(1)
    h, r, t = rng.randint(N, size=B), rng.randint(R, size=B), rng.randint(N, size=B)
    hr_t = np.zeros((B, N), dtype=np.float32)
    tr_h = np.zeros((B, N), dtype=np.float32)
    hr_t[np.arange(B), t] = 1.0
    tr_h[np.arange(B), h] = 1.0
    losses = [tr.train_batch([h, r, t, torch.from_numpy(hr_t), torch.from_numpy(tr_h)]) for _ in range(8)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    assert tr.last_h2d_bytes >= 2 * B * N * 4


@pytest.mark.parametrize("with_rows", [False, True])
def test_label_rows_kernel(with_rows):
    L = _L()
    rng = np.random.RandomState(9)
    K, N, B = 40, 14541, 33
    sizes = rng.randint(0, 600, size=K)
    sizes[3] = 0
    ptr = np.concatenate([[0], np.cumsum(sizes)]).astype(np.int64)
    idx = np.concatenate([rng.choice(N, size=n, replace=False) for n in sizes]).astype(np.int64)
    rows = rng.randint(K, size=B).astype(np.int64) if with_rows else None
    if not with_rows:
        B = K
    got = L.proj_labels(_cuda(rows) if with_rows else None, _cuda(ptr), _cuda(idx), B, N).cpu().numpy()
    want = np.zeros((B, N), dtype=np.float32)
    for b in range(B):
        row = rows[b] if with_rows else b
        want[b, idx[ptr[row]:ptr[row + 1]]] = 1.0
    assert np.array_equal(got, want)


def test_conve_epoch_on_device_generator():
    """Generator (device label rows) -> Trainer.train_model_epoch -> batched Evaluator, ConvE, adam"""
    from pykg2vec_b200 import import_model
    from pykg2vec_b200.generator import Generator
    from pykg2vec_b200.synthetic import SyntheticConfig, SyntheticKnowledgeGraph
    from pykg2vec_b200.trainer import Trainer
    kg = SyntheticKnowledgeGraph(200, 5, 2000, 60, 60, seed=3)
    cfg = SyntheticConfig(kg, device="cuda", optimizer="adam", learning_rate=0.003, batch_size=64, neg_rate=0,
                          hidden_size=48, hidden_size_1=8, lmbda=0.1, input_dropout=0.0, feature_map_dropout=0.0,
                          hidden_dropout=0.0, label_smoothing=0.1, test_num=50)
    torch.manual_seed(0)
    model = import_model("conve")(**cfg.__dict__)
    tr = Trainer(model, cfg)
    tr.build_model()
    gen = Generator(model, cfg, seed=1)
    # the label rows of a batch are the dense hr_t_train / tr_h_train rows of its triples
    gen.start_one_epoch(1)
    h, r, t, hr_t, tr_h = next(gen)
    known = kg.read_cache_data("hr_t_train")
    hn, rn = h.cpu().numpy(), r.cpu().numpy()
    lab = hr_t.cpu().numpy()
    assert lab.shape == (64, 200) and tr_h.shape == (64, 200)
    for i in range(64):
        assert set(np.nonzero(lab[i])[0].tolist()) == known[(int(hn[i]), int(rn[i]))]
    losses = [tr.train_model_epoch(gen, num_batch=10) for _ in range(4)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    model.eval()
    scores = tr.evaluator.mini_test(epoch=0)
    assert set(scores) == {"mr", "fmr", "mrr", "fmrr"} and 1.0 <= scores["fmr"] <= scores["mr"] <= 200.0


def _tucker(g, train=False):
    from pykg2vec_b200 import import_model
    m = import_model("tucker")(tot_entity=int(g["N"]), tot_relation=int(g["R"]),
                               ent_hidden_size=int(g["ent_hidden_size"]), rel_hidden_size=int(g["rel_hidden_size"]),
                               lmbda=0.1, input_dropout=0.0, hidden_dropout1=0.0, hidden_dropout2=0.0)
    m.load_state_dict({k_: torch.from_numpy(np.asarray(v)) for k_, v in gu.proj_state(g).items()}, strict=True)
    return m.cuda().train(train)


def test_tucker_matches_reference():
    """TuckER (bias-free tail, torch trunk): predictions, Evaluator ranks, training loss and gradients
    vs the reference's own outputs"""
    from pykg2vec_b200.evaluator import Evaluator
    from pykg2vec_b200.trainer import Trainer
    g = gu.load("tucker_d32")
    m = _tucker(g)
    h, r, t = _cuda(g["h"]), _cuda(g["r"]), _cuda(g["t"])
    with torch.no_grad():
        assert gu.rel_err(m(h, r, direction="tail").cpu().numpy(), g["preds_tail"]).max() < 1e-4
        assert gu.rel_err(m(t, r, direction="head").cpu().numpy(), g["preds_head"]).max() < 1e-4
    Q = g["ranks"].shape[0]
    ev = object.__new__(Evaluator)
    ev.model, ev.config = m, types.SimpleNamespace(device="cuda", tot_entity=int(g["N"]))
    ev._filter_cache, ev._workspace = {}, None
    with torch.no_grad():
        got = ev.rank_triples(g["h"][:Q], g["r"][:Q], g["t"][:Q], (g["filt_t_ptr"], g["filt_t_idx"]),
                              (g["filt_h_ptr"], g["filt_h_idx"]))
    assert np.array_equal(got, g["ranks"])
    m = _tucker(g, train=True)
    tr = object.__new__(Trainer)
    tr.model = m
    tr.config = types.SimpleNamespace(device="cuda", label_smoothing=float(g["label_smoothing"]), tot_entity=int(g["N"]))
    m.zero_grad()
    loss = tr.train_step_projection(h, r, t, _cuda(g["tr_labels_tail"]), _cuda(g["tr_labels_head"]))
    loss.backward()
    assert abs(loss.item() - float(g["tr_loss"])) <= 1e-5 * abs(float(g["tr_loss"]))
    for key, p in m.named_parameters():
        want = g["grad_" + key]
        assert np.abs(p.grad.cpu().numpy() - want).max() <= 5e-4 * np.abs(want).max() + 1e-7, key
