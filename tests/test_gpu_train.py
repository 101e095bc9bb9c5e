"""-m gpu tests of the training-side kernels: backward vs the fp64 autograd oracle
(oracle/ref_port.py) and vs gradients produced by the reference itself (golden), losses,
regularisers, the model classes through autograd, and the fused sparse training steps vs
dense torch optimizers on the oracle formulas."""
import numpy as np
import pytest
import torch

import golden_util as gu
import gpu_util as gpu

pytestmark = pytest.mark.gpu
CASES = [n for n in gu.case_names() if "pretrained" not in n]
GRAD_TOL = 2e-4  # relative to the largest |gradient| of the table


def _L():
    from pykg2vec_b200 import _lib
    return _lib


def _cuda(a):
    return torch.from_numpy(np.ascontiguousarray(a)).cuda()


def _check_grads(got, want, tol=GRAD_TOL, what=""):
    for k, (g, w) in enumerate(zip(got, want)):
        if w is None:
            continue
        w = np.asarray(w, dtype=np.float64)
        scale = max(np.abs(w).max(), 1e-12)
        err = np.abs(g.astype(np.float64) - w).max() / scale
        assert err < tol, "%s table %d: rel err %.3g" % (what, k, err)


@pytest.mark.parametrize("name", CASES)
def test_score_bwd_vs_reference_autograd(name):
    L = _L()
    g = gu.load(name)
    desc = gpu.desc_from_golden(g)
    grads = [torch.zeros_like(t) for t in desc.tables]
    L.score_bwd(desc, _cuda(g["h"]), _cuda(g["r"]), _cuda(g["t"]), _cuda(g["upstream"]), grads)
    want = [g.get("grad%d" % k) for k in range(len(grads))]   # (ConvKB: no reference gradient for the collapsed tables)
    _check_grads([x.cpu().numpy() for x in grads], want, what=name)


@pytest.mark.parametrize("spec", [
    ("transe", 300, 7, 200, None, False, 0.0), ("transe", 300, 7, 50, None, True, 0.0),
    ("transh", 300, 7, 100, None, True, 0.0), ("transd", 300, 7, 64, None, False, 0.0),
    ("transr", 120, 5, 40, 24, False, 0.0), ("transm", 300, 7, 36, None, False, 0.0),
    ("rotate", 300, 7, 100, None, False, 12.0), ("distmult", 300, 7, 200, None, False, 0.0),
    ("cp", 300, 7, 30, None, False, 0.0), ("complex", 300, 7, 200, None, False, 0.0),
    ("hole", 200, 5, 30, None, False, 0.0), ("hole", 200, 5, 52, None, False, 0.0),
    ("rescal", 150, 4, 24, None, False, 0.0), ("simple", 300, 7, 48, None, False, 0.0),
    ("simple_ignr", 300, 7, 50, None, False, 0.0), ("analogy", 300, 7, 48, None, False, 0.0),
    ("slm", 200, 5, 24, 16, False, 0.0), ("ntn", 120, 4, 16, 12, False, 0.0), ("sme", 150, 5, 24, None, False, 0.0),
    ("sme_bl", 150, 5, 20, None, False, 0.0), ("kg2e", 200, 5, 40, None, False, 0.0), ("quate", 200, 5, 24, None, False, 0.0), ("octonione", 150, 5, 12, None, False, 0.0),
    ("convkb", 200, 5, 50, None, False, 0.0), ("convkb", 150, 5, 37, None, False, 0.0),
], ids=lambda s: "%s-d%d" % (s[0], s[3]))
def test_score_bwd_vs_fp64_oracle(spec):
    """duplicates in the batch (few entities) exercise the atomic scatter."""
    from oracle import ref_port
    L = _L()
    name, N, R, d, dr, l1, margin = spec
    om, tabs = gpu.synthetic_case(name, N, R, d, seed=11, dr=dr, l1=l1, margin=margin, scale=0.3)
    desc = gpu.desc_from_oracle_model(om)
    rng = np.random.RandomState(1)
    n = 777
    h, r, t = rng.randint(N, size=n), rng.randint(R, size=n), rng.randint(N, size=n)
    up = rng.standard_normal(n).astype(np.float32)
    grads = [torch.zeros_like(x) for x in desc.tables]
    L.score_bwd(desc, _cuda(h), _cuda(r), _cuda(t), _cuda(up), grads)
    t64 = [torch.from_numpy(x.astype(np.float64)).requires_grad_(name != "transm" or i < 2)
           for i, x in enumerate(tabs)]
    s = ref_port.score(name, t64, torch.from_numpy(h), torch.from_numpy(r), torch.from_numpy(t),
                       l1_flag=l1, margin=margin, embedding_range=((margin + 2.0) / d if name == "rotate" else None),
                       rel_dim=dr)
    (s * torch.from_numpy(up.astype(np.float64))).sum().backward()
    want = [x.grad.numpy() if x.grad is not None else None for x in t64]
    _check_grads([x.cpu().numpy() for x in grads], want, tol=5e-5, what=name)


def test_losses_vs_golden_and_oracle():
    import oracle
    L = _L()
    g = gu.load("losses")
    loss, gp, gn = L.loss_pairwise_hinge(_cuda(g["pos"]), _cuda(g["neg"]), float(g["margin"]))
    assert abs(loss.item() - g["hinge"]) <= 1e-5 * abs(g["hinge"])
    np.testing.assert_array_equal(gp.cpu().numpy(), g["hinge_gpos"])
    np.testing.assert_array_equal(gn.cpu().numpy(), g["hinge_gneg"])
    loss, gg = L.loss_pointwise_logistic(_cuda(g["preds"]), _cuda(g["target"]))
    assert abs(loss.item() - g["logistic"]) <= 1e-5 * abs(g["logistic"])
    np.testing.assert_allclose(gg.cpu().numpy(), g["logistic_g"], rtol=1e-4, atol=1e-8)
    for tag, alpha in (("a1", 1.0), ("a01", 0.1)):
        loss, gp, gn = L.loss_selfadv(_cuda(g["sa_pos"]), _cuda(g["sa_neg"]), int(g["sa_neg_rate"]), alpha)
        assert abs(loss.item() - g["sa_" + tag]) <= 2e-5 * abs(g["sa_" + tag])
        np.testing.assert_allclose(gp.cpu().numpy(), g["sa_%s_gpos" % tag], rtol=1e-4, atol=1e-8)
        np.testing.assert_allclose(gn.cpu().numpy(), g["sa_%s_gneg" % tag], rtol=1e-4, atol=1e-8)
    # larger, seeded: vs the C oracle
    rng = np.random.RandomState(3)
    pos = (rng.standard_normal(5000) * 2).astype(np.float32)
    neg = (rng.standard_normal(5000) * 2).astype(np.float32)
    want, _ = oracle.loss_pairwise_hinge(pos, neg, 1.5)
    got, _, _ = L.loss_pairwise_hinge(_cuda(pos), _cuda(neg), 1.5)
    assert abs(got.item() - want) <= 1e-5 * abs(want)
    neg2 = (rng.standard_normal(5000 * 16) * 3).astype(np.float32)
    want = oracle.loss_selfadv(pos, neg2, 16, 0.5)
    got, _, _ = L.loss_selfadv(_cuda(pos), _cuda(neg2), 16, 0.5)
    assert abs(got.item() - want) <= 2e-5 * abs(want)


@pytest.mark.parametrize("name", [n for n in CASES if n.split("_")[0] in ("distmult", "complex", "cp", "analogy", "quate", "octonione")])
def test_regulariser_vs_golden_and_autograd(name):
    from oracle import ref_port
    L = _L()
    g = gu.load(name)
    desc = gpu.desc_from_golden(g)
    h, r, t = _cuda(g["h"]), _cuda(g["r"]), _cuda(g["t"])
    lm = float(g["kw_lmbda"])
    for code, key in ((0, "reg_f2"), (1, "reg_n3"), (2, "reg_absn3")):
        if key not in g:
            continue
        grads = [torch.zeros_like(x) for x in desc.tables]
        out = L.reg_fwd_bwd(desc, code, lm, h, r, t, grad_scale=1.0, grad_tables=grads)
        assert abs(out.item() - g[key]) <= 1e-4 * abs(g[key]) + 1e-9, (name, key)
        t64 = [torch.from_numpy(x.astype(np.float64)).requires_grad_() for x in gu.tables_of(g)]
        ref_port.reg(str(g["model"]), t64, torch.from_numpy(g["h"]), torch.from_numpy(g["r"]),
                     torch.from_numpy(g["t"]), lm, code).backward()
        _check_grads([x.cpu().numpy() for x in grads], [x.grad.numpy() for x in t64], tol=5e-5, what=name + key)


def _make_model(g, device="cuda"):
    """the pykg2vec_b200 model class for a golden case, loaded with the golden tables through
    the reference's state_dict keys."""
    import pykg2vec_b200
    kw = {k[3:]: (g[k].item() if g[k].ndim == 0 else g[k]) for k in g if k.startswith("kw_")}
    cls = pykg2vec_b200.import_model(str(g["model"]))
    if str(g["model"]) == "convkb":
        kw["device"] = device   # the reference's ConvKB takes the device of its (unregistered) conv_list
    m = cls(tot_entity=int(g["N"]), tot_relation=int(g["R"]), **kw)
    sd = {str(k) + ".weight": torch.from_numpy(g["table%d" % i]) for i, k in enumerate(g["table_keys"])}
    if str(g["model"]) == "convkb":
        raw = gu.raw_tables_of(g)
        sd["fc1.weight"], sd["fc1.bias"] = torch.from_numpy(raw[-2]), torch.from_numpy(raw[-1])
        with torch.no_grad():
            for i, conv in enumerate(m.conv_list):   # plain Python list: not part of state_dict
                conv.weight.copy_(torch.from_numpy(raw[2 + 2 * i]))
                conv.bias.copy_(torch.from_numpy(raw[3 + 2 * i]))
    missing = m.load_state_dict(sd, strict=False)  # (QuatE/OctonionE register tables forward() never reads)
    assert not missing.unexpected_keys
    return m.to(device)


@pytest.mark.parametrize("name", CASES)
def test_model_classes_forward_backward_like_reference(name):
    """Drop-in surface: state_dict keys load, forward() scores and .backward() dense grads
    match what the reference classes produced."""
    g = gu.load(name)
    if str(g["model"]) == "transm":
        pytest.skip("needs a knowledge graph")
    if str(g["model"]) == "rescal":
        pytest.skip("forward() re-normalises the stored (already normalised) tables: covered by test_rescal_model")
    m = _make_model(g)
    h, r, t = _cuda(g["h"]), _cuda(g["r"]), _cuda(g["t"])
    s = m(h, r, t)
    ref = g["scores"]
    floor = 1e-2 * np.abs(ref).max()
    err = np.abs(s.detach().cpu().numpy().astype(np.float64) - ref) / np.maximum(np.abs(ref), floor)
    assert err.max() < 1e-4
    (s * _cuda(g["upstream"])).sum().backward()
    got = [getattr(m, str(k)).weight.grad.cpu().numpy() for k in g["table_keys"]]
    _check_grads(got, [g["grad%d" % i] for i in range(len(got))], what=name)
    if str(g["model"]) == "convkb":   # the Linear layer trains through the collapse
        raw_n = len(gu.raw_tables_of(g))
        _check_grads([m.fc1.weight.grad.cpu().numpy(), m.fc1.bias.grad.cpu().numpy()],
                     [g["rawgrad%d" % (raw_n - 2)], g["rawgrad%d" % (raw_n - 1)]], what=name + " fc1")
        assert "conv_list" not in "".join(m.state_dict().keys())
    embs = m.embed(h, r, t)
    assert all(e.shape[0] == h.numel() for e in embs)


def test_fused_hinge_sgd_matches_dense_sgd():
    """kge_train_pairwise_hinge_sgd == (oracle formulas + torch autograd + optim.SGD)."""
    from oracle import ref_port
    L = _L()
    for name, d, l1 in (("transe", 200, False), ("transe", 50, True), ("transh", 48, False), ("transd", 40, True)):
        N, R, B = 500, 9, 512
        om, tabs = gpu.synthetic_case(name, N, R, d, seed=21, l1=l1, scale=0.4)
        desc = gpu.desc_from_oracle_model(om)
        scratch = [torch.zeros_like(x) for x in desc.tables]
        ref = [torch.from_numpy(x.astype(np.float64)).requires_grad_() for x in tabs]
        opt = torch.optim.SGD(ref, lr=0.05)
        rng = np.random.RandomState(4)
        for step in range(3):
            ids = [rng.randint(N if k % 3 != 1 else R, size=B) for k in range(6)]
            loss = L.train_pairwise_hinge_sgd(desc, scratch, *[_cuda(x) for x in ids], margin=0.7, lr=0.05)
            opt.zero_grad()
            tid = [torch.from_numpy(x) for x in ids]
            pos = ref_port.score(name, ref, tid[0], tid[1], tid[2], l1_flag=l1)
            neg = ref_port.score(name, ref, tid[3], tid[4], tid[5], l1_flag=l1)
            want = ref_port.pairwise_hinge(pos, neg, 0.7)
            want.backward()
            opt.step()
            assert abs(loss.item() - want.item()) <= 2e-4 * abs(want.item()), (name, step)
            for a, b in zip(desc.tables, ref):
                np.testing.assert_allclose(a.cpu().numpy(), b.detach().numpy(), rtol=0, atol=3e-5)
            assert all(float(s.abs().max()) == 0.0 for s in scratch), "gradient scratch must be left zeroed"


@pytest.mark.parametrize("neg_rate,B", [(1, 70), (3, 129), (4, 64), (5, 33), (16, 64), (256, 40)])
def test_fused_selfadv_step_equals_the_five_launch_path(neg_rate, B):
    """kge_train_pairwise_selfadv (RotatE: forward + self-adversarial loss + backward in one kernel; a warp per
    positive up to neg_rate 4, a CTA per positive beyond) == kge_score_fwd x2 + kge_loss_selfadv + kge_score_bwd x2:
    the loss terms are the same bits (summed by unordered atomics -> compared to fp32 rounding), the gradients
    equal up to the order of the float atomics."""
    L = _L()
    N, R, d = 700, 9, 100
    om, tabs = gpu.synthetic_case("rotate", N, R, d, seed=31)
    desc = gpu.desc_from_oracle_model(om)
    rng = np.random.RandomState(neg_rate)
    ids = [_cuda(rng.randint(N if k % 3 != 1 else R, size=B if k < 3 else B * neg_rate)) for k in range(6)]
    g_fused = [torch.zeros_like(x) for x in desc.tables]
    loss_fused = L.train_pairwise_selfadv(desc, g_fused, *ids, neg_rate=neg_rate, alpha=0.5)
    pos, neg = L.score_fwd(desc, *ids[:3]), L.score_fwd(desc, *ids[3:])
    loss, gp, gn = L.loss_selfadv(pos, neg, neg_rate, 0.5)
    g_ref = [torch.zeros_like(x) for x in desc.tables]
    L.score_bwd(desc, *ids[:3], gp, g_ref)
    L.score_bwd(desc, *ids[3:], gn, g_ref)
    assert abs(loss_fused.item() - loss.item()) <= 2e-6 * abs(loss.item())
    _check_grads([g.cpu().numpy() for g in g_fused], [g.cpu().numpy() for g in g_ref], tol=2e-5, what="selfadv fused")
    with pytest.raises(L.KgeNotSupported):   # other models do not train with this loss
        om2, _ = gpu.synthetic_case("transe", N, R, d, seed=1)
        d2 = gpu.desc_from_oracle_model(om2)
        L.train_pairwise_selfadv(d2, [torch.zeros_like(x) for x in d2.tables], *ids, neg_rate=neg_rate, alpha=0.5)


def _trainer_for(model_name, kg, **cfgkw):
    import pykg2vec_b200
    from pykg2vec_b200.synthetic import SyntheticConfig
    from pykg2vec_b200.trainer import Trainer
    cfg = SyntheticConfig(kg, **cfgkw)
    torch.manual_seed(0)
    model = pykg2vec_b200.import_model(model_name)(**cfg.__dict__)
    tr = Trainer(model, cfg)
    tr.build_model()
    return tr


@pytest.mark.parametrize("model_name,opt", [
    ("transe", "sgd"), ("transe", "adagrad"), ("distmult", "sgd"), ("complex", "adagrad"), ("rotate", "adagrad"),
    # every pointwise model with its own get_reg default (SimplE's id-tensor regulariser, QuatE /
    # OctonionE |x|^3, CP signed x^3, ComplexN3 |x|^3), and Rescal whose forward() normalises in place
    ("cp", "sgd"), ("complexn3", "sgd"), ("analogy", "adagrad"), ("simple", "sgd"), ("simple_ignr", "adagrad"),
    ("quate", "sgd"), ("octonione", "adagrad"), ("rescal", "sgd"), ("rescal", "adagrad"), ("hole", "sgd"),
    ("transh", "sgd"), ("transd", "adagrad")])
def test_trainer_fused_equals_autograd_mode(model_name, opt):
    """Trainer.train_batch in fused mode follows the same weight trajectory as the autograd
    mode (reference step order) with the dense torch optimizer."""
    from pykg2vec_b200.synthetic import SyntheticKnowledgeGraph
    kg = SyntheticKnowledgeGraph(400, 6, 2000, 50, 50, seed=1)
    kw = dict(optimizer=opt, learning_rate=0.05, hidden_size=32 if model_name == "rescal" else 64,
              ent_hidden_size=64, rel_hidden_size=64, margin=1.0 if model_name != "rotate" else 6.0,
              l1_flag=False, lmbda=0.01, neg_rate=4 if model_name == "rotate" else 1, alpha=0.5,
              cmax=0.5, cmin=-0.5, batch_size=256)
    a = _trainer_for(model_name, kg, fused_step=True, **kw)
    b = _trainer_for(model_name, kg, fused_step=False, **kw)
    b.model.load_state_dict(a.model.state_dict())
    assert a._fused and not b._fused
    rng = np.random.RandomState(2)
    B = 256
    for step in range(3):
        if a.model.training_strategy.name == "PAIRWISE_BASED":
            nr = kw["neg_rate"]
            data = [rng.randint(400, size=B), rng.randint(6, size=B), rng.randint(400, size=B),
                    rng.randint(400, size=B * nr), rng.randint(6, size=B * nr), rng.randint(400, size=B * nr)]
        else:
            data = [rng.randint(400, size=B), rng.randint(6, size=B), rng.randint(400, size=B),
                    np.where(np.arange(B) % 2 == 0, 1, -1)]
        la, lb = a.train_batch(data), b.train_batch(data)
        assert abs(la - lb) <= 1e-4 * max(abs(lb), 1e-6), (step, la, lb)
        for (ka, va), (kb, vb) in zip(a.model.state_dict().items(), b.model.state_dict().items()):
            np.testing.assert_allclose(va.cpu().numpy(), vb.cpu().numpy(), rtol=0, atol=2e-5, err_msg=ka)


def test_evaluator_matches_oracle_and_reference_metrics():
    """Evaluator.test (batched rank kernel) == oracle rank counts; settle() metrics follow."""
    import oracle
    from oracle import ref_port
    from pykg2vec_b200.synthetic import SyntheticKnowledgeGraph
    kg = SyntheticKnowledgeGraph(600, 5, 3000, 40, 30, seed=5)
    tr = _trainer_for("complex", kg, hidden_size=32, lmbda=0.1)
    ev = tr.evaluator
    res = ev.full_test(epoch=0)
    tabs = [w.detach().cpu().numpy() for w in tr.model.kge_tables()]
    om = oracle.Model("complex", tabs, 32)
    test = kg.arrays["test"]
    from pykg2vec_b200.evaluator import build_filter_csr
    hr_t, tr_h = kg.read_cache_data("hr_t"), kg.read_cache_data("tr_h")
    ft = build_filter_csr([(int(h), int(r)) for h, r, t in test], hr_t)
    fh = build_filter_csr([(int(t), int(r)) for h, r, t in test], tr_h)
    want = oracle.rank_1vsall(om, test[:, 0], test[:, 1], test[:, 2], ft, fh)
    mc = ev.metric_calculator
    got = np.stack([mc.rank_tail, mc.f_rank_tail, mc.rank_head, mc.f_rank_head], axis=1)
    np.testing.assert_array_equal(got, want)
    ref = ref_port.settle(want)
    assert res["mr"] == pytest.approx(ref["mr"]) and res["fmrr"] == pytest.approx(ref["fmrr"])
    # infer-style single query API: descending score order, length topk (evaluator.py:249-260)
    top = ev.test_tail_rank(int(test[0, 0]), int(test[0, 1]), topk=5)
    assert top.shape == (5,)
    full = oracle.sweep_scores(om, oracle.GROUP_TAIL, int(test[0, 0]), int(test[0, 1]), 0)
    assert set(top.cpu().tolist()) == set(np.argsort(-full, kind="stable")[:5].tolist())


def test_rescal_model_normalises_in_place_like_reference():
    """Rescal.forward mutates its tables (pairwise.py:843-844): rows become unit-norm, then scores
    match the oracle on the normalised tables; backward reaches both tables."""
    import oracle
    import pykg2vec_b200
    torch.manual_seed(3)
    m = pykg2vec_b200.import_model("rescal")(tot_entity=90, tot_relation=4, hidden_size=20, margin=1.0).cuda()
    before = [w.detach().cpu().numpy().copy() for w in m.kge_tables()]
    rng = np.random.RandomState(0)
    h, r, t = rng.randint(90, size=40), rng.randint(4, size=40), rng.randint(90, size=40)
    s = m(_cuda(h), _cuda(r), _cuda(t))
    want_tabs = [oracle.normalize_rows(b.copy()) for b in before]
    for w, wt in zip(m.kge_tables(), want_tabs):
        np.testing.assert_array_equal(w.detach().cpu().numpy(), wt)
    om = oracle.Model("rescal", want_tabs, 20)
    np.testing.assert_array_equal(gpu.bits(s.detach().cpu().numpy()), gpu.bits(oracle.score_fwd(om, h, r, t)))
    s.sum().backward()
    assert all(w.grad is not None and float(w.grad.abs().sum()) > 0 for w in m.kge_tables())


def test_device_sampler_matches_oracle_and_rules():
    """kge_sample_negatives == oracle bit-for-bit; negatives never hit a positive; layout and
    labels follow generator.py:42-158; Bernoulli probabilities steer the corrupted side."""
    import oracle
    L = _L()
    rng = np.random.RandomState(0)
    N, R, n_train = 60, 4, 2500  # dense graph: rejections do happen
    train = np.unique(np.stack([rng.randint(N, size=n_train), rng.randint(R, size=n_train),
                                rng.randint(N, size=n_train)], 1), axis=0)
    slots = L.tripleset_build(_cuda(train[:, 0].copy()), _cuda(train[:, 1].copy()), _cuda(train[:, 2].copy()), N, R)
    B = 300
    sel = rng.randint(len(train), size=B)
    ph, pr, pt = train[sel, 0].copy(), train[sel, 1].copy(), train[sel, 2].copy()
    positives = set(map(tuple, train.tolist()))
    for neg_rate, probs in ((1, None), (4, np.array([0.05, 0.5, 0.95, 0.3], dtype=np.float32))):
        hp = _cuda(probs) if probs is not None else None
        nh, nr, nt = L.sample_negatives(slots, _cuda(ph), _cuda(pr), _cuda(pt), neg_rate, hp, N, seed=7, step=3)
        want = oracle.sample_negatives(train, ph, pr, pt, neg_rate, probs, N, 7, 3)
        for a, b in zip((nh, nr, nt), want):
            np.testing.assert_array_equal(a.cpu().numpy(), b)
        nh, nr, nt = nh.cpu().numpy(), nr.cpu().numpy(), nt.cpu().numpy()
        rep_h, rep_t = np.repeat(ph, neg_rate), np.repeat(pt, neg_rate)
        assert np.array_equal(nr, np.repeat(pr, neg_rate))
        assert all((nh[i] == rep_h[i]) or (nt[i] == rep_t[i]) for i in range(len(nh)))  # one side kept
        assert not any((int(a), int(b), int(c)) in positives for a, b, c in zip(nh, nr, nt))
        if probs is not None:
            head_corrupted = nt == rep_t
            r_rep = np.repeat(pr, neg_rate)
            assert head_corrupted[r_rep == 2].mean() > 0.8 and head_corrupted[r_rep == 0].mean() < 0.2
        # pointwise layout: each positive followed by its negatives, labels +1/-1
        h4, r4, t4, y4 = L.sample_negatives(slots, _cuda(ph), _cuda(pr), _cuda(pt), neg_rate, hp, N, seed=7, step=3, layout=1)
        w4 = oracle.sample_negatives(train, ph, pr, pt, neg_rate, probs, N, 7, 3, layout=1)
        for a, b in zip((h4, r4, t4, y4), w4):
            np.testing.assert_array_equal(a.cpu().numpy(), b)
        y = y4.cpu().numpy().reshape(B, 1 + neg_rate)
        assert (y[:, 0] == 1).all() and (y[:, 1:] == -1).all()
        assert np.array_equal(h4.cpu().numpy().reshape(B, -1)[:, 0], ph)
    # a different step gives a different draw
    a = L.sample_negatives(slots, _cuda(ph), _cuda(pr), _cuda(pt), 1, None, N, seed=7, step=4)
    b = L.sample_negatives(slots, _cuda(ph), _cuda(pr), _cuda(pt), 1, None, N, seed=7, step=3)
    assert not (torch.equal(a[0], b[0]) and torch.equal(a[2], b[2]))


def test_generator_feeds_trainer_epoch():
    from pykg2vec_b200.generator import Generator, relation_property
    from pykg2vec_b200.synthetic import SyntheticKnowledgeGraph
    kg = SyntheticKnowledgeGraph(300, 5, 4000, 50, 50, seed=3)
    for model_name, opt in (("transe", "sgd"), ("distmult", "adagrad")):
        tr = _trainer_for(model_name, kg, optimizer=opt, learning_rate=0.05, hidden_size=32, margin=1.0,
                          l1_flag=False, lmbda=0.01, batch_size=256, neg_rate=2 if model_name == "distmult" else 1,
                          sampling="bern")
        gen = Generator(tr.model, tr.config, seed=1)
        assert gen.head_prob is not None and gen.head_prob.shape[0] == 5
        before = [w.detach().clone() for w in tr.model.kge_tables()]
        loss = tr.train_model_epoch(gen, num_batch=5)
        assert np.isfinite(loss) and loss > 0
        assert any(not torch.equal(a, b.detach()) for a, b in zip(before, tr.model.kge_tables()))
        gen.start_one_epoch(1)
        batch = next(gen)
        assert len(batch) == (6 if model_name == "transe" else 4) and all(x.is_cuda for x in batch)
        with pytest.raises(StopIteration):
            next(gen)
    p = relation_property(kg.arrays["train"], 5)
    assert p.shape == (5,) and ((p > 0) & (p < 1)).all()


def test_convkb_trains_like_the_written_chain():
    """ConvKB (kernel on the collapsed affine form) follows the as-written conv -> concat -> Linear
    chain through whole optimizer steps: same loss and same updated parameters as torch autograd
    on oracle.ref_port's restatement, with the convolution filters left untouched (they are not
    registered parameters in the reference either, pointwise.py:280)."""
    import pykg2vec_b200
    from oracle import ref_port
    from pykg2vec_b200.criterion import Criterion
    g = gu.load("convkb_d24")
    m = _make_model(g)
    raw = [torch.from_numpy(x.astype(np.float64)) for x in gu.raw_tables_of(g)]
    train_idx = [0, 1, len(raw) - 2, len(raw) - 1]          # ent, rel, fc1.weight, fc1.bias
    params = [raw[i].clone().requires_grad_() for i in train_idx]
    opt_ref = torch.optim.SGD(params, lr=0.05)
    opt = torch.optim.SGD(m.parameters(), lr=0.05)
    assert len(list(m.parameters())) == 4
    conv_before = [c.weight.detach().clone() for c in m.conv_list]
    h, r, t = g["h"], g["r"], g["t"]
    y = np.where(np.arange(len(h)) % 2 == 0, 1.0, -1.0).astype(np.float32)
    for step in range(3):
        tabs = list(raw)
        for i, p in zip(train_idx, params):
            tabs[i] = p
        s_ref = ref_port.score("convkb_raw", tabs, torch.from_numpy(h), torch.from_numpy(r), torch.from_numpy(t))
        loss_ref = ref_port.pointwise_logistic(s_ref, torch.from_numpy(y).double())
        opt_ref.zero_grad(); loss_ref.backward(); opt_ref.step()
        loss = Criterion.pointwise_logistic(m(_cuda(h), _cuda(r), _cuda(t)), _cuda(y)) + m.get_reg(None, None, None)
        opt.zero_grad(); loss.backward(); opt.step()
        assert abs(loss.item() - loss_ref.item()) <= 2e-5 * abs(loss_ref.item())
    got = [m.ent_embeddings.weight, m.rel_embeddings.weight, m.fc1.weight, m.fc1.bias]
    for a, b in zip(got, params):
        np.testing.assert_allclose(a.detach().cpu().numpy(), b.detach().numpy(), rtol=2e-4, atol=2e-6)
    for c, w0 in zip(m.conv_list, conv_before):
        assert torch.equal(c.weight.detach(), w0)


def test_train_batch_async_handle_equals_sync():
    """train_batch(sync=False) enqueues the same graph-staged step and hands back the loss lazily:
    identical losses and weights to the synchronous calls."""
    from pykg2vec_b200.synthetic import SyntheticKnowledgeGraph
    kg = SyntheticKnowledgeGraph(400, 6, 2000, 50, 50, seed=1)
    kw = dict(optimizer="sgd", learning_rate=0.05, hidden_size=64, margin=1.0, l1_flag=False, neg_rate=1)
    a = _trainer_for("transe", kg, fused_step=True, **kw)
    b = _trainer_for("transe", kg, fused_step=True, **kw)
    b.model.load_state_dict(a.model.state_dict())
    rng = np.random.RandomState(4)
    B = 128
    batches = [[rng.randint(400, size=B), rng.randint(6, size=B), rng.randint(400, size=B),
                rng.randint(400, size=B), rng.randint(6, size=B), rng.randint(400, size=B)] for _ in range(4)]
    want = [a.train_batch(d) for d in batches]
    got = []
    for d in batches:
        h = b.train_batch(d, sync=False)      # enqueued only
        _ = sum(int(x.sum()) for x in d)      # host work overlapping the step
        got.append(float(h))                  # waits for this step's D2H (the handle is valid until the next call)
    # (the loss and the row gradients are accumulated with float atomics: equal up to summation order)
    np.testing.assert_allclose(got, want, rtol=1e-5)
    for (ka, va), (kb, vb) in zip(a.model.state_dict().items(), b.model.state_dict().items()):
        np.testing.assert_allclose(va.cpu().numpy(), vb.cpu().numpy(), rtol=0, atol=2e-6, err_msg=ka)
