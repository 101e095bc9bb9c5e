"""CPU tests (-m "not gpu"): the oracle's projection-tail / ConvE-trunk restatement against the
golden vectors produced by the reference itself (tests/golden/make_golden_proj.py), and the host
surface of the ConvE mirror."""
import numpy as np
import pytest
import torch

import golden_util as gu
import oracle

CASES = gu.proj_case_names()


def _scalars(g):
    N = int(g["N"])
    ls = float(g["label_smoothing"])
    return np.float32(1.0 - ls), np.float32(1.0 / N)


def test_cases_present():
    assert CASES == ["conve_d100", "conve_d48"]


@pytest.mark.parametrize("name", CASES)
def test_tail_forward_matches_reference(name):
    """sigmoid(x.E^T + b) on the reference's own x: <= 1e-4 relative (north_star tolerance), in fact ~1e-7"""
    g = gu.load(name)
    E, b = g["sd_ent_embeddings.weight"], g["sd_b.weight"]
    for tag in ("tail", "head"):
        p = oracle.proj_tail_fwd(g["x_" + tag], E, b)
        assert gu.rel_err(p, g["preds_" + tag]).max() < 1e-6


@pytest.mark.parametrize("name", CASES)
def test_trunk_matches_reference(name):
    g = gu.load(name)
    st = gu.proj_state(g)
    k, k1, R = int(g["hidden_size"]), int(g["hidden_size_1"]), int(g["R"])
    xt = oracle.conve_trunk_fwd(st, k, k1, g["h"], g["r"])
    xh = oracle.conve_trunk_fwd(st, k, k1, g["t"], g["r"] + R)
    assert np.abs(xt - g["x_tail"]).max() < 1e-5 and np.abs(xh - g["x_head"]).max() < 1e-5
    # end to end: trunk + tail within the 1e-4 relative bar on the predictions
    E, b = st["ent_embeddings.weight"], st["b.weight"]
    assert gu.rel_err(oracle.proj_tail_fwd(xt, E, b), g["preds_tail"]).max() < 1e-4
    assert gu.rel_err(oracle.proj_tail_fwd(xh, E, b), g["preds_head"]).max() < 1e-4


@pytest.mark.parametrize("name", CASES)
@pytest.mark.parametrize("own_trunk", [False, True])
def test_ranks_match_reference_evaluator(name, own_trunk):
    """count formulation == the reference's predict_*_rank + MetricCalculator walk, exactly"""
    g = gu.load(name)
    st = gu.proj_state(g)
    k, k1, R = int(g["hidden_size"]), int(g["hidden_size_1"]), int(g["R"])
    Q = g["ranks"].shape[0]
    h, r, t = g["h"][:Q], g["r"][:Q], g["t"][:Q]
    if own_trunk:
        xt, xh = oracle.conve_trunk_fwd(st, k, k1, h, r), oracle.conve_trunk_fwd(st, k, k1, t, r + R)
    else:
        xt, xh = g["x_tail"][:Q], g["x_head"][:Q]
    E, b = st["ent_embeddings.weight"], st["b.weight"]
    c = np.zeros((Q, 4), dtype=np.int32)
    oracle.proj_rank(xt, E, b, t, (g["filt_t_ptr"], g["filt_t_idx"]), 0, c)
    oracle.proj_rank(xh, E, b, h, (g["filt_h_ptr"], g["filt_h_idx"]), 1, c)
    assert np.array_equal(c, g["ranks"])


@pytest.mark.parametrize("name", CASES)
def test_bce_and_tail_backward_match_reference_autograd(name):
    g = gu.load(name)
    E = g["sd_ent_embeddings.weight"]
    scale, shift = _scalars(g)
    total = 0.0
    for tag in ("tail", "head"):
        loss, gp = oracle.proj_bce(g["tr_preds_" + tag], g["tr_labels_" + tag], scale, shift, 1.0)
        assert abs(loss - float(g["tr_loss_" + tag])) <= 2e-6 * abs(loss)
        total += loss
        gx, ge, gb = oracle.proj_tail_bwd(gp, g["tr_preds_" + tag], g["tr_x_" + tag], E)
        for got, key in ((gx, "tr_gx_"), (ge, "tr_gE_"), (gb, "tr_gb_")):
            want = g[key + tag].reshape(got.shape)
            assert np.abs(got - want).max() <= 2e-6 * np.abs(want).max()
    assert abs(total - float(g["tr_loss"])) <= 2e-6 * abs(total)


def test_bce_without_smoothing_matches_torch():
    rng = np.random.RandomState(0)
    p = (1 / (1 + np.exp(-rng.standard_normal((5, 40)) * 4))).astype(np.float32)
    p[0, :3] = [0.0, 1.0, 0.5]
    y = (rng.rand(5, 40) < 0.2).astype(np.float32)
    pt = torch.from_numpy(p).requires_grad_()
    want = torch.mean(torch.nn.BCEWithLogitsLoss()(pt, torch.from_numpy(y)))
    want.backward()
    loss, g = oracle.proj_bce(p, y, 1.0, 0.0, 1.0)
    assert abs(loss - want.item()) < 1e-6
    assert np.abs(g - pt.grad.numpy()).max() < 1e-8


# ---- host surface of the mirror (no GPU, no compute) ---------------------------------------------
@pytest.mark.parametrize("name", CASES)
def test_conve_mirror_state_dict_is_the_reference_layout(name):
    from pykg2vec_b200 import import_model
    g = gu.load(name)
    st = gu.proj_state(g)
    m = import_model("conve")(tot_entity=int(g["N"]), tot_relation=int(g["R"]), hidden_size=int(g["hidden_size"]),
                              hidden_size_1=int(g["hidden_size_1"]), lmbda=0.1, input_dropout=0.0,
                              feature_map_dropout=0.0, hidden_dropout=0.0)
    own = m.state_dict()
    assert sorted(own) == sorted(st)
    for k_, v in st.items():
        assert tuple(own[k_].shape) == tuple(v.shape), k_
    m.load_state_dict({k_: torch.from_numpy(np.asarray(v)) for k_, v in st.items()}, strict=True)
    assert [p.name for p in m.parameter_list] == ["ent_embedding", "rel_embedding", "b"]
    assert m.model_name == "conve" and m.training_strategy.name == "PROJECTION_BASED"


def test_conve_mirror_contract_errors():
    from pykg2vec_b200 import import_model, _lib
    C = import_model("conve")
    with pytest.raises(Exception, match="hyperparameter hidden_size_1 not found!"):
        C(tot_entity=10, tot_relation=2, hidden_size=20, lmbda=0.1, input_dropout=0.0, feature_map_dropout=0.0,
          hidden_dropout=0.0)
    m = C(tot_entity=10, tot_relation=2, hidden_size=20, hidden_size_1=5, lmbda=0.1, input_dropout=0.0,
          feature_map_dropout=0.0, hidden_dropout=0.0)
    e, r = torch.tensor([1, 2]), torch.tensor([0, 1])
    with pytest.raises(AssertionError, match="Unknown forward direction"):
        m(e, r, direction="sideways")
    with pytest.raises(_lib.KgeError, match="no CPU"):      # no CPU fallback for the product path
        m(e, r, direction="tail")


def test_label_csr_host_logic():
    """Generator's CSR of hr_t_train / tr_h_train over distinct keys + row of every training triple"""
    from pykg2vec_b200.generator import _label_csr
    from pykg2vec_b200.synthetic import SyntheticKnowledgeGraph
    kg = SyntheticKnowledgeGraph(50, 3, 400, 10, 10, seed=1)
    arr = kg.arrays["train"]
    for key, a, b in (("hr_t_train", 0, 1), ("tr_h_train", 2, 1)):
        known = kg.read_cache_data(key)
        rows, ptr, idx = (x.numpy() for x in _label_csr(known, arr[:, a], arr[:, b], "cpu"))
        assert len(ptr) - 1 == len(known) and ptr[-1] == len(idx) == sum(len(v) for v in known.values())
        for i in range(len(arr)):
            assert set(idx[ptr[rows[i]]:ptr[rows[i] + 1]].tolist()) == known[(int(arr[i, a]), int(arr[i, b]))]


def test_tucker_tail_and_ranks_match_reference():
    """TuckER's tail has no bias row (projection.py:335-336): oracle vs the reference's outputs"""
    g = gu.load("tucker_d32")
    E = g["sd_ent_embeddings.weight"]
    for tag in ("tail", "head"):
        assert gu.rel_err(oracle.proj_tail_fwd(g["x_" + tag], E, None), g["preds_" + tag]).max() < 1e-6
    Q = g["ranks"].shape[0]
    c = np.zeros((Q, 4), dtype=np.int32)
    oracle.proj_rank(g["x_tail"][:Q], E, None, g["t"][:Q], (g["filt_t_ptr"], g["filt_t_idx"]), 0, c)
    oracle.proj_rank(g["x_head"][:Q], E, None, g["h"][:Q], (g["filt_h_ptr"], g["filt_h_idx"]), 1, c)
    assert np.array_equal(c, g["ranks"])


def test_tucker_mirror_surface():
    from pykg2vec_b200 import import_model
    g = gu.load("tucker_d32")
    st = gu.proj_state(g)
    m = import_model("tucker")(tot_entity=int(g["N"]), tot_relation=int(g["R"]), ent_hidden_size=32,
                               rel_hidden_size=16, lmbda=0.1, input_dropout=0.0, hidden_dropout1=0.0,
                               hidden_dropout2=0.0)
    assert sorted(m.state_dict()) == sorted(st) == ["W.weight", "ent_embeddings.weight", "rel_embeddings.weight"]
    m.load_state_dict({k_: torch.from_numpy(np.asarray(v)) for k_, v in st.items()}, strict=True)
    assert [p.name for p in m.parameter_list] == ["ent_embedding", "rel_embedding", "W"]
    assert m.training_strategy.name == "PROJECTION_BASED" and m.model_name == "tucker"
