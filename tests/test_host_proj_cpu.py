"""The projection path's PYTHON GLUE on the CPU (no GPU): pykg2vec_b200.projection / functional /
criterion / trainer / evaluator / generator run unmodified, with the ctypes entry points of `_lib`
replaced by the SAME kernels executed under the host emulation (tests/emu/, through the product's
launch plans).  Checked against the golden vectors the reference produced: ConvE and TuckER
predictions, Evaluator ranks, the multi-class BCE training step (loss + gradient of every parameter),
and a Generator -> Trainer epoch with label rows built by the emulated kernel.

This is a test double for the DEVICE only — the product still refuses CPU tensors (see
test_oracle_proj.py::test_conve_mirror_contract_errors); here that guard is lifted explicitly."""
import ctypes
import types

import numpy as np
import pytest
import torch

import golden_util as gu
from test_emu_proj import emu  # noqa: F401  (fixture: the emulated projection kernels)


def _p(t):
    return ctypes.c_void_p(t.data_ptr()) if t is not None else None


@pytest.fixture
def cpu_device_double(emu, monkeypatch):  # noqa: F811
    from pykg2vec_b200 import _lib, functional, projection

    def proj_tail_fwd(x, ent, bias=None, out=None):
        B, k = x.shape
        N = ent.shape[0]
        out = torch.empty((B, N), dtype=torch.float32) if out is None else out
        b = bias.contiguous().view(-1) if bias is not None else None
        emu.emu_proj_tail_fwd(_p(x.contiguous()), _p(ent), _p(b), ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_int32(k),
                              _p(out), ctypes.c_int32(B % 3))          # rotate through the CTA tiles
        return out

    def proj_tail_bwd(gp, preds, x, ent, gx=None, ge=None, gb=None):
        B, k = x.shape
        emu.emu_proj_tail_bwd(_p(gp), _p(preds), _p(x), _p(ent), ctypes.c_int64(B), ctypes.c_int64(ent.shape[0]),
                              ctypes.c_int32(k), _p(gx), _p(ge), _p(gb), ctypes.c_int32(24))

    def proj_bce(preds, labels, scale=1.0, shift=0.0, gscale=1.0, want_grad=True):
        B, N = preds.shape
        loss = torch.zeros(1)
        g = torch.empty_like(preds) if want_grad else None
        emu.emu_proj_bce(_p(preds), _p(labels), ctypes.c_int64(B), ctypes.c_int64(N), ctypes.c_float(scale),
                         ctypes.c_float(shift), ctypes.c_float(gscale), _p(loss), _p(g), ctypes.c_int32(4))
        return loss, g

    def proj_rank(x, ent, bias, tgt, filt=None, direction=0, counts=None, workspace=None):
        Q, k = x.shape
        counts = torch.zeros((Q, 4), dtype=torch.int32) if counts is None else counts
        thr = torch.zeros(Q)
        fp, fi = filt if filt is not None else (None, None)
        emu.emu_proj_rank(_p(x), _p(ent), _p(bias), ctypes.c_int64(Q), ctypes.c_int64(ent.shape[0]), ctypes.c_int32(k),
                          _p(tgt), _p(fp), _p(fi), ctypes.c_int64(fi.numel() if fi is not None else 0),
                          ctypes.c_int32(direction), _p(counts), _p(thr), ctypes.c_int32(Q % 3))
        return counts

    def conve_trunk_fwd(model, e, r, out=None):
        import oracle
        sd = {k: v.detach().contiguous() for k, v in model.state_dict().items()}
        p = oracle.KgeConve()
        p.hidden_size, p.hidden_size_1 = int(model.hidden_size), int(model.hidden_size_1)
        p.bn0_eps, p.bn1_eps = float(model.bn0.eps), float(model.bn1.eps)
        for f, key in oracle.CONVE_KEYS.items():
            setattr(p, f, sd[key].data_ptr())
        Q, k, k1 = e.numel(), p.hidden_size, p.hidden_size_1
        F = 32 * (2 * (k // k1) - 2) * (k1 - 2)
        x = torch.empty((Q, k))
        ws = torch.empty(Q * F + ((F + 511) // 512) * Q * k)
        emu.emu_conve_trunk_fwd(ctypes.byref(p), _p(e), _p(r), ctypes.c_int64(Q), _p(x), _p(ws))
        return x

    def proj_labels(rows, ptr, idx, B, N, out=None):
        out = torch.empty((B, N), dtype=torch.float32) if out is None else out
        emu.emu_proj_labels(_p(rows), _p(ptr), _p(idx), ctypes.c_int64(B), ctypes.c_int64(N), _p(out))
        return out

    for name, fn in (("proj_tail_fwd", proj_tail_fwd), ("proj_tail_bwd", proj_tail_bwd), ("proj_bce", proj_bce),
                     ("proj_rank", proj_rank), ("conve_trunk_fwd", conve_trunk_fwd), ("proj_labels", proj_labels)):
        monkeypatch.setattr(_lib, name, fn)
    for mod in (functional, projection):
        monkeypatch.setattr(mod, "_require_cuda", lambda *a: None)
    monkeypatch.setattr(torch.Tensor, "pin_memory", lambda self: self)
    return _lib


def _conve(g, train=False):
    from pykg2vec_b200 import import_model
    m = import_model("conve")(tot_entity=int(g["N"]), tot_relation=int(g["R"]), hidden_size=int(g["hidden_size"]),
                              hidden_size_1=int(g["hidden_size_1"]), lmbda=0.1, input_dropout=0.0,
                              feature_map_dropout=0.0, hidden_dropout=0.0)
    m.load_state_dict({k_: torch.from_numpy(np.asarray(v)) for k_, v in gu.proj_state(g).items()}, strict=True)
    return m.train(train)


def _evaluator(model, N):
    from pykg2vec_b200.evaluator import Evaluator
    ev = object.__new__(Evaluator)
    ev.model, ev.config = model, types.SimpleNamespace(device="cpu", tot_entity=N)
    ev._filter_cache, ev._workspace = {}, None
    return ev


@pytest.mark.parametrize("name", gu.proj_case_names())
def test_conve_glue_eval(cpu_device_double, name):
    g = gu.load(name)
    m = _conve(g)
    h, r, t = (torch.from_numpy(g[k]) for k in ("h", "r", "t"))
    with torch.no_grad():
        pt = m(h, r, direction="tail").numpy()          # native trunk (emulated) + tail kernel (emulated)
        ph = m(t, r, direction="head").numpy()
    assert gu.rel_err(pt, g["preds_tail"]).max() < 1e-4 and gu.rel_err(ph, g["preds_head"]).max() < 1e-4
    pt2 = m(h, r, direction="tail").detach().numpy()    # grad-enabled: torch layers + tail kernel
    assert gu.rel_err(pt2, g["preds_tail"]).max() < 1e-4
    Q = g["ranks"].shape[0]
    with torch.no_grad():
        got = _evaluator(m, int(g["N"])).rank_triples(g["h"][:Q], g["r"][:Q], g["t"][:Q],
                                                      (g["filt_t_ptr"], g["filt_t_idx"]),
                                                      (g["filt_h_ptr"], g["filt_h_idx"]))
    assert np.array_equal(got, g["ranks"])


@pytest.mark.parametrize("name", gu.proj_case_names())
def test_conve_glue_training_step(cpu_device_double, name):
    from pykg2vec_b200.trainer import Trainer
    g = gu.load(name)
    m = _conve(g, train=True)
    tr = object.__new__(Trainer)
    tr.model = m
    tr.config = types.SimpleNamespace(device="cpu", label_smoothing=float(g["label_smoothing"]), tot_entity=int(g["N"]))
    m.zero_grad()
    loss = tr.train_step_projection(*(torch.from_numpy(g[k]) for k in ("h", "r", "t", "tr_labels_tail", "tr_labels_head")))
    loss.backward()
    assert abs(loss.item() - float(g["tr_loss"])) <= 1e-5 * abs(float(g["tr_loss"]))
    for key, p in m.named_parameters():
        got = p.grad.numpy()
        if "grad_" + key in g:
            want = g["grad_" + key]
        else:
            got, want = got.reshape(-1)[::37], g["gradsample_" + key]
        assert np.abs(got - want).max() <= 5e-4 * np.abs(want).max() + 1e-7, key
    for key, v in m.state_dict().items():
        if "running" in key:
            assert np.abs(v.numpy() - g["sd_after_" + key]).max() < 1e-5, key


def test_tucker_glue(cpu_device_double):
    from pykg2vec_b200 import import_model
    from pykg2vec_b200.trainer import Trainer
    g = gu.load("tucker_d32")
    m = import_model("tucker")(tot_entity=int(g["N"]), tot_relation=int(g["R"]), ent_hidden_size=32, rel_hidden_size=16,
                               lmbda=0.1, input_dropout=0.0, hidden_dropout1=0.0, hidden_dropout2=0.0)
    m.load_state_dict({k_: torch.from_numpy(np.asarray(v)) for k_, v in gu.proj_state(g).items()}, strict=True)
    h, r, t = (torch.from_numpy(g[k]) for k in ("h", "r", "t"))
    m.eval()
    with torch.no_grad():
        assert gu.rel_err(m(h, r, direction="tail").numpy(), g["preds_tail"]).max() < 1e-4
        Q = g["ranks"].shape[0]
        got = _evaluator(m, int(g["N"])).rank_triples(g["h"][:Q], g["r"][:Q], g["t"][:Q],
                                                      (g["filt_t_ptr"], g["filt_t_idx"]),
                                                      (g["filt_h_ptr"], g["filt_h_idx"]))
    assert np.array_equal(got, g["ranks"])
    m.train()
    tr = object.__new__(Trainer)
    tr.model = m
    tr.config = types.SimpleNamespace(device="cpu", label_smoothing=float(g["label_smoothing"]), tot_entity=int(g["N"]))
    m.zero_grad()
    loss = tr.train_step_projection(h, r, t, torch.from_numpy(g["tr_labels_tail"]), torch.from_numpy(g["tr_labels_head"]))
    loss.backward()
    assert abs(loss.item() - float(g["tr_loss"])) <= 1e-5 * abs(float(g["tr_loss"]))
    for key, p in m.named_parameters():
        want = g["grad_" + key]
        assert np.abs(p.grad.numpy() - want).max() <= 5e-4 * np.abs(want).max() + 1e-7, key


def test_generator_trainer_epoch_glue(cpu_device_double):
    from pykg2vec_b200 import import_model
    from pykg2vec_b200.generator import Generator
    from pykg2vec_b200.synthetic import SyntheticConfig, SyntheticKnowledgeGraph
    from pykg2vec_b200.trainer import Trainer
    kg = SyntheticKnowledgeGraph(60, 3, 400, 20, 20, seed=3)
    cfg = SyntheticConfig(kg, device="cpu", optimizer="adam", learning_rate=0.01, batch_size=16, neg_rate=0,
                          hidden_size=48, hidden_size_1=8, lmbda=0.1, input_dropout=0.0, feature_map_dropout=0.0,
                          hidden_dropout=0.0, label_smoothing=0.1, test_num=10)
    torch.manual_seed(0)
    model = import_model("conve")(**cfg.__dict__)
    tr = Trainer(model, cfg)
    tr.build_model()
    gen = Generator(model, cfg, seed=1)
    gen.start_one_epoch(1)
    h, r, t, hr_t, tr_h = next(gen)
    known = kg.read_cache_data("hr_t_train")
    for i in range(16):
        assert set(np.nonzero(hr_t[i].numpy())[0].tolist()) == known[(int(h[i]), int(r[i]))]
    losses = [tr.train_model_epoch(gen, num_batch=4) for _ in range(3)]
    assert all(np.isfinite(losses)) and losses[-1] < losses[0]
    model.eval()
    scores = tr.evaluator.mini_test(epoch=0)
    assert 1.0 <= scores["fmr"] <= scores["mr"] <= 60.0
