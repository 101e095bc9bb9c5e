"""Pairwise (margin-based) models — same constructor kwargs, attribute names,
state_dict keys, parameter_list and loss bindings as pykg2vec/models/pairwise.py; the
batch score forward() is the fused gather+score CUDA kernel (kge_score_fwd), its
backward the scatter kernel (kge_score_bwd).

embed() is not on the hot path (it serves export / visualisation,
pykg2vec/utils/trainer.py:462-471, visualization.py:91) and stays plain tensor code.
"""
import math

import numpy as np
import torch
import torch.nn as nn
import torch.nn.functional as F

from . import _lib
from .criterion import Criterion
from .Domain import NamedEmbedding
from .functional import ScoreFunction
from .KGMeta import PairwiseModel


class ModelSpec:
    """Static description handed to the autograd functions."""

    def __init__(self, name, dim, rel_dim=None, l1_flag=False, margin=0.0, phase_scale=0.0):
        self.name, self.dim, self.rel_dim = name, dim, rel_dim if rel_dim is not None else dim
        self.l1_flag, self.margin, self.phase_scale = l1_flag, margin, phase_scale

    def desc(self, tables):
        return _lib.ModelDesc(self.name, tables, self.dim, rel_dim=self.rel_dim, l1_flag=self.l1_flag,
                              margin=self.margin, phase_scale=self.phase_scale)


class _KernelScored:
    """Mixin: forward() through the CUDA kernel; kge_tables() gives the weights in C-ABI order."""

    def kge_tables(self):
        raise NotImplementedError

    def kge_spec(self):
        raise NotImplementedError

    def kge_desc(self):
        """ModelDesc over the live weights (no copy) for the rank / fused-train entry points."""
        return self.kge_spec().desc([w.detach() for w in self.kge_tables()])

    def kge_pre_score(self):
        """Side effects the reference's forward() has on the tables BEFORE scoring (Rescal's in-place
        row normalisation, pairwise.py:843-844).  The fused training steps and Evaluator.rank_triples
        call the kernels directly, so they call this first; no-op for every other model."""

    def kge_fused_reg(self):
        """(reg_code, lmbda) of the row regulariser get_reg() applies with its DEFAULT reg_type (what
        Trainer.train_step_pointwise calls, trainer.py:179) for kge_reg_fwd_bwd, or None when get_reg is
        not a function of the weights (then the fused step adds its value as a constant)."""
        return None

    def forward(self, h, r, t):
        return ScoreFunction.apply(self.kge_spec(), h, r, t, *self.kge_tables())


class TransE(_KernelScored, PairwiseModel):
    """pykg2vec/models/pairwise.py:12-93."""

    def __init__(self, **kwargs):
        super(TransE, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "l1_flag"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, self.hidden_size)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, self.hidden_size)
        nn.init.xavier_uniform_(self.ent_embeddings.weight)
        nn.init.xavier_uniform_(self.rel_embeddings.weight)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings]
        self.loss = Criterion.pairwise_hinge

    def kge_tables(self):
        return [self.ent_embeddings.weight, self.rel_embeddings.weight]

    def kge_spec(self):
        return ModelSpec("transe", self.hidden_size, l1_flag=bool(self.l1_flag))

    def embed(self, h, r, t):
        return self.ent_embeddings(h), self.rel_embeddings(r), self.ent_embeddings(t)


class TransH(_KernelScored, PairwiseModel):
    """pykg2vec/models/pairwise.py:96-182."""

    def __init__(self, **kwargs):
        super(TransH, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "l1_flag"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, self.hidden_size)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, self.hidden_size)
        self.w = NamedEmbedding("w", self.tot_relation, self.hidden_size)
        nn.init.xavier_uniform_(self.ent_embeddings.weight)
        nn.init.xavier_uniform_(self.rel_embeddings.weight)
        nn.init.xavier_uniform_(self.w.weight)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings, self.w]
        self.loss = Criterion.pairwise_hinge

    def kge_tables(self):
        return [self.ent_embeddings.weight, self.rel_embeddings.weight, self.w.weight]

    def kge_spec(self):
        return ModelSpec("transh", self.hidden_size, l1_flag=bool(self.l1_flag))

    def embed(self, h, r, t):
        emb_h, emb_r, emb_t = self.ent_embeddings(h), self.rel_embeddings(r), self.ent_embeddings(t)
        proj_vec = self.w(r)
        return self._projection(emb_h, proj_vec), emb_r, self._projection(emb_t, proj_vec)

    @staticmethod
    def _projection(emb_e, proj_vec):
        proj_vec = F.normalize(proj_vec, p=2, dim=-1)
        return emb_e - torch.sum(emb_e * proj_vec, dim=-1, keepdim=True) * proj_vec


class TransD(_KernelScored, PairwiseModel):
    """pykg2vec/models/pairwise.py:185-278 (valid only for ent_hidden_size == rel_hidden_size,
    like the reference whose _projection broadcast fails otherwise)."""

    def __init__(self, **kwargs):
        super(TransD, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "rel_hidden_size", "ent_hidden_size", "l1_flag"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, self.ent_hidden_size)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, self.rel_hidden_size)
        self.ent_mappings = NamedEmbedding("ent_mappings", self.tot_entity, self.ent_hidden_size)
        self.rel_mappings = NamedEmbedding("rel_mappings", self.tot_relation, self.rel_hidden_size)
        for e in (self.ent_embeddings, self.rel_embeddings, self.ent_mappings, self.rel_mappings):
            nn.init.xavier_uniform_(e.weight)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings, self.ent_mappings, self.rel_mappings]
        self.loss = Criterion.pairwise_hinge

    def kge_tables(self):
        return [self.ent_embeddings.weight, self.rel_embeddings.weight, self.ent_mappings.weight,
                self.rel_mappings.weight]

    def kge_spec(self):
        if self.ent_hidden_size != self.rel_hidden_size:
            raise RuntimeError("TransD requires ent_hidden_size == rel_hidden_size (the reference's "
                               "_projection broadcast, pairwise.py:275-278)")
        return ModelSpec("transd", self.ent_hidden_size, l1_flag=bool(self.l1_flag))

    def embed(self, h, r, t):
        emb_h, emb_r, emb_t = self.ent_embeddings(h), self.rel_embeddings(r), self.ent_embeddings(t)
        h_m, r_m, t_m = self.ent_mappings(h), self.rel_mappings(r), self.ent_mappings(t)
        return self._projection(emb_h, h_m, r_m), emb_r, self._projection(emb_t, t_m, r_m)

    @staticmethod
    def _projection(emb_e, emb_m, proj_vec):
        return emb_e + torch.sum(emb_e * emb_m, dim=-1, keepdim=True) * proj_vec


class TransM(_KernelScored, PairwiseModel):
    """pykg2vec/models/pairwise.py:281-364.  theta is a per-relation constant computed from
    the training triples at construction (pairwise.py:304-314)."""

    def __init__(self, **kwargs):
        super(TransM, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "l1_flag"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, self.hidden_size)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, self.hidden_size)
        rel_head = {x: [] for x in range(self.tot_relation)}
        rel_tail = {x: [] for x in range(self.tot_relation)}
        rel_counts = {x: 0 for x in range(self.tot_relation)}
        for tr in kwargs["knowledge_graph"].read_cache_data('triplets_train'):
            rel_head[tr.r].append(tr.h)
            rel_tail[tr.r].append(tr.t)
            rel_counts[tr.r] += 1
        theta = [1 / np.log(2 + rel_counts[x] / (1 + len(rel_tail[x])) + rel_counts[x] / (1 + len(rel_head[x])))
                 for x in range(self.tot_relation)]
        self.theta = torch.from_numpy(np.asarray(theta, dtype=np.float32)).to(kwargs["device"])
        nn.init.xavier_uniform_(self.ent_embeddings.weight)
        nn.init.xavier_uniform_(self.rel_embeddings.weight)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings]
        self.loss = Criterion.pairwise_hinge

    def kge_tables(self):
        return [self.ent_embeddings.weight, self.rel_embeddings.weight, self.theta]

    def kge_spec(self):
        return ModelSpec("transm", self.hidden_size, l1_flag=bool(self.l1_flag))

    def embed(self, h, r, t):
        return self.ent_embeddings(h), self.rel_embeddings(r), self.ent_embeddings(t)


class TransR(_KernelScored, PairwiseModel):
    """pykg2vec/models/pairwise.py:367-470."""

    def __init__(self, **kwargs):
        super(TransR, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "rel_hidden_size", "ent_hidden_size", "l1_flag"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, self.ent_hidden_size)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, self.rel_hidden_size)
        self.rel_matrix = NamedEmbedding("rel_matrix", self.tot_relation, self.ent_hidden_size * self.rel_hidden_size)
        nn.init.xavier_uniform_(self.ent_embeddings.weight)
        nn.init.xavier_uniform_(self.rel_embeddings.weight)
        nn.init.xavier_uniform_(self.rel_matrix.weight)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings, self.rel_matrix]
        self.loss = Criterion.pairwise_hinge

    def kge_tables(self):
        return [self.ent_embeddings.weight, self.rel_embeddings.weight, self.rel_matrix.weight]

    def kge_spec(self):
        return ModelSpec("transr", self.ent_hidden_size, rel_dim=self.rel_hidden_size, l1_flag=bool(self.l1_flag))

    def embed(self, h, r, t):
        h_e = F.normalize(self.ent_embeddings(h), p=2, dim=-1)
        r_e = F.normalize(self.rel_embeddings(r), p=2, dim=-1)
        t_e = F.normalize(self.ent_embeddings(t), p=2, dim=-1)
        m = self.rel_matrix(r).view(-1, self.ent_hidden_size, self.rel_hidden_size)
        return torch.matmul(h_e.unsqueeze(1), m).squeeze(1), r_e, torch.matmul(t_e.unsqueeze(1), m).squeeze(1)


class RotatE(_KernelScored, PairwiseModel):
    """pykg2vec/models/pairwise.py:727-791."""

    def __init__(self, **kwargs):
        super(RotatE, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "margin"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.embedding_range = (self.margin + 2.0) / self.hidden_size
        self.ent_embeddings = NamedEmbedding("ent_embeddings_real", self.tot_entity, self.hidden_size)
        self.ent_embeddings_imag = NamedEmbedding("ent_embeddings_imag", self.tot_entity, self.hidden_size)
        self.rel_embeddings = NamedEmbedding("rel_embeddings_real", self.tot_relation, self.hidden_size)
        nn.init.uniform_(self.ent_embeddings.weight, -self.embedding_range, self.embedding_range)
        nn.init.uniform_(self.ent_embeddings_imag.weight, -self.embedding_range, self.embedding_range)
        nn.init.uniform_(self.rel_embeddings.weight, -self.embedding_range, self.embedding_range)
        self.parameter_list = [self.ent_embeddings, self.ent_embeddings_imag, self.rel_embeddings]
        self.loss = Criterion.pariwise_logistic

    def kge_tables(self):
        return [self.ent_embeddings.weight, self.ent_embeddings_imag.weight, self.rel_embeddings.weight]

    def kge_spec(self):
        # theta = r / (embedding_range / pi)  (pairwise.py:776-782) as one fp32 multiplier
        phase = float(np.float32(math.pi / self.embedding_range))
        return ModelSpec("rotate", self.hidden_size, margin=float(self.margin), phase_scale=phase)

    def embed(self, h, r, t):
        pi = 3.14159265358979323846
        r_e_r = self.rel_embeddings(r) / (self.embedding_range / pi)
        return (self.ent_embeddings(h), self.ent_embeddings_imag(h), torch.cos(r_e_r), torch.sin(r_e_r),
                self.ent_embeddings(t), self.ent_embeddings_imag(t))


class Rescal(_KernelScored, PairwiseModel):
    """pykg2vec/models/pairwise.py:794-865.  Like the reference, every forward() first replaces
    both tables by their row-normalised versions IN PLACE (embed(), :843-844) — here one kernel
    per table (kge_normalize_rows) — and then scores -h^T M_r t."""

    def __init__(self, **kwargs):
        super(Rescal, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "margin"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, self.hidden_size)
        self.rel_matrices = NamedEmbedding("rel_matrices", self.tot_relation, self.hidden_size * self.hidden_size)
        nn.init.xavier_uniform_(self.ent_embeddings.weight)
        nn.init.xavier_uniform_(self.rel_matrices.weight)
        self.parameter_list = [self.ent_embeddings, self.rel_matrices]
        self.loss = Criterion.pairwise_hinge

    def kge_tables(self):
        return [self.ent_embeddings.weight, self.rel_matrices.weight]

    def kge_spec(self):
        return ModelSpec("rescal", self.hidden_size)

    def normalize_tables_(self):
        with torch.no_grad():
            _lib.normalize_rows(self.ent_embeddings.weight.data)
            _lib.normalize_rows(self.rel_matrices.weight.data)

    def kge_pre_score(self):
        self.normalize_tables_()

    def forward(self, h, r, t):
        self.normalize_tables_()
        return ScoreFunction.apply(self.kge_spec(), h, r, t, *self.kge_tables())

    def embed(self, h, r, t):
        k = self.hidden_size
        self.normalize_tables_()
        return (self.ent_embeddings(h).view(-1, k, 1), self.rel_matrices(r).view(-1, k, k),
                self.ent_embeddings(t).view(-1, k, 1))


class HoLE(_KernelScored, PairwiseModel):
    """pykg2vec/models/pairwise.py:1087-1142.  forward() reproduces what the reference's lines
    :1119-1125 evaluate under its pinned torch<1.7: e = circconv(even(h),
    even(t)), score = -sigmoid(<normalize(r), e>).  The legacy torch.fft/ifft the reference
    calls no longer exist, so parity for HoLE is pinned on an emulation (tests/golden)."""

    def __init__(self, **kwargs):
        super(HoLE, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "cmax", "cmin"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, self.hidden_size)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, self.hidden_size)
        nn.init.xavier_uniform_(self.ent_embeddings.weight)
        nn.init.xavier_uniform_(self.rel_embeddings.weight)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings]
        self.loss = Criterion.pairwise_hinge

    def kge_tables(self):
        return [self.ent_embeddings.weight, self.rel_embeddings.weight]

    def kge_spec(self):
        return ModelSpec("hole", self.hidden_size)

    def embed(self, h, r, t):
        return self.ent_embeddings(h), self.rel_embeddings(r), self.ent_embeddings(t)


class KG2E(_KernelScored, PairwiseModel):
    """pykg2vec/models/pairwise.py:966-1084 (KL-divergence score on row-normalised means/variances)."""

    def __init__(self, **kwargs):
        super(KG2E, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "cmax", "cmin"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.ent_embeddings_mu = NamedEmbedding("ent_embeddings_mu", self.tot_entity, self.hidden_size)
        self.rel_embeddings_mu = NamedEmbedding("rel_embeddings_mu", self.tot_relation, self.hidden_size)
        self.ent_embeddings_sigma = NamedEmbedding("ent_embeddings_sigma", self.tot_entity, self.hidden_size)
        self.rel_embeddings_sigma = NamedEmbedding("rel_embeddings_sigma", self.tot_relation, self.hidden_size)
        for e in (self.ent_embeddings_mu, self.rel_embeddings_mu, self.ent_embeddings_sigma, self.rel_embeddings_sigma):
            nn.init.xavier_uniform_(e.weight)
        self.parameter_list = [self.ent_embeddings_mu, self.ent_embeddings_sigma, self.rel_embeddings_mu,
                               self.rel_embeddings_sigma]
        # sigma <- clamp(sigma + 1, cmin, cmax), re-registered as new Parameters (pairwise.py:1011-1014)
        for e in (self.ent_embeddings_sigma, self.rel_embeddings_sigma):
            e.weight = nn.Parameter(torch.clamp(e.weight.detach() + 1.0, min=float(self.cmin), max=float(self.cmax)))
        self.loss = Criterion.pairwise_hinge

    def kge_tables(self):
        return [e.weight for e in self.parameter_list]

    def kge_spec(self):
        return ModelSpec("kg2e", self.hidden_size)

    @staticmethod
    def get_normalized_data(embedding, p=2, dim=1):
        norms = torch.norm(embedding, p, dim)
        return embedding.div(norms.view(-1, 1).expand_as(embedding))

    def embed(self, h, r, t):
        n = self.get_normalized_data
        return (n(self.ent_embeddings_mu(h)), n(self.ent_embeddings_sigma(h)), n(self.rel_embeddings_mu(r)),
                n(self.rel_embeddings_sigma(r)), n(self.ent_embeddings_mu(t)), n(self.ent_embeddings_sigma(t)))


class _DenseLayerModel(_KernelScored, PairwiseModel):
    """Models whose trailing tables are global dense parameters (not indexed by ids): the fused
    sparse optimizer does not apply to them, training goes through autograd + a torch optimizer."""
    kge_dense_params = True

    def kge_tables(self):
        return [e.weight for e in self.parameter_list]

    def embed(self, h, r, t):
        return self.ent_embeddings(h), self.rel_embeddings(r), self.ent_embeddings(t)


class SLM(_DenseLayerModel):
    """pykg2vec/models/pairwise.py:473-541."""

    def __init__(self, **kwargs):
        super(SLM, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "rel_hidden_size", "ent_hidden_size"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, self.ent_hidden_size)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, self.rel_hidden_size)
        self.mr1 = NamedEmbedding("mr1", self.ent_hidden_size, self.rel_hidden_size)
        self.mr2 = NamedEmbedding("mr2", self.ent_hidden_size, self.rel_hidden_size)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings, self.mr1, self.mr2]
        for e in self.parameter_list:
            nn.init.xavier_uniform_(e.weight)
        self.loss = Criterion.pairwise_hinge

    def kge_spec(self):
        return ModelSpec("slm", self.ent_hidden_size, rel_dim=self.rel_hidden_size)


class NTN(_DenseLayerModel):
    """pykg2vec/models/pairwise.py:868-963 (fp32 CUDA-core kernels; the grouped tensor-core
    formulation of the bilinear term is round-2 work)."""

    def __init__(self, **kwargs):
        super(NTN, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "ent_hidden_size", "rel_hidden_size", "lmbda"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        d, k = self.ent_hidden_size, self.rel_hidden_size
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, d)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, k)
        self.mr1 = NamedEmbedding("mr1", d, k)
        self.mr2 = NamedEmbedding("mr2", d, k)
        self.br = NamedEmbedding("br", 1, k)
        self.mr = NamedEmbedding("mr", k, d * d)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings, self.mr1, self.mr2, self.br, self.mr]
        for e in self.parameter_list:
            nn.init.xavier_uniform_(e.weight)
        self.loss = Criterion.pairwise_hinge

    def kge_spec(self):
        return ModelSpec("ntn", self.ent_hidden_size, rel_dim=self.rel_hidden_size)

    def get_reg(self, h, r, t):
        # lmbda * sqrt(sum over ALL parameters of ||theta||^2) (pairwise.py:962-963): a whole-table
        # reduction, not part of the batch score; plain tensor code as in the reference
        return self.lmbda * torch.sqrt(sum(torch.sum(torch.pow(var.weight, 2)) for var in self.parameter_list))


class SME(_DenseLayerModel):
    """pykg2vec/models/pairwise.py:544-661 (linear semantic matching energy)."""
    _kge_name = "sme"

    def __init__(self, **kwargs):
        super(SME, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        d = self.hidden_size
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, d)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, d)
        self.mu1 = NamedEmbedding("mu1", d, d)
        self.mu2 = NamedEmbedding("mu2", d, d)
        self.bu = NamedEmbedding("bu", d, 1)
        self.mv1 = NamedEmbedding("mv1", d, d)
        self.mv2 = NamedEmbedding("mv2", d, d)
        self.bv = NamedEmbedding("bv", d, 1)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings, self.mu1, self.mu2, self.bu, self.mv1,
                               self.mv2, self.bv]
        for e in self.parameter_list:
            nn.init.xavier_uniform_(e.weight)
        self.loss = Criterion.pairwise_hinge

    def kge_spec(self):
        return ModelSpec(self._kge_name, self.hidden_size)


class SME_BL(SME):
    """pykg2vec/models/pairwise.py:664-724 (bilinear variant; note the POSITIVE sign of its score)."""
    _kge_name = "sme_bl"

    def __init__(self, **kwargs):
        super(SME_BL, self).__init__(**kwargs)
        self.model_name = self.__class__.__name__.lower()
        self.loss = Criterion.pairwise_hinge
