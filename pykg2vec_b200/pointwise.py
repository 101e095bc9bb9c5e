"""Pointwise (logistic) models — same surface as pykg2vec/models/pointwise.py; forward()
and get_reg() are CUDA kernels (kge_score_fwd / kge_reg_fwd_bwd)."""
import torch
import torch.nn as nn

from .criterion import Criterion
from .Domain import NamedEmbedding
from .functional import RegFunction
from .KGMeta import PointwiseModel
from .pairwise import ModelSpec, _KernelScored

_REG_TYPES = {"f2": 0, "n3": 1}


class _RowRegularised(_KernelScored):
    _abs_n3 = False
    _default_reg = "F2"   # default reg_type of the model's get_reg (what the Trainer calls)

    def kge_fused_reg(self):
        code = _REG_TYPES[self._default_reg.lower()]
        if code == 1 and self._abs_n3:
            code = 2
        return code, float(self.lmbda)

    def _reg(self, h, r, t, reg_type):
        key = reg_type.lower()
        if key not in _REG_TYPES:
            raise NotImplementedError('Unknown regularizer type: %s' % reg_type)
        code = _REG_TYPES[key]
        if code == 1 and self._abs_n3:
            code = 2  # ComplexN3: |x|**3 (pointwise.py:224-238); |x|**2 == x**2 for its "f2"
        return RegFunction.apply(self.kge_spec(), code, float(self.lmbda), h, r, t, *self.kge_tables())


class DistMult(_RowRegularised, PointwiseModel):
    """pykg2vec/models/pointwise.py:391-458."""

    def __init__(self, **kwargs):
        super(DistMult, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "lmbda"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, self.hidden_size)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, self.hidden_size)
        nn.init.xavier_uniform_(self.ent_embeddings.weight)
        nn.init.xavier_uniform_(self.rel_embeddings.weight)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings]
        self.loss = Criterion.pointwise_logistic

    def kge_tables(self):
        return [self.ent_embeddings.weight, self.rel_embeddings.weight]

    def kge_spec(self):
        return ModelSpec("distmult", self.hidden_size)

    def embed(self, h, r, t):
        return self.ent_embeddings(h), self.rel_embeddings(r), self.ent_embeddings(t)

    def get_reg(self, h, r, t, reg_type="F2"):
        return self._reg(h, r, t, reg_type)


class CP(_RowRegularised, PointwiseModel):
    """pykg2vec/models/pointwise.py:321-388."""
    _default_reg = "N3"

    def __init__(self, **kwargs):
        super(CP, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "lmbda"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.sub_embeddings = NamedEmbedding("sub_embedding", self.tot_entity, self.hidden_size)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, self.hidden_size)
        self.obj_embeddings = NamedEmbedding("obj_embedding", self.tot_entity, self.hidden_size)
        nn.init.xavier_uniform_(self.sub_embeddings.weight)
        nn.init.xavier_uniform_(self.rel_embeddings.weight)
        nn.init.xavier_uniform_(self.obj_embeddings.weight)
        self.parameter_list = [self.sub_embeddings, self.rel_embeddings, self.obj_embeddings]
        self.loss = Criterion.pointwise_logistic

    def kge_tables(self):
        return [self.sub_embeddings.weight, self.rel_embeddings.weight, self.obj_embeddings.weight]

    def kge_spec(self):
        return ModelSpec("cp", self.hidden_size)

    def embed(self, h, r, t):
        return self.sub_embeddings(h), self.rel_embeddings(r), self.obj_embeddings(t)

    def get_reg(self, h, r, t, reg_type='N3'):
        return self._reg(h, r, t, reg_type)


class Complex(_RowRegularised, PointwiseModel):
    """pykg2vec/models/pointwise.py:122-202."""

    def __init__(self, **kwargs):
        super(Complex, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "lmbda"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        k = self.hidden_size
        self.ent_embeddings_real = NamedEmbedding("emb_e_real", self.tot_entity, k)
        self.ent_embeddings_img = NamedEmbedding("emb_e_img", self.tot_entity, k)
        self.rel_embeddings_real = NamedEmbedding("emb_rel_real", self.tot_relation, k)
        self.rel_embeddings_img = NamedEmbedding("emb_rel_img", self.tot_relation, k)
        for e in (self.ent_embeddings_real, self.ent_embeddings_img, self.rel_embeddings_real,
                  self.rel_embeddings_img):
            nn.init.xavier_uniform_(e.weight)
        self.parameter_list = [self.ent_embeddings_real, self.ent_embeddings_img, self.rel_embeddings_real,
                               self.rel_embeddings_img]
        self.loss = Criterion.pointwise_logistic

    def kge_tables(self):
        return [self.ent_embeddings_real.weight, self.ent_embeddings_img.weight,
                self.rel_embeddings_real.weight, self.rel_embeddings_img.weight]

    def kge_spec(self):
        return ModelSpec("complex", self.hidden_size)

    def embed(self, h, r, t):
        return (self.ent_embeddings_real(h), self.ent_embeddings_img(h), self.rel_embeddings_real(r),
                self.rel_embeddings_img(r), self.ent_embeddings_real(t), self.ent_embeddings_img(t))

    def get_reg(self, h, r, t, reg_type="F2"):
        return self._reg(h, r, t, reg_type)


class ComplexN3(Complex):
    """pykg2vec/models/pointwise.py:205-238."""
    _abs_n3 = True
    _default_reg = "N3"

    def __init__(self, **kwargs):
        super(ComplexN3, self).__init__(**kwargs)
        self.model_name = 'complexn3'
        self.loss = Criterion.pointwise_logistic

    def get_reg(self, h, r, t, reg_type="N3"):
        return self._reg(h, r, t, reg_type)


class SimplE(_KernelScored, PointwiseModel):
    """pykg2vec/models/pointwise.py:461-536 (including its quirks: only the inverse-relation sum
    is halved, :525; get_reg() regularises the id tensors, :528-536)."""
    _kge_name = "simple"

    def __init__(self, **kwargs):
        super(SimplE, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "lmbda"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        k = self.hidden_size
        self.tot_train_triples = kwargs['tot_train_triples']
        self.batch_size = kwargs['batch_size']
        self.ent_head_embeddings = NamedEmbedding("ent_head_embedding", self.tot_entity, k)
        self.ent_tail_embeddings = NamedEmbedding("ent_tail_embedding", self.tot_entity, k)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, k)
        self.rel_inv_embeddings = NamedEmbedding("rel_inv_embedding", self.tot_relation, k)
        for e in (self.ent_head_embeddings, self.ent_tail_embeddings, self.rel_embeddings, self.rel_inv_embeddings):
            nn.init.xavier_uniform_(e.weight)
        self.parameter_list = [self.ent_head_embeddings, self.ent_tail_embeddings, self.rel_embeddings,
                               self.rel_inv_embeddings]
        self.loss = Criterion.pointwise_logistic

    def kge_tables(self):
        return [self.ent_head_embeddings.weight, self.ent_tail_embeddings.weight, self.rel_embeddings.weight,
                self.rel_inv_embeddings.weight]

    def kge_spec(self):
        return ModelSpec(self._kge_name, self.hidden_size)

    def embed(self, h, r, t):
        return (self.ent_head_embeddings(h), self.ent_head_embeddings(t), self.rel_embeddings(r),
                self.rel_inv_embeddings(r), self.ent_tail_embeddings(t), self.ent_tail_embeddings(h))

    def get_reg(self, h, r, t, reg_type="F2"):
        # the reference applies the regulariser to the ID tensors (a constant w.r.t. the weights)
        import torch
        p = {"f2": 2, "n3": 3}.get(reg_type.lower())
        if p is None:
            raise NotImplementedError('Unknown regularizer type: %s' % reg_type)
        term = torch.mean(torch.sum(h.float() ** p, -1) + torch.sum(r.float() ** p, -1) + torch.sum(t.float() ** p, -1))
        return self.lmbda * term


class SimplE_ignr(SimplE):
    """pykg2vec/models/pointwise.py:539-581."""
    _kge_name = "simple_ignr"

    def __init__(self, **kwargs):
        super(SimplE_ignr, self).__init__(**kwargs)
        self.model_name = 'simple_ignr'
        self.loss = Criterion.pointwise_logistic

    def embed(self, h, r, t):
        import torch
        cat = lambda e1, i1, e2, i2: torch.cat([e1.weight[i1], e2.weight[i2]], 1)
        return (cat(self.ent_head_embeddings, h, self.ent_head_embeddings, t),
                cat(self.rel_embeddings, r, self.rel_inv_embeddings, r),
                cat(self.ent_tail_embeddings, t, self.ent_tail_embeddings, h))


class ANALOGY(_RowRegularised, PointwiseModel):
    """pykg2vec/models/pointwise.py:13-119: ComplEx on half-width tables + DistMult on full-width."""

    def __init__(self, **kwargs):
        super(ANALOGY, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "lmbda"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        k = self.hidden_size
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, k)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, k)
        self.ent_embeddings_real = NamedEmbedding("emb_e_real", self.tot_entity, k // 2)
        self.ent_embeddings_img = NamedEmbedding("emb_e_img", self.tot_entity, k // 2)
        self.rel_embeddings_real = NamedEmbedding("emb_rel_real", self.tot_relation, k // 2)
        self.rel_embeddings_img = NamedEmbedding("emb_rel_img", self.tot_relation, k // 2)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings, self.ent_embeddings_real,
                               self.ent_embeddings_img, self.rel_embeddings_real, self.rel_embeddings_img]
        for e in self.parameter_list:
            nn.init.xavier_uniform_(e.weight)
        self.loss = Criterion.pointwise_logistic

    def kge_tables(self):
        return [e.weight for e in self.parameter_list]

    def kge_spec(self):
        return ModelSpec("analogy", self.hidden_size)

    def embed(self, h, r, t):
        return self.ent_embeddings(h), self.rel_embeddings(r), self.ent_embeddings(t)

    def embed_complex(self, h, r, t):
        return (self.ent_embeddings_real(h), self.ent_embeddings_img(h), self.rel_embeddings_real(r),
                self.rel_embeddings_img(r), self.ent_embeddings_real(t), self.ent_embeddings_img(t))

    def get_reg(self, h, r, t, reg_type="F2"):
        return self._reg(h, r, t, reg_type)


class _HyperComplex(_KernelScored, PointwiseModel):
    """Shared surface of QuatE / OctonionE: C entity-part tables, C relation-part tables (+ the unused
    rel_w table the reference also registers), score = -sum <h (x) unit(r), t>."""
    _kge_name = None
    _parts = 0

    def kge_tables(self):
        return [e.weight for e in self.parameter_list[:2 * self._parts]]

    def kge_spec(self):
        return ModelSpec(self._kge_name, self.hidden_size)

    def kge_fused_reg(self):
        return 2, float(self.lmbda)   # get_reg's default reg_type is 'N3' = |x|**3 (pointwise.py:696-727, :901-960)

    def get_reg(self, h, r, t, reg_type='N3'):
        key = reg_type.lower()
        if key not in ("f2", "n3"):
            raise NotImplementedError('Unknown regularizer type: %s' % reg_type)
        # |x|**2 == x**2 (code 0); |x|**3 (code 2); averaged over batch and width by the kernel
        return RegFunction.apply(self.kge_spec(), 0 if key == "f2" else 2, float(self.lmbda), h, r, t,
                                 *self.kge_tables())


class QuatE(_HyperComplex):
    """pykg2vec/models/pointwise.py:584-769 (fc / bn / dropout members are registered but unused by
    forward(), exactly as in the reference, so checkpoints keep their keys)."""
    _kge_name, _parts = "quate", 4

    def __init__(self, **kwargs):
        super(QuatE, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "lmbda"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        k = self.hidden_size
        self.ent_s_embedding = NamedEmbedding("ent_s_embedding", self.tot_entity, k)
        self.ent_x_embedding = NamedEmbedding("ent_x_embedding", self.tot_entity, k)
        self.ent_y_embedding = NamedEmbedding("ent_y_embedding", self.tot_entity, k)
        self.ent_z_embedding = NamedEmbedding("ent_z_embedding", self.tot_entity, k)
        self.rel_s_embedding = NamedEmbedding("rel_s_embedding", self.tot_relation, k)
        self.rel_x_embedding = NamedEmbedding("rel_x_embedding", self.tot_relation, k)
        self.rel_y_embedding = NamedEmbedding("rel_y_embedding", self.tot_relation, k)
        self.rel_z_embedding = NamedEmbedding("rel_z_embedding", self.tot_relation, k)
        self.rel_w_embedding = NamedEmbedding("rel_w_embedding", self.tot_relation, k)
        self.fc = nn.Linear(100, 50, bias=False)
        self.ent_dropout = nn.Dropout(0)
        self.rel_dropout = nn.Dropout(0)
        self.bn = nn.BatchNorm1d(k)
        self.parameter_list = [self.ent_s_embedding, self.ent_x_embedding, self.ent_y_embedding, self.ent_z_embedding,
                               self.rel_s_embedding, self.rel_x_embedding, self.rel_y_embedding, self.rel_z_embedding,
                               self.rel_w_embedding]
        # Reference quirk kept for checkpoint compatibility (pointwise.py:621-636): the four relation
        # part tables are re-assigned data drawn with in_features = tot_entity, so their weights have
        # tot_entity rows (only the first tot_relation are ever addressed); everything is then
        # overwritten by xavier_uniform_.
        import torch
        for e in (self.rel_s_embedding, self.rel_x_embedding, self.rel_y_embedding, self.rel_z_embedding):
            e.weight.data = torch.empty(self.tot_entity, k)
        for e in self.parameter_list:
            nn.init.xavier_uniform_(e.weight.data)
        self.loss = Criterion.pointwise_logistic

    def embed(self, h, r, t):
        e, rl = self.parameter_list[:4], self.parameter_list[4:8]
        return tuple(x(h) for x in e) + tuple(x(t) for x in e) + tuple(x(r) for x in rl)


class OctonionE(_HyperComplex):
    """pykg2vec/models/pointwise.py:772-1002."""
    _kge_name, _parts = "octonione", 8

    def __init__(self, **kwargs):
        super(OctonionE, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "lmbda"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        k = self.hidden_size
        self.parameter_list = []
        for i in range(1, 9):
            emb = NamedEmbedding("ent_embedding_%d" % i, self.tot_entity, k)
            setattr(self, "ent_embedding_%d" % i, emb)
            self.parameter_list.append(emb)
        for i in range(1, 9):
            emb = NamedEmbedding("rel_embedding_%d" % i, self.tot_relation, k)
            setattr(self, "rel_embedding_%d" % i, emb)
            self.parameter_list.append(emb)
        self.rel_w_embedding = NamedEmbedding("rel_w_embedding", self.tot_relation, k)
        self.parameter_list.append(self.rel_w_embedding)
        for e in self.parameter_list:
            nn.init.xavier_uniform_(e.weight.data)
        self.loss = Criterion.pointwise_logistic

    def embed(self, h, r, t):
        e, rl = self.parameter_list[:8], self.parameter_list[8:16]
        return tuple(x(h) for x in e) + tuple(x(t) for x in e) + tuple(x(r) for x in rl)


class ConvKB(_KernelScored, PointwiseModel):
    """pykg2vec/models/pointwise.py:241-318.

    The reference stacks [h; r; t] into a [b,1,3,k] image, applies Conv2d(1 -> F, (3, w)) for every
    filter size w, concatenates, flattens and applies Linear -> 1 — with NO nonlinearity in between
    (pointwise.py:311-316).  The score is therefore an affine map of the three rows,
        score = <a_h, h> + <a_r, r> + <a_t, t> + c0,
    and the per-triple work is three length-k dot products: that is what the CUDA kernel evaluates
    (KGE_CONVKB, include/kge_b200.h).  The [3,k] coefficient rows and c0 are the collapse of the
    convolution filters with the Linear weights (`_collapse`, tiny: 3k outputs), recomputed from the
    live parameters on every call and differentiable, so `fc1` trains through autograd exactly as in
    the reference.  As in the reference, `conv_list` is a plain Python list: its filters are not
    registered parameters, are not in state_dict() and are never updated by the optimizer."""
    kge_dense_params = True

    def __init__(self, **kwargs):
        super(ConvKB, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "num_filters", "filter_sizes"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        k = self.hidden_size
        device = kwargs["device"]
        self.filter_sizes = [int(w) for w in self.filter_sizes]
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, k)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, k)
        nn.init.xavier_uniform_(self.ent_embeddings.weight)
        nn.init.xavier_uniform_(self.rel_embeddings.weight)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings]
        self.conv_list = [nn.Conv2d(1, self.num_filters, (3, w), stride=(1, 1)).to(device) for w in self.filter_sizes]
        conv_out_dim = self.num_filters * sum(k - w + 1 for w in self.filter_sizes)
        self.fc1 = nn.Linear(in_features=conv_out_dim, out_features=1, bias=True)
        self.loss = Criterion.pointwise_logistic

    def _collapse(self):
        """(A [3,k], c0 [1]):  A[row, j] = sum_w sum_f sum_q K_w[f,0,row,q] * W[f, off_w + j - q],
        c0 = fc_b + sum_w sum_f b_w[f] * sum_p W[f, off_w + p],  W = fc1.weight.view(F, -1)
        (the concat is along the width, the flatten is filter-major: index f*sumP + off_w + p)."""
        k, nf = self.hidden_size, self.num_filters
        W = self.fc1.weight.view(nf, -1)
        dev = W.device
        A = torch.zeros((3, k), dtype=W.dtype, device=dev)
        c0 = self.fc1.bias.reshape(1)
        off = 0
        for conv, w in zip(self.conv_list, self.filter_sizes):
            P = k - w + 1
            Wf = W[:, off:off + P]
            K = conv.weight.detach().to(dev)[:, 0]                        # [F, 3, w], not trained
            parts = [torch.nn.functional.pad(torch.matmul(K[:, :, q].t(), Wf), (q, w - 1 - q)) for q in range(w)]
            A = A + sum(parts)
            c0 = c0 + torch.dot(conv.bias.detach().to(dev), Wf.sum(1)).reshape(1)
            off += P
        return A.contiguous(), c0

    def kge_tables(self):
        A, c0 = self._collapse()
        return [self.ent_embeddings.weight, self.rel_embeddings.weight, A, c0]

    def kge_spec(self):
        return ModelSpec("convkb", self.hidden_size)

    def embed(self, h, r, t):
        return self.ent_embeddings(h), self.rel_embeddings(r), self.ent_embeddings(t)
