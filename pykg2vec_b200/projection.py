"""Projection models — same surface as pykg2vec/models/projection.py.

ConvE (projection.py:12-125).  The reference's hot spot for these models is the last layer,
`sigmoid(x . E^T + b)` against ALL entities, its BCE loss over the dense [b, N] label matrix
and — at evaluation — one such forward plus a full `topk` per query and direction.  Those are
the hand-written kernels here (include/kge_b200.h kge_proj_*): forward() ends in
ProjTailFunction, the loss is Criterion.multi_class_bce -> kge_proj_bce, and the Evaluator ranks
batches of queries with kge_proj_rank (counts only, no [Q, N] matrix, no sort).

The trunk in front of it (BatchNorm -> 3x3 conv -> BatchNorm -> ReLU -> Linear -> BatchNorm ->
ReLU on a [b, 1, 2*h2, h1] image) runs in two ways:
  * training (autograd needed, BatchNorm batch statistics): the reference's own torch layers,
    with TF32 convolutions disabled in forward and backward so the trunk stays fp32 like the
    reference's CPU path;
  * evaluation (`proj_query`, used by the batched Evaluator): kge_conve_trunk_fwd — gather +
    BN0 + conv + BN1 + ReLU in one kernel and the Linear layer through the same tiled GEMM as the
    tail.
"""
import torch
import torch.nn as nn

from . import _lib
from .criterion import Criterion
from .Domain import NamedEmbedding
from .functional import ProjTailFunction, _require_cuda
from .KGMeta import ProjectionModel


class _Fp32Conv2d(torch.autograd.Function):
    """conv2d whose forward AND backward run with TF32 convolutions disabled: the reference's
    numbers are the fp32 CPU path's, and cuDNN would otherwise round operands to 10-bit
    mantissas (torch.backends.cudnn.allow_tf32 defaults to True).  A context manager around the
    forward call alone would not cover the backward pass, which runs later inside
    loss.backward()."""

    @staticmethod
    def forward(ctx, x, weight, bias, stride, padding, dilation, groups):
        ctx.save_for_backward(x, weight)
        ctx.conf = (stride, padding, dilation, groups, bias is not None)
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            return torch.nn.functional.conv2d(x, weight, bias, stride, padding, dilation, groups)

    @staticmethod
    def backward(ctx, gout):
        x, weight = ctx.saved_tensors
        stride, padding, dilation, groups, has_bias = ctx.conf
        mask = [ctx.needs_input_grad[0], ctx.needs_input_grad[1], has_bias and ctx.needs_input_grad[2]]
        with torch.backends.cudnn.flags(enabled=True, allow_tf32=False):
            gx, gw, gb = torch.ops.aten.convolution_backward(
                gout.contiguous(), x, weight, [weight.shape[0]] if has_bias else None, list(stride),
                list(padding), list(dilation), False, [0, 0], groups, mask)
        return gx, gw, gb, None, None, None, None


def _conv2d_fp32(conv, x):
    return _Fp32Conv2d.apply(x, conv.weight, conv.bias, conv.stride, conv.padding, conv.dilation, conv.groups)


class ConvE(ProjectionModel):
    """pykg2vec/models/projection.py:12-125 — same kwargs, sub-module names (state_dict keys
    ent_embeddings / rel_embeddings / b / bn0 / conv2d_1 / bn1 / fc / bn2), parameter_list and
    loss binding."""

    def __init__(self, **kwargs):
        super(ConvE, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "hidden_size", "hidden_size_1",
                      "lmbda", "input_dropout", "feature_map_dropout", "hidden_dropout"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.hidden_size_2 = self.hidden_size // self.hidden_size_1
        k = self.hidden_size
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, k)
        # reciprocal relations: every relation has a mirrored reverse (projection.py:40-42)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation * 2, k)
        self.b = NamedEmbedding("b", 1, self.tot_entity)
        self.bn0 = nn.BatchNorm2d(1)
        self.inp_drop = nn.Dropout(self.input_dropout)
        self.conv2d_1 = nn.Conv2d(1, 32, (3, 3), stride=(1, 1))
        self.bn1 = nn.BatchNorm2d(32)
        self.feat_drop = nn.Dropout2d(self.feature_map_dropout)
        self.fc = nn.Linear((2 * self.hidden_size_2 - 3 + 1) * (self.hidden_size_1 - 3 + 1) * 32, k)
        self.hidden_drop = nn.Dropout(self.hidden_dropout)
        self.bn2 = nn.BatchNorm1d(k)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings, self.b]
        self.loss = Criterion.multi_class_bce

    def embed(self, h, r, t):
        return self.ent_embeddings(h), self.rel_embeddings(r), self.ent_embeddings(t)

    def embed2(self, e, r):
        return self.ent_embeddings(e), self.rel_embeddings(r)

    # ---- trunk: everything of inner_forward before the x . E^T product (projection.py:88-99) ----
    def _trunk_layers(self, e, r):
        e_emb, r_emb = self.embed2(e, r)
        stacked_e = e_emb.view(-1, 1, self.hidden_size_2, self.hidden_size_1)
        stacked_r = r_emb.view(-1, 1, self.hidden_size_2, self.hidden_size_1)
        x = torch.cat([stacked_e, stacked_r], 2)
        x = self.bn0(x)
        x = self.inp_drop(x)
        x = _conv2d_fp32(self.conv2d_1, x)
        x = self.bn1(x)
        x = torch.relu(x)
        x = self.feat_drop(x)
        x = x.view(e.shape[0], -1)
        x = self.fc(x)
        x = self.hidden_drop(x)
        if self.training:
            x = self.bn2(x)   # the reference applies bn2 in training mode only (projection.py:97-98)
        return torch.relu(x)

    def _rel_ids(self, r, direction):
        assert direction in ("head", "tail"), "Unknown forward direction"
        return r + self.tot_relation if direction == "head" else r

    def proj_query(self, e, r, direction="tail"):
        """x [b, k]: the operand of the tail product, for the batched Evaluator (eval mode)."""
        _require_cuda(e, r, self.ent_embeddings.weight)
        r = self._rel_ids(r, direction)
        if not self.training and not torch.is_grad_enabled():
            return _lib.conve_trunk_fwd(self, e.contiguous(), r.contiguous())
        return self._trunk_layers(e, r)

    def proj_tail_tables(self):
        """(entity table [N,k], bias row [N]) of the tail product."""
        return self.ent_embeddings.weight, self.b.weight.view(-1)

    def forward(self, e, r, direction="tail"):
        x = self._trunk_layers(e, self._rel_ids(r, direction)) if (self.training or torch.is_grad_enabled()) \
            else self.proj_query(e, r, direction)
        return ProjTailFunction.apply(x, self.ent_embeddings.weight, self.b.weight)

    def predict_tail_rank(self, e, r, topk=-1):
        _, rank = torch.topk(-self.forward(e, r, direction="tail"), k=topk)
        return rank

    def predict_head_rank(self, e, r, topk=-1):
        _, rank = torch.topk(-self.forward(e, r, direction="head"), k=topk)
        return rank


class TuckER(ProjectionModel):
    """pykg2vec/models/projection.py:258-345 — same kwargs, tables (ent_embeddings, rel_embeddings, W),
    parameter_list and loss binding.  The trunk (normalise, contract the relation with the core tensor,
    multiply, normalise) is a handful of small dense ops and stays on torch; the product against every
    entity, its loss and the evaluation ranks are the kge_proj_* kernels (no bias row).  As in the
    reference the `direction` argument is validated and otherwise ignored: both directions apply the
    same function, to (h, r) and to (t, r)."""

    def __init__(self, **kwargs):
        super(TuckER, self).__init__(self.__class__.__name__.lower())
        param_list = ["tot_entity", "tot_relation", "ent_hidden_size", "rel_hidden_size", "lmbda",
                      "input_dropout", "hidden_dropout1", "hidden_dropout2"]
        param_dict = self.load_params(param_list, kwargs)
        self.__dict__.update(param_dict)
        self.d1, self.d2 = self.ent_hidden_size, self.rel_hidden_size
        self.ent_embeddings = NamedEmbedding("ent_embedding", self.tot_entity, self.d1)
        self.rel_embeddings = NamedEmbedding("rel_embedding", self.tot_relation, self.d2)
        self.W = NamedEmbedding("W", self.d2, self.d1 * self.d1)
        nn.init.xavier_uniform_(self.ent_embeddings.weight)
        nn.init.xavier_uniform_(self.rel_embeddings.weight)
        nn.init.xavier_uniform_(self.W.weight)
        self.parameter_list = [self.ent_embeddings, self.rel_embeddings, self.W]
        self.inp_drop = nn.Dropout(self.input_dropout)
        self.hidden_dropout1 = nn.Dropout(self.hidden_dropout1)
        self.hidden_dropout2 = nn.Dropout(self.hidden_dropout2)
        self.loss = Criterion.multi_class_bce

    def proj_query(self, e1, r, direction="tail"):
        assert direction in ("head", "tail"), "Unknown forward direction"
        _require_cuda(e1, r, self.ent_embeddings.weight)
        e1 = torch.nn.functional.normalize(self.ent_embeddings(e1), p=2, dim=1)
        e1 = self.inp_drop(e1).view(-1, 1, self.d1)
        W_mat = torch.matmul(self.rel_embeddings(r), self.W.weight.view(self.d2, -1)).view(-1, self.d1, self.d1)
        W_mat = self.hidden_dropout1(W_mat)
        x = torch.matmul(e1, W_mat).view(-1, self.d1)
        return self.hidden_dropout2(torch.nn.functional.normalize(x, p=2, dim=1))

    def proj_tail_tables(self):
        return self.ent_embeddings.weight, None

    def forward(self, e1, r, direction="head"):
        return ProjTailFunction.apply(self.proj_query(e1, r, direction), self.ent_embeddings.weight, None)

    def predict_tail_rank(self, e, r, topk=-1):
        _, rank = torch.topk(-self.forward(e, r, direction="tail"), k=topk)
        return rank

    def predict_head_rank(self, e, r, topk=-1):
        _, rank = torch.topk(-self.forward(e, r, direction="head"), k=topk)
        return rank
