"""Criterion — mirror of pykg2vec/utils/criterion.py with the same static-method
signatures; each loss is one CUDA kernel producing value and d loss/d score
(kge_loss_* in include/kge_b200.h)."""
from . import functional as F_


class Criterion:
    @staticmethod
    def pariwise_logistic(pos_preds, neg_preds, neg_rate, alpha):
        """RotatE self-adversarial negative sampling loss (criterion.py:14-23; sic)."""
        return F_.SelfAdvFunction.apply(pos_preds, neg_preds, neg_rate, alpha)

    @staticmethod
    def pairwise_hinge(pos_preds, neg_preds, margin):
        """criterion.py:26-29: sum_i max(pos_i + margin - neg_i, 0)."""
        return F_.HingeFunction.apply(pos_preds, neg_preds, margin)

    @staticmethod
    def pointwise_logistic(preds, target):
        """criterion.py:32-34: mean_i softplus(target_i * preds_i)."""
        return F_.LogisticFunction.apply(preds, target)
