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

    @staticmethod
    def multi_class_bce(pred_heads, pred_tails, tr_h, hr_t, label_smoothing, tot_entity):
        """criterion.py:41-50: BCEWithLogits (applied to the sigmoided preds, as the reference does)
        against the optionally smoothed dense labels, head direction + tail direction."""
        scale, shift = 1.0, 0.0
        if label_smoothing is not None and tot_entity is not None:
            scale, shift = 1.0 - label_smoothing, 1.0 / tot_entity
        loss_heads = F_.MultiClassBceFunction.apply(pred_heads, tr_h, scale, shift)
        loss_tails = F_.MultiClassBceFunction.apply(pred_tails, hr_t, scale, shift)
        return loss_heads + loss_tails

    @staticmethod
    def multi_class(pred_heads, pred_tails):
        """criterion.py:52-55 (ProjE: the model's forward already returns the loss terms)."""
        return pred_heads + pred_tails
