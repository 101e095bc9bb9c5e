"""Base classes of the model surface — mirror of pykg2vec/models/KGMeta.py:14-80.

`training_strategy` uses the reference's own `TrainingStrategy` enum when pykg2vec is
importable (its Trainer compares members by identity, pykg2vec/utils/trainer.py:274-296),
otherwise an identical local enum (pykg2vec/common.py:19-24).
"""
from abc import ABCMeta
from enum import Enum

import torch.nn as nn

try:  # running inside a pykg2vec installation: share its enum so Trainer dispatch works
    from pykg2vec.common import TrainingStrategy  # type: ignore
except Exception:  # standalone (e.g. the GPU box)
    class TrainingStrategy(Enum):
        PROJECTION_BASED = "projection_based"  # matching models with neural network
        PAIRWISE_BASED = "pairwise_based"      # translational distance models
        POINTWISE_BASED = "pointwise_based"    # semantic matching models

        # If pykg2vec becomes importable only AFTER this module was loaded, its Trainer / Generator compare
        # `model.training_strategy == pykg2vec.common.TrainingStrategy.X` (trainer.py:274-296,
        # generator.py:300-308): members of the two enums with the same value must then compare equal.
        def __eq__(self, other):
            return isinstance(other, Enum) and type(other).__name__ == "TrainingStrategy" and other.value == self.value

        def __hash__(self):
            return hash(self.value)


class Model:
    """Meta class of KGE models (KGMeta.py:14-38)."""

    def __init__(self):
        self.database = None

    def embed(self, h, r, t):
        raise NotImplementedError

    def forward(self, h, r, t):
        raise NotImplementedError

    def load_params(self, param_list, kwargs):
        for param_name in param_list:
            if param_name not in kwargs:
                raise Exception("hyperparameter %s not found!" % param_name)
            self.database[param_name] = kwargs[param_name]
        return self.database

    def get_reg(self, h, r, t, **kwargs):
        return 0.0


class PairwiseModel(nn.Module, Model):
    __metaclass__ = ABCMeta

    def __init__(self, model_name):
        super(PairwiseModel, self).__init__()
        self.model_name = model_name
        self.training_strategy = TrainingStrategy.PAIRWISE_BASED
        self.database = {}


class PointwiseModel(nn.Module, Model):
    __metaclass__ = ABCMeta

    def __init__(self, model_name):
        super(PointwiseModel, self).__init__()
        self.model_name = model_name
        self.training_strategy = TrainingStrategy.POINTWISE_BASED
        self.database = {}


class ProjectionModel(nn.Module, Model):
    """Meta class of the models with a neural projection trunk (KGMeta.py:67-80)."""
    __metaclass__ = ABCMeta

    def __init__(self, model_name):
        super(ProjectionModel, self).__init__()
        self.model_name = model_name
        self.training_strategy = TrainingStrategy.PROJECTION_BASED
        self.database = {}
