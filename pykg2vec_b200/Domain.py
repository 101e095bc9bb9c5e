"""NamedEmbedding — mirror of pykg2vec/models/Domain.py:8-17 (an nn.Embedding with a name;
dense gradients, fp32 row-major weight: the layout the CUDA kernels read)."""
from torch.nn import Embedding


class NamedEmbedding(Embedding):
    def __init__(self, name, *args, **kwargs):
        super(NamedEmbedding, self).__init__(*args, **kwargs)
        self._name = name

    @property
    def name(self):
        return self._name
