"""Dataset-shaped synthetic knowledge graphs (no dataset can be downloaded offline:
pykg2vec/data/datasets.py:127-139 fetches by URL).  Exposes the slice of the
KnowledgeGraph cache interface the hot path consumes
(pykg2vec/data/kgcontroller.py:258-330: read_cache_data('triplets_*' | 'hr_t' | 'tr_h'))."""
from collections import namedtuple

import numpy as np

Triple = namedtuple("Triple", ["h", "r", "t"])

# public dataset statistics: name -> (entities, relations, train, valid, test)
SHAPES = {
    "umls": (135, 46, 5216, 652, 661),
    "fb15k_237": (14541, 237, 272115, 17535, 20466),
    "wn18rr": (40943, 11, 86835, 3034, 3134),
    "fb15k": (14951, 1345, 483142, 50000, 59071),
    "yago3_10": (123182, 37, 1079040, 5000, 5000),
}


class SyntheticKnowledgeGraph:
    def __init__(self, tot_entity, tot_relation, n_train, n_valid, n_test, seed=0, name="synthetic"):
        rng = np.random.RandomState(seed)
        self.dataset_name = name
        self.tot_entity, self.tot_relation = tot_entity, tot_relation
        total = n_train + n_valid + n_test
        arr = np.stack([rng.randint(tot_entity, size=total), rng.randint(tot_relation, size=total),
                        rng.randint(tot_entity, size=total)], axis=1).astype(np.int64)
        self.arrays = {"train": arr[:n_train], "valid": arr[n_train:n_train + n_valid],
                       "test": arr[n_train + n_valid:]}
        self._cache = {}

    @classmethod
    def shaped_like(cls, dataset, seed=0, scale=1.0):
        n, r, tr, va, te = SHAPES[dataset]
        return cls(n, r, max(1, int(tr * scale)), max(1, int(va * scale)), max(1, int(te * scale)),
                   seed=seed, name=dataset + "-shaped-synthetic")

    def _triples(self, split):
        return [Triple(int(h), int(r), int(t)) for h, r, t in self.arrays[split]]

    def read_cache_data(self, key):
        if key in self._cache:
            return self._cache[key]
        if key in ("triplets_train", "triplets_valid", "triplets_test"):
            val = self._triples(key.split("_")[1])
        elif key in ("hr_t", "tr_h"):
            # all splits, as kgcontroller.py:410-428
            hr_t, tr_h = {}, {}
            for split in ("train", "valid", "test"):
                for h, r, t in self.arrays[split]:
                    hr_t.setdefault((int(h), int(r)), set()).add(int(t))
                    tr_h.setdefault((int(t), int(r)), set()).add(int(h))
            self._cache["hr_t"], self._cache["tr_h"] = hr_t, tr_h
            return self._cache[key]
        elif key in ("hr_t_train", "tr_h_train"):
            # training split only, as kgcontroller.py builds them for the multi-class label rows
            hr_t, tr_h = {}, {}
            for h, r, t in self.arrays["train"]:
                hr_t.setdefault((int(h), int(r)), set()).add(int(t))
                tr_h.setdefault((int(t), int(r)), set()).add(int(h))
            self._cache["hr_t_train"], self._cache["tr_h_train"] = hr_t, tr_h
            return self._cache[key]
        else:
            raise ValueError("Unknown cache data key %s" % key)
        self._cache[key] = val
        return val


class SyntheticConfig:
    """The attributes of pykg2vec.config.Config the hot path reads (config.py:46-83)."""

    def __init__(self, kg, device="cuda", **kw):
        self.knowledge_graph = kg
        self.dataset_name = kg.dataset_name
        self.tot_entity, self.tot_relation = kg.tot_entity, kg.tot_relation
        self.tot_train_triples = len(kg.arrays["train"])
        self.tot_valid_triples = len(kg.arrays["valid"])
        self.tot_test_triples = len(kg.arrays["test"])
        self.device = device
        self.hits = [1, 3, 5, 10]
        self.test_num = 1000
        self.debug = False
        self.batch_size = 128
        self.neg_rate = 1
        self.margin = 0.8
        self.alpha = 0.1
        self.learning_rate = 0.01
        self.optimizer = "adam"
        self.l1_flag = True
        self.hidden_size = 50
        self.lmbda = 0.1
        self.__dict__.update(kw)
