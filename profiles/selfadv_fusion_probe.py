"""RotatE training step: fused forward+self-adversarial loss+backward kernel (kge_train_pairwise_selfadv) vs the
five-launch path (2 x kge_score_fwd, kge_loss_selfadv, 2 x kge_score_bwd), per neg_rate.  CUDA events, L2 flushed.
Usage (GPU box): python profiles/selfadv_fusion_probe.py > gpurun_out/r2_selfadv_fusion.jsonl"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pykg2vec_b200  # noqa: E402
from pykg2vec_b200.synthetic import SyntheticConfig, SyntheticKnowledgeGraph  # noqa: E402
from pykg2vec_b200.trainer import Trainer  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    N, R = 14951, 1345
    for d in (200, 1000):
        for B, neg in ((512, 1), (512, 4), (512, 16), (512, 64), (1024, 256)):
            kg = SyntheticKnowledgeGraph(N, R, 4096, 64, 64, seed=0)
            cfg = SyntheticConfig(kg, device=dev, optimizer="adagrad", learning_rate=0.01, batch_size=B, neg_rate=neg,
                                  hidden_size=d, margin=6.0, alpha=0.5)
            torch.manual_seed(0)
            tr = Trainer(pykg2vec_b200.import_model("rotate")(**cfg.__dict__), cfg)
            tr.build_model()
            rng = np.random.RandomState(1)
            to = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.int64)).to(dev)
            ids = [to(rng.randint(N, size=B)), to(rng.randint(R, size=B)), to(rng.randint(N, size=B)),
                   to(rng.randint(N, size=B * neg)), to(rng.randint(R, size=B * neg)), to(rng.randint(N, size=B * neg))]
            out = {"d": d, "B": B, "neg": neg}
            for fused in (True, False):
                tr._selfadv_fused = fused
                ts = []
                with torch.no_grad():
                    for it in range(8):
                        flush.zero_()
                        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                        e0.record()
                        tr._fused_pairwise(ids)
                        e1.record()
                        torch.cuda.synchronize()
                        ts.append(e0.elapsed_time(e1))
                out["fused_ms" if fused else "unfused_ms"] = float(np.median(ts[3:]))
            print(json.dumps(out), flush=True)
            del tr
            torch.cuda.empty_cache()


if __name__ == "__main__":
    main()
