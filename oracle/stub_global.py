"""TEST INFRASTRUCTURE ONLY -- a deterministic stand-in for the reference's global model (global_model.py:81-92).

The global model is outside this repo's scope; the test-time path only calls ``predict(t, graph_dict, subject)`` and
uses the returned ``(embedding [1,1,h], logits [1,1,num_e], prob [num_e])``.  The stub derives all three from a seed and
the timestamp alone, so the reference and the CUDA path see identical distributions (CPU tensors: the sampling RNG
stream is then the same on both sides)."""
import torch


class StubGlobalModel:
    def __init__(self, num_e, h, seed):
        self.num_e, self.h, self.seed = num_e, h, seed
        self.calls = []

    def predict(self, t, graph_dict, subject=True):
        self.calls.append((int(t), bool(subject)))
        g = torch.Generator().manual_seed(self.seed + 7 * int(t) + (0 if subject else 100003))
        emb = 0.1 * torch.randn(1, 1, self.h, generator=g)
        logits = 2.0 * torch.randn(1, 1, self.num_e, generator=g)
        return emb, logits, torch.softmax(logits.view(-1), dim=0)
