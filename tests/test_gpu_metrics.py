"""mgcn_binary_confusion and the metrics module (meta_gcn_b200/gcn_meta/optim/metrics.py) against the values of
the UNMODIFIED reference src/gcn_meta/optim/metrics.py (tests/golden/metrics.npz): counters bit-exact,
derived ratios identical, the same ZeroDivisionError / -1 / 0 behaviour."""
import os

import numpy as np
import pytest
import torch

from meta_gcn_b200.gcn_meta.optim import metrics as M
from oracle import port

pytestmark = pytest.mark.gpu
DEV = "cuda"
G = np.load(os.path.join(os.path.dirname(__file__), "golden", "metrics.npz"))
NAMES = ["accuracy", "true_positive", "false_positive", "true_negative", "false_negative", "recall",
         "precision", "f1_score", "false_positive_rate", "false_negative_rate"]


@pytest.mark.parametrize("k", [0, 1, 2])
def test_metrics_match_reference_golden(k):
    logits = torch.from_numpy(G[f"logits{k}"]).to(DEV)
    y = torch.from_numpy(G[f"y{k}"]).long().to(DEV)
    want = G[f"vals{k}"]
    pred = logits.argmax(1)
    c_pred, c_logits = M.confusion(pred, y), M.confusion_from_logits(logits, y)
    assert c_pred == c_logits
    assert [c_pred.correct / c_pred.numel, c_pred.tp, c_pred.fp, c_pred.tn, c_pred.fn] == list(want[:5])
    for i, name in enumerate(NAMES):
        fn = getattr(M, name)
        if np.isnan(want[i]):
            with pytest.raises(ZeroDivisionError):
                fn(pred, y)
        else:
            assert fn(pred, y) == want[i], name


def test_confusion_large_and_ties():
    n = 3_000_001
    g = torch.Generator().manual_seed(0)
    logits = torch.randn(n, 2, generator=g)
    logits[::7, 1] = logits[::7, 0]                      # ties: the first maximal class (0) wins
    y = (torch.rand(n, generator=g) < 0.1).long()
    c = M.confusion_from_logits(logits.to(DEV), y.to(DEV))
    pred = (logits[:, 1] > logits[:, 0]).long()
    want = port.binary_metrics(pred.numpy(), y.numpy())
    assert [c.tp, c.fp, c.tn, c.fn] == [int(v) for v in want[1:5]]
    assert c.tp + c.fp + c.tn + c.fn == n and c.correct == c.tp + c.tn
    empty = M.confusion(torch.zeros(0, dtype=torch.long, device=DEV), torch.zeros(0, dtype=torch.long, device=DEV))
    assert empty[:5] == (0, 0, 0, 0, 0)
