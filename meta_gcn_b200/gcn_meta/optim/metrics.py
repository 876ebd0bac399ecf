"""Evaluation metrics for binary classification — the names, arguments and error behaviour of
src/gcn_meta/optim/metrics.py:8-60 (`pred`, `target`: LongTensors of one length), computed by ONE pass of
mgcn_binary_confusion instead of a boolean-mask pass and a host synchronisation per counter.

`confusion(pred, target)` / `confusion_from_logits(logits, target)` return the five counters as a
`Confusion`; the train loop of train_botnet.py:296-305 needs one of them per report instead of ~14 syncs:

    c = confusion_from_logits(x, batch.y.long())        # argmax fused, one D2H of 5 integers
    acc, fpr, fnr, rec, prc, f1 = c.accuracy(), c.false_positive_rate(), ...
"""
from collections import namedtuple

from ... import ops


class Confusion(namedtuple("Confusion", "tp fp tn fn correct numel")):
    """the counters of metrics.py:12-24 plus what accuracy needs; the derived values follow
    metrics.py:8-60 including its division-by-zero behaviour"""
    __slots__ = ()

    def accuracy(self):
        return self.correct / self.numel                           # metrics.py:9

    def recall(self):
        return self.tp / (self.tp + self.fn)                       # metrics.py:31  (target == 1).sum()

    def precision(self):
        try:
            return self.tp / (self.tp + self.fp)                   # metrics.py:36  (pred == 1).sum()
        except ZeroDivisionError:
            return -1

    def f1_score(self):
        prec, rec = self.precision(), self.recall()
        try:
            return 2 * (prec * rec) / (prec + rec)                 # metrics.py:46
        except ZeroDivisionError:
            return 0

    def false_positive_rate(self):
        return self.fp / (self.fp + self.tn)                       # metrics.py:52  (target == 0).sum()

    def false_negative_rate(self):
        return self.fn / (self.tp + self.fn)                       # metrics.py:59


def _counts(t, numel):
    tp, fp, tn, fn, correct = (int(v) for v in t.tolist())         # the one host synchronisation
    return Confusion(tp, fp, tn, fn, correct, numel)


def confusion(pred, target):
    return _counts(ops.binary_confusion_impl(target, pred=pred), target.numel())


def confusion_from_logits(logits, target):
    """pred = logits.argmax(1) fused into the counting pass (train_botnet.py:297)"""
    return _counts(ops.binary_confusion_impl(target, logits=logits), target.numel())


def accuracy(pred, target):
    return confusion(pred, target).accuracy()


def true_positive(pred, target):
    return confusion(pred, target).tp


def false_positive(pred, target):
    return confusion(pred, target).fp


def true_negative(pred, target):
    return confusion(pred, target).tn


def false_negative(pred, target):
    return confusion(pred, target).fn


def recall(pred, target):
    """Or true positive rate."""
    return confusion(pred, target).recall()


def precision(pred, target):
    return confusion(pred, target).precision()


def f1_score(pred, target):
    return confusion(pred, target).f1_score()


def false_positive_rate(pred, target):
    return confusion(pred, target).false_positive_rate()


def false_negative_rate(pred, target):
    """Or 1 - recall/true_positive_rate"""
    return confusion(pred, target).false_negative_rate()
