"""Host-side mirror of the reference's workflow API surface that the hot path is reached through.

Reference (K/ = src/main/scala/keystoneml/):
  Transformer        K/workflow/Transformer.scala:18-55     apply(datum) / apply(batch)
  Estimator          K/workflow/Estimator.scala:10-62       fit(data) -> Transformer
  LabelEstimator     K/workflow/LabelEstimator.scala:13-100 fit(data, labels) -> Transformer
  Chainable.andThen  K/workflow/Chainable.scala:26-125      (next) | (est, data) | (est, data, labels)
  Pipeline.gather    K/workflow/Pipeline.scala:119-154      branches applied to the same input -> Seq
  WeightedNode       K/workflow/WeightedNode.scala:7-9      weight = number of passes over the input

Only the API shape is kept (same names and argument meaning) so that pipelines written against the
reference read the same; the DAG optimizer / auto-cache rules are out of scope.
Estimators chained with ``andThen(est, data[, labels])`` are fitted lazily, exactly once
(T/workflow/PipelineSuite.scala:28 "Do not fit estimators multiple times").
"""
from __future__ import annotations

from typing import Any, Callable, List, Optional, Sequence


class Chainable:
    def to_pipeline(self) -> "Pipeline":
        raise NotImplementedError

    def andThen(self, nxt, data=None, labels=None) -> "Pipeline":
        """Chainable.scala:26-31 (next), :44-50 (estimator + data), :67-73 (label estimator + data + labels)."""
        me = self.to_pipeline()
        if isinstance(nxt, LabelEstimator):
            if data is None or labels is None:
                raise TypeError("andThen(LabelEstimator, data, labels) needs data and labels")
            return Pipeline(me.stages + [_LazyFit(nxt, me, data, labels)])
        if isinstance(nxt, Estimator):
            if data is None:
                raise TypeError("andThen(Estimator, data) needs data")
            return Pipeline(me.stages + [_LazyFit(nxt, me, data, None)])
        if isinstance(nxt, Chainable):
            return Pipeline(me.stages + nxt.to_pipeline().stages)
        raise TypeError(f"cannot chain {type(nxt).__name__}")

    # Scala-style alias
    and_then = andThen


class Transformer(Chainable):
    """apply() accepts one datum (1-D numpy vector) or a batch (Dataset / 2-D numpy array)."""

    def apply(self, data):
        raise NotImplementedError

    def __call__(self, data):
        return self.apply(data)

    def to_pipeline(self) -> "Pipeline":
        return Pipeline([self])


class Estimator:
    def fit(self, data) -> Transformer:
        raise NotImplementedError

    def withData(self, data) -> "Pipeline":  # Estimator.scala:33-61
        return Pipeline([_LazyFit(self, Pipeline([]), data, None)])


class LabelEstimator:
    def fit(self, data, labels) -> Transformer:
        raise NotImplementedError

    def withData(self, data, labels) -> "Pipeline":  # LabelEstimator.scala:58-82
        return Pipeline([_LazyFit(self, Pipeline([]), data, labels)])


class WeightedNode:
    weight: int = 1


class _LazyFit:
    """An estimator grafted into a pipeline together with its training data; fitted on first use."""

    def __init__(self, est, prefix: "Pipeline", data, labels):
        self.est, self.prefix, self.data, self.labels = est, prefix, data, labels
        self.fitted: Optional[Transformer] = None
        self.fit_count = 0

    def get(self) -> Transformer:
        if self.fitted is None:
            feats = self.prefix.apply(self.data)
            self.fitted = self.est.fit(feats) if self.labels is None else self.est.fit(feats, self.labels)
            self.fit_count += 1
        return self.fitted


class Pipeline(Transformer):
    def __init__(self, stages: Sequence[Any]):
        self.stages: List[Any] = list(stages)

    def to_pipeline(self) -> "Pipeline":
        return self

    def apply(self, data):
        for s in self.stages:
            t = s.get() if isinstance(s, _LazyFit) else s
            data = t.apply(data)
        return data

    def fit(self) -> "Pipeline":
        """Pipeline.fit (Pipeline.scala:38-62): force every estimator, return a transformer-only pipeline."""
        return Pipeline([s.get() if isinstance(s, _LazyFit) else s for s in self.stages])

    @staticmethod
    def gather(branches: Sequence[Chainable]) -> "Pipeline":
        """Pipeline.gather (Pipeline.scala:119-154): apply every branch to the same input, emit the list."""
        return Pipeline([_Gather([b.to_pipeline() for b in branches])])


class _Gather(Transformer):
    def __init__(self, branches: Sequence[Pipeline]):
        self.branches = list(branches)

    @staticmethod
    def _context_of(pipe: "Pipeline"):
        for s in pipe.stages:
            ctx = getattr(s, "ctx", None)
            if ctx is not None:
                return ctx
        return None

    def apply(self, data):
        # a host batch is uploaded ONCE and every branch sees the same device matrix (gathered lazy feature maps must
        # share their input; the reference's branches all read the same RDD, Pipeline.scala:119-154)
        import numpy as np
        if isinstance(data, np.ndarray) and data.ndim == 2:
            ctx = next((c for c in (self._context_of(b) for b in self.branches) if c is not None), None)
            if ctx is not None:
                data = ctx.matrix(data)
        return [b.apply(data) for b in self.branches]


class FunctionNode(Transformer):
    def __init__(self, fn: Callable):
        self.fn = fn

    def apply(self, data):
        return self.fn(data)
