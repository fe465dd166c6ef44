"""Retrieval metrics of the hot path: IndexBasedMeter and the ranx-backed meters under their registry names."""
from .index_base_metric import IndexBasedMeter, search_topk  # noqa: F401
from .representation_ranx import (HitAtKMeter, MeanAveragePrecisionAtKMeter, NDCGAtKMeter, PrecisionAtKMeter,  # noqa: F401
                                  RecallAtKMeter)
from .metrics_manager import (Accuracy, F1Score, JaccardIndex, Metric, MetricsManager, MetricWithUtils, Phase,  # noqa: F401
                              Precision, Recall)
