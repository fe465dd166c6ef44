"""HitAtK / Precision / Recall / MAP / NDCG @k meters (torchok/metrics/representation_ranx.py:17-121).

The reference formats per-query Python lists for ranx 0.3.8 (numba) — `process_data_for_metric_func` :28-51.  Here a
metric is a function of device tensors:
    gains (nq, k)  gain of each retrieved item for its query (0 = not relevant), in rank order
    n_rel (nq,)    number of relevant items of the query in the whole gallery
    ideal (nq, k)  the k largest gains of the query (None = binary relevance)
    k              cut-off
returning one value per query; formulas are ranx's (pinned by the reference's golden answers, see
tests/golden/retrieval_kat.json): precision divides by k, recall and average precision by ALL relevants.
"""
import torch

from ..constructor import METRICS
from .index_base_metric import IndexBasedMeter


def hit_rate(gains, n_rel, ideal, k):
    return (gains[:, :k] > 0).any(dim=1).float()


def precision(gains, n_rel, ideal, k):
    return (gains[:, :k] > 0).float().sum(dim=1) / k


def recall(gains, n_rel, ideal, k):
    return (gains[:, :k] > 0).float().sum(dim=1) / n_rel.clamp_min(1).float()


def average_precision(gains, n_rel, ideal, k):
    hits = (gains[:, :k] > 0).float()
    ranks = torch.arange(1, hits.shape[1] + 1, device=hits.device, dtype=torch.float32)
    prec_at = hits.cumsum(dim=1) / ranks
    return (prec_at * hits).sum(dim=1) / n_rel.clamp_min(1).float()


def ndcg(gains, n_rel, ideal, k):
    g = gains[:, :k].float()
    disc = 1.0 / torch.log2(torch.arange(2, g.shape[1] + 2, device=g.device, dtype=torch.float32))
    dcg = (g * disc).sum(dim=1)
    if ideal is None:  # binary relevance: the ideal list is min(n_rel, k) ones
        ranks = torch.arange(g.shape[1], device=g.device).unsqueeze(0)
        ideal = (ranks < n_rel.unsqueeze(1)).float()
    idcg = (ideal[:, :k].float() * disc[:ideal.shape[1]]).sum(dim=1)
    return torch.where(idcg > 0, dcg / idcg.clamp_min(1e-30), torch.zeros_like(dcg))


class RanxBasedMeter(IndexBasedMeter):
    metric = None

    def __init__(self, dataset_type, exact_index=True, metric_distance='IP', k=None, search_batch_size=None,
                 normalize_vectors=False, group_averaging=False, k_as_target_len=False, use_batching_search=True,
                 raise_empty_query=True, **kwargs):
        super().__init__(exact_index=exact_index, dataset_type=dataset_type, metric_distance=metric_distance,
                         metric_func=type(self).metric, k=k, search_batch_size=search_batch_size,
                         normalize_vectors=normalize_vectors, group_averaging=group_averaging,
                         k_as_target_len=k_as_target_len, use_batching_search=use_batching_search,
                         raise_empty_query=raise_empty_query, **kwargs)


@METRICS.register_class
class HitAtKMeter(RanxBasedMeter):
    metric = staticmethod(hit_rate)


@METRICS.register_class
class PrecisionAtKMeter(RanxBasedMeter):
    metric = staticmethod(precision)


@METRICS.register_class
class RecallAtKMeter(RanxBasedMeter):
    metric = staticmethod(recall)


@METRICS.register_class
class MeanAveragePrecisionAtKMeter(RanxBasedMeter):
    metric = staticmethod(average_precision)


@METRICS.register_class
class NDCGAtKMeter(RanxBasedMeter):
    metric = staticmethod(ndcg)
