"""numpy CPU restatement of the reference's retrieval metric path.  TEST INFRASTRUCTURE (see oracle/__init__.py).

PINNED: reproduces every known answer of the reference's own tests
(tests/base_tests/metrics/representation/data.py:123-148 classification, :197-230 representation incl. NDCG,
:312-329 query-as-relevant) — checked by tests/test_oracle_retrieval.py against tests/golden/retrieval_kat.json.

Follows torchok/metrics/index_base_metric.py:
    compute .......................... :170-270   -> IndexBasedMeterOracle.compute
    prepare_representation_data ...... :297-373   -> prepare_representation_data
    prepare_classification_data ...... :378-416   -> prepare_classification_data
    clear_faiss_output ............... :418-442   -> drop_self_or_last
    query_generator .................. :444-521   -> (loop inside compute)
    build_index / IndexFlatIP|L2 ..... :523-545   -> flat_search   (faiss-cpu 1.7.2, not vendored: exact brute force,
                                                     IP = descending inner product, L2 = ascending squared distance,
                                                     ties -> lower index first)
and torchok/metrics/representation_ranx.py:28-51 + ranx 0.3.8 metrics (not vendored; formulas pinned by the goldens):
    hit_rate, precision, recall, average_precision, ndcg.

Deviation kept on purpose: the reference normalises with np.linalg.norm(vectors, axis=0) (per column,
index_base_metric.py:192-193) but its golden answers are only reproduced by per-row L2 (true cosine); `normalize`
follows the goldens, `normalize_axis0=True` gives the literal behaviour.
"""
import math

import numpy as np


# ------------------------------------------------------------------------------------------------ search
def flat_search(gallery, queries, k, metric='IP'):
    """faiss.IndexFlatIP / IndexFlatL2 .search(queries, k) on an exact index: (scores, idx), both (nq, k).
    Fewer than k gallery rows -> padded with idx -1 and score -inf (IP) / +inf (L2), like faiss."""
    gallery = np.asarray(gallery, dtype=np.float32)
    queries = np.asarray(queries, dtype=np.float32)
    nq, ng = len(queries), len(gallery)
    if metric == 'IP':
        s = queries @ gallery.T
        key = -s
    elif metric == 'L2':
        s = ((queries[:, None, :] - gallery[None, :, :]) ** 2).sum(-1) if nq * ng <= 1 << 22 else \
            (queries ** 2).sum(1)[:, None] + (gallery ** 2).sum(1)[None, :] - 2.0 * (queries @ gallery.T)
        key = s
    else:
        raise ValueError(f'unknown metric {metric}')
    order = np.argsort(key, axis=1, kind='stable')[:, :k]
    scores = np.take_along_axis(s, order, axis=1)
    if ng < k:
        pad = k - ng
        order = np.concatenate([order, -np.ones((nq, pad), dtype=order.dtype)], 1)
        scores = np.concatenate([scores, np.full((nq, pad), -np.inf if metric == 'IP' else np.inf, np.float32)], 1)
    return scores, order


def drop_self_or_last(out, query_in_gallery):
    """clear_faiss_output (:418-442): k+1 were searched; drop column 0 for queries that are in the index (assumed to
    be their own nearest neighbour), else drop the last column."""
    out = np.asarray(out)
    res = np.empty((out.shape[0], out.shape[1] - 1), dtype=np.float64)
    res[query_in_gallery] = out[query_in_gallery][:, 1:]
    res[~query_in_gallery] = out[~query_in_gallery][:, :-1]
    return res


# ------------------------------------------------------------------------------------------------ data preparation
def prepare_classification_data(targets, raise_empty_query=True):
    """Every vector is a query; relevants = other vectors with the same label (:378-416).  Query order follows the
    reference's pandas groupby: labels ascending, members in storage order."""
    targets = np.asarray(targets)
    relevant, query_rows = [], []
    for label in np.unique(targets):
        group = np.where(targets == label)[0]
        for qi in group:
            rel = [int(g) for g in group if g != qi]
            if not rel and raise_empty_query:
                raise ValueError(f'Representation metric. The class {label} has only one element.')
            query_rows.append(int(qi))
            relevant.append(rel)
    n = len(targets)
    return relevant, np.arange(n), np.asarray(query_rows), np.ones(n, dtype=bool)


def prepare_representation_data(query_idxs, scores, raise_empty_query=True):
    """(:297-373) queries are rows with query_idxs >= 0 (value = their column in `scores`); a query stays in the
    gallery only if it is relevant to some other query; relevants sorted by descending score."""
    query_idxs = np.asarray(query_idxs)
    scores = np.asarray(scores)
    is_query = query_idxs >= 0
    query_cols = query_idxs[is_query]
    query_rows = np.where(is_query)[0]
    query_in_gallery = np.any(scores[query_rows, :] > 0, axis=-1)
    gallery = np.delete(np.arange(len(scores)), query_rows[~query_in_gallery])
    relevant = []
    for col in query_cols:
        rel = np.where(scores[:, col] > 0.0)[0]
        if len(rel) == 0:
            if raise_empty_query:
                raise ValueError('Representation metric. The dataset contains a query vector that does not has '
                                 'relevants. Set parameter raise_empty_query to False for compute.')
            relevant.append([])
        else:
            order = np.argsort(scores[rel, col])[::-1]
            relevant.append([int(r) for r in rel[order]])
    return relevant, gallery, query_cols, query_rows, query_in_gallery


# ------------------------------------------------------------------------------------------------ ranx 0.3.8 metrics
# qrels: list of (doc_id, gain) for one query; run: ranked doc ids; k: cut-off (0 = all).
def _cut(run, k):
    return run if k == 0 else run[:k]


def hit_rate(qrels, run, k):
    rel = {d for d, g in qrels if g > 0}
    return 1.0 if any(d in rel for d in _cut(run, k)) else 0.0


def precision(qrels, run, k):
    rel = {d for d, g in qrels if g > 0}
    run = _cut(run, k)
    denom = k if k > 0 else len(run)
    return sum(d in rel for d in run) / denom if denom else 0.0


def recall(qrels, run, k):
    rel = {d for d, g in qrels if g > 0}
    return sum(d in rel for d in _cut(run, k)) / len(rel) if rel else 0.0


def average_precision(qrels, run, k):
    rel = {d for d, g in qrels if g > 0}
    if not rel:
        return 0.0
    hits, acc = 0, 0.0
    for i, d in enumerate(_cut(run, k), 1):
        if d in rel:
            hits += 1
            acc += hits / i
    return acc / len(rel)  # denominator = ALL relevants (data.py:116-121)


def ndcg(qrels, run, k):
    gain = {d: g for d, g in qrels}
    dcg = sum(gain.get(d, 0.0) / math.log2(i + 1) for i, d in enumerate(_cut(run, k), 1))
    ideal = sorted((g for g in gain.values()), reverse=True)
    ideal = ideal if k == 0 else ideal[:k]
    idcg = sum(g / math.log2(i + 1) for i, g in enumerate(ideal, 1))
    return dcg / idcg if idcg > 0 else 0.0


METRIC_FUNCS = {'hit_rate': hit_rate, 'precision': precision, 'recall': recall,
                'average_precision': average_precision, 'ndcg': ndcg}


# ------------------------------------------------------------------------------------------------ the meter
class IndexBasedMeterOracle:
    """IndexBasedMeter + RanxBasedMeter.process_data_for_metric_func, batch accumulation included."""

    def __init__(self, metric, dataset_type, k=None, metric_distance='IP', normalize_vectors=False,
                 group_averaging=False, k_as_target_len=False, search_batch_size=None, use_batching_search=True,
                 raise_empty_query=True, normalize_axis0=False, **unused):
        if dataset_type not in ('classification', 'representation'):
            raise ValueError(f'unknown dataset type {dataset_type}')
        if metric_distance not in ('IP', 'L2'):
            raise ValueError(f'unknown metric distance {metric_distance}')
        self.metric_func = METRIC_FUNCS[metric]
        self.dataset_type, self.metric_distance = dataset_type, metric_distance
        self.normalize_vectors, self.normalize_axis0 = normalize_vectors, normalize_axis0
        self.group_averaging, self.k_as_target_len = group_averaging, k_as_target_len
        self.use_batching_search = use_batching_search
        self.search_batch_size = search_batch_size or 8
        self.raise_empty_query = raise_empty_query
        k = 1 if k is None else k
        self.search_k, self.metric_compute_k = k + 1, k
        self.vectors, self.group_labels, self.query_idxs, self.scores = [], [], [], []

    def update(self, vectors, group_labels=None, query_idxs=None, scores=None):
        self.vectors.append(np.atleast_2d(np.asarray(vectors, dtype=np.float32)))
        if self.dataset_type == 'classification':
            if group_labels is None:
                raise ValueError('In classification dataset group_labels must be not None.')
            self.group_labels.append(np.atleast_1d(np.asarray(group_labels)))
        else:
            if query_idxs is None:
                raise ValueError('In representation dataset query_numbers must be not None.')
            if scores is None:
                raise ValueError('In representation dataset scores must be not None')
            self.query_idxs.append(np.atleast_1d(np.asarray(query_idxs)))
            self.scores.append(np.atleast_2d(np.asarray(scores)))
            self.group_labels.append(np.atleast_1d(np.asarray(group_labels)))

    def neighbours(self):
        """(query_rows, closest global idx (nq, k) after the self-hit drop, relevant lists, ...) — the part the CUDA
        kernel (tok_cosine_topk) replaces."""
        vectors = np.concatenate(self.vectors).astype(np.float32)
        if self.normalize_vectors:
            axis = 0 if self.normalize_axis0 else 1
            vectors = vectors / np.linalg.norm(vectors, axis=axis, keepdims=True)
        labels = np.concatenate(self.group_labels)
        if self.dataset_type == 'classification':
            relevant, gallery, q_rows, q_in = prepare_classification_data(labels, self.raise_empty_query)
            scores = q_cols = None
        else:
            scores = np.concatenate(self.scores)
            relevant, gallery, q_cols, q_rows, q_in = prepare_representation_data(
                np.concatenate(self.query_idxs), scores, self.raise_empty_query)
        return vectors, labels, relevant, gallery, q_rows, q_in, scores, q_cols

    def compute(self):
        vectors, labels, relevant, gallery, q_rows, q_in, scores, q_cols = self.neighbours()
        if self.group_averaging:
            groups = [np.where(labels == u)[0] for u in np.unique(labels)]
        else:
            groups = [np.arange(len(labels))]
        values = []
        for group in groups:
            sel = np.where(np.isin(q_rows, group))[0]
            if self.k_as_target_len:
                k = len(group) + 1 - int((~q_in[sel]).sum())
            else:
                k = self.search_k
            total = 0.0
            bs = self.search_batch_size if self.use_batching_search else max(len(sel), 1)
            for i in range(0, len(sel), bs):
                b = sel[i:i + bs]
                sc, local = flat_search(vectors[gallery], vectors[q_rows[b]], k, self.metric_distance)
                idx = np.where(local >= 0, gallery[np.clip(local, 0, None)], -1)
                idx = drop_self_or_last(idx, q_in[b])
                if min(idx.shape) == 0:
                    continue
                batch_vals = []
                for j, qi in enumerate(b):
                    if q_cols is None:
                        qrels = [(r, 1.0) for r in relevant[qi]]
                    else:
                        qrels = [(r, float(scores[r, q_cols[qi]])) for r in relevant[qi]]
                    run = [int(d) for d in idx[j]]
                    batch_vals.append(self.metric_func(qrels, run, k - 1))
                total += len(b) * float(np.mean(batch_vals))
            values.append(total / len(sel))
        return float(np.mean(values))


def cosine_topk(vectors, k, metric='IP', normalize=True, exclude_self=True):
    """Reference semantics of the N x N retrieval (every row a query against all rows): top-(k+1) then drop the
    first hit (assumed self).  Returns (scores (N,k) float32, idx (N,k) int64).  Used to check tok_cosine_topk."""
    v = np.asarray(vectors, dtype=np.float32)
    if normalize:
        v = v / np.linalg.norm(v, axis=1, keepdims=True)
    sc, idx = flat_search(v, v, k + 1 if exclude_self else k, metric)
    if exclude_self:
        sc, idx = sc[:, 1:], idx[:, 1:]
    return sc.astype(np.float32), idx.astype(np.int64)
