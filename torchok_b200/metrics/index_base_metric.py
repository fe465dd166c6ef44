"""IndexBasedMeter: accumulate embeddings, search nearest neighbours, feed a ranking metric.

Host-side mirror of torchok/metrics/index_base_metric.py:32-545 (same constructor arguments, `update` keywords and
error behaviour).  What changed is WHERE the work runs: the reference copies everything to the CPU and searches a
faiss IndexFlatIP / IndexFlatL2 in Python-driven batches (`compute` :170-270, `query_generator` :444-521); here the
vectors stay on the GPU, the N x N search is `search_topk` (tok_l2_normalize_rows -> tok_topk_candidates on tcgen05
-> tok_topk_rerank in exact fp32, see include/tokb200.h) and the per-query ranking metrics are evaluated on the
device over the (n_queries, k) result.  Under torch.distributed every rank gathers the vectors once and searches only
its own slice of the query rows (the reference repeats the whole CPU search on every rank).

Deliberate deviation: `normalize_vectors=True` normalises each ROW to unit L2 norm — the behaviour the
reference's own golden answers encode — not `np.linalg.norm(vectors, axis=0)` as index_base_metric.py:192-193 literally
does; `normalize_axis0=True` restores the literal behaviour.
"""
import torch
import torch.distributed as dist

from .. import kernels as K
from .._lib import lib

_DATASETS = ('classification', 'representation')
_DISTANCES = {'IP': 0, 'L2': 1}
MAX_SEARCH_K = 29  # k + 1 <= 29 leaves >= 3 slack candidates in the 32-wide candidate list


def search_topk(queries, gallery, k, metric='IP', gallery_bf16=None, gallery_sqnorm=None):
    """faiss `IndexFlat{IP,L2}(d).add(gallery); .search(queries, k)` on the GPU.

    queries (nq, d), gallery (ng, d): fp32 CUDA tensors.  Returns (scores (nq, k) fp32, idx (nq, k) int64) in faiss
    order (IP descending / squared L2 ascending, ties -> lower index, -1 / -+inf padding when ng < k).
    """
    K.require_cuda(queries, 'queries')
    K.require_cuda(gallery, 'gallery')
    if metric not in _DISTANCES:
        raise ValueError(f'unknown metric distance {metric}')
    if k < 1 or k > 32:
        raise ValueError('search_topk supports 1 <= k <= 32')
    L = lib()
    st = K._st()
    queries = queries.float().contiguous()
    gallery = gallery.float().contiguous()
    nq, d = queries.shape
    ng = gallery.shape[0]
    dev = queries.device
    kp = 8 if k + 3 <= 8 else (16 if k + 3 <= 16 else 32)
    dp = K.ceil8(d)
    if dp > 512:
        raise NotImplementedError('search_topk: embedding size > 512 is not supported by the resident-query kernel')

    def bf16_copy(x):
        n = x.shape[0]
        xb = torch.empty((n, dp), dtype=torch.bfloat16, device=dev)
        sq = torch.empty((n,), dtype=torch.float32, device=dev)
        L.tok_l2_normalize_rows(n, d, 0, K._p(x), None, K._p(xb), dp, K._p(sq), st)
        return xb, sq

    qb, _ = bf16_copy(queries)
    if gallery_bf16 is None:
        gallery_bf16, gallery_sqnorm = bf16_copy(gallery)
    cand_s = torch.empty((nq, kp), dtype=torch.float32, device=dev)
    cand_i = torch.empty((nq, kp), dtype=torch.int32, device=dev)
    L.tok_topk_candidates(nq, ng, dp, kp, K._p(qb), K._p(gallery_bf16),
                          K._p(gallery_sqnorm) if metric == 'L2' else None, K._p(cand_s), K._p(cand_i), st)
    out_s = torch.empty((nq, k), dtype=torch.float32, device=dev)
    out_i = torch.empty((nq, k), dtype=torch.int64, device=dev)
    L.tok_topk_rerank(nq, d, kp, k, _DISTANCES[metric], K._p(queries), K._p(gallery), K._p(cand_i), K._p(out_s),
                      K._p(out_i), st)
    return out_s, out_i


def normalize_rows(x):
    K.require_cuda(x, 'vectors')
    x = x.float().contiguous()
    out = torch.empty_like(x)
    lib().tok_l2_normalize_rows(x.shape[0], x.shape[1], 1, K._p(x), K._p(out), None, 0, None, K._st())
    return out


class IndexBasedMeter:
    """Base class of the retrieval meters (torchmetrics.Metric in the reference; `update` / `compute` / `reset`)."""

    def __init__(self, exact_index, dataset_type, metric_distance, metric_func, k_as_target_len=False, k=None,
                 use_batching_search=True, search_batch_size=None, normalize_vectors=False, group_averaging=False,
                 raise_empty_query=True, normalize_axis0=False, **kwargs):
        if dataset_type not in _DATASETS:
            raise ValueError(f'dataset_type must be one of {_DATASETS}, got {dataset_type}')
        if metric_distance not in _DISTANCES:
            raise ValueError(f'metric_distance must be one of {tuple(_DISTANCES)}, got {metric_distance}')
        if not exact_index:
            raise NotImplementedError('approximate (IVF) index: the GPU search is exact and fast enough to replace it')
        self.exact_index = exact_index
        self.dataset_type = dataset_type
        self.metric_distance = metric_distance
        self.metric_func = metric_func
        self.normalize_vectors = normalize_vectors
        self.normalize_axis0 = normalize_axis0
        self.group_averaging = group_averaging
        self.k_as_target_len = k_as_target_len
        self.use_batching_search = use_batching_search  # kept for signature parity: one fused search here
        self.search_batch_size = search_batch_size
        self.raise_empty_query = raise_empty_query
        k = 1 if k is None else k
        self.search_k = k + 1  # the query itself may be in the index (index_base_metric.py:104-109)
        self.metric_compute_k = k
        self.reset()

    # ---------------------------------------------------------------------------------------------- state
    def reset(self):
        self.vectors, self.group_labels, self.query_idxs, self.scores = [], [], [], []

    def update(self, vectors, group_labels=None, query_idxs=None, scores=None):
        self.vectors.append(vectors.detach())
        if self.dataset_type == 'classification':
            if group_labels is None:
                raise ValueError('In classification dataset group_labels must be not None.')
            self.group_labels.append(group_labels.detach())
        else:
            if query_idxs is None:
                raise ValueError('In representation dataset query_numbers must be not None.')
            if scores is None:
                raise ValueError('In representation dataset scores must be not None')
            self.query_idxs.append(query_idxs.detach())
            self.scores.append(scores.detach())
            self.group_labels.append(group_labels.detach())

    def __call__(self, *args, **kwargs):
        self.update(*args, **kwargs)

    @staticmethod
    def _cat(parts):
        parts = [p.reshape(1, *p.shape) if p.dim() == 0 else p for p in parts]
        return torch.cat(parts)

    def _gathered(self):
        """torchmetrics' dist_reduce_fx='cat' sync (index_base_metric.py:112-120): every rank ends with all rows."""
        vectors = self._cat(self.vectors).float()
        labels = self._cat(self.group_labels)
        qidx = self._cat(self.query_idxs) if self.dataset_type != 'classification' else None
        scores = self._cat(self.scores) if self.dataset_type != 'classification' else None
        if dist.is_available() and dist.is_initialized() and dist.get_world_size() > 1:
            vectors, labels, qidx, scores = (_all_gather_rows(t) for t in (vectors, labels, qidx, scores))
        return vectors, labels, qidx, scores

    # ---------------------------------------------------------------------------------------------- search
    def _search(self, queries, gallery, k):
        """(scores, idx) of the k nearest gallery rows per query — the hot loop; CUDA only."""
        return search_topk(queries, gallery, k, self.metric_distance)

    def _normalize(self, vectors):
        if self.normalize_axis0:
            return vectors / vectors.norm(dim=0, keepdim=True)
        return normalize_rows(vectors) if vectors.is_cuda else vectors / vectors.norm(dim=1, keepdim=True)

    # ---------------------------------------------------------------------------------------------- compute
    def compute(self):
        vectors, labels, qidx, scores = self._gathered()
        if self.normalize_vectors:
            vectors = self._normalize(vectors)
        dev = vectors.device
        n = vectors.shape[0]
        if self.dataset_type == 'classification':
            labels = labels.to(dev).long().reshape(-1)
            counts = torch.bincount(labels - labels.min())
            n_rel = counts[labels - labels.min()] - 1
            if self.raise_empty_query and bool((n_rel == 0).any()):
                bad = int(labels[(n_rel == 0).nonzero()[0, 0]])
                raise ValueError(f'Representation metric. The class {bad} has only one element.')
            query_rows = torch.arange(n, device=dev)
            gallery_rows = query_rows
            query_in_gallery = torch.ones(n, dtype=torch.bool, device=dev)
            group_of_query = labels
        else:
            scores = scores.to(dev).float()
            qidx = qidx.to(dev).long().reshape(-1)
            labels = labels.to(dev).long().reshape(-1)
            query_rows = (qidx >= 0).nonzero().reshape(-1)
            query_cols = qidx[query_rows]
            query_in_gallery = (scores[query_rows] > 0).any(dim=-1)
            keep = torch.ones(n, dtype=torch.bool, device=dev)
            keep[query_rows[~query_in_gallery]] = False
            gallery_rows = keep.nonzero().reshape(-1)
            gains_all = scores[:, query_cols].t().contiguous()  # (nq, n): gain of every row for every query
            n_rel = (gains_all > 0).sum(dim=1)
            if self.raise_empty_query and bool((n_rel == 0).any()):
                raise ValueError('Representation metric. The dataset contains a query vector that does not has '
                                 'relevants. Set parameter raise_empty_query to False for compute.')
            group_of_query = labels[query_rows]

        if self.group_averaging:
            groups = [(group_of_query == g).nonzero().reshape(-1) for g in torch.unique(labels)]
            groups = [g for g in groups if g.numel() > 0]
        else:
            groups = [torch.arange(query_rows.numel(), device=dev)]

        # rank-sharded search: this rank handles a contiguous slice of every group's queries
        rank, world = (dist.get_rank(), dist.get_world_size()) if (dist.is_available() and dist.is_initialized()) \
            else (0, 1)
        gallery = vectors[gallery_rows] if gallery_rows.numel() != n else vectors
        values = []
        for sel in groups:
            if self.k_as_target_len:
                in_group = (labels == labels[query_rows[sel[0]]]).sum() if self.group_averaging else n
                k = int(in_group) + 1 - int((~query_in_gallery[sel]).sum())
            else:
                k = self.search_k
            if k > MAX_SEARCH_K:
                raise NotImplementedError(f'search depth k+1 = {k} exceeds the fused top-k width ({MAX_SEARCH_K})')
            lo, hi = (sel.numel() * rank) // world, (sel.numel() * (rank + 1)) // world
            mine = sel[lo:hi]
            total = torch.zeros((), dtype=torch.float64, device=dev)
            if mine.numel() > 0:
                q_rows = query_rows[mine]
                _, local = self._search(vectors[q_rows], gallery, k)
                closest = torch.where(local >= 0, gallery_rows[local.clamp_min(0)], local)
                # clear_faiss_output (:418-442): drop the first hit when the query is in the index, else the last
                inq = query_in_gallery[mine].unsqueeze(1)
                closest = torch.where(inq, closest[:, 1:], closest[:, :-1])
                if self.dataset_type == 'classification':
                    gains = ((labels[closest.clamp_min(0)] == labels[q_rows].unsqueeze(1)) & (closest >= 0) &
                             (closest != q_rows.unsqueeze(1))).float()
                    ideal = None
                    rel = n_rel[q_rows]
                else:
                    g_all = gains_all[mine]
                    gains = torch.gather(g_all, 1, closest.clamp_min(0)) * (closest >= 0)
                    ideal = torch.sort(g_all, dim=1, descending=True).values[:, :k - 1]
                    rel = n_rel[mine]
                total = self.metric_func(gains, rel, ideal, k - 1).double().sum()
            if world > 1:
                dist.all_reduce(total)
            values.append(total / sel.numel())
        return float(torch.stack(values).mean())


def _all_gather_rows(t):
    """all_gather of tensors whose first dimension differs per rank (pad to the maximum, then trim)."""
    if t is None:
        return None
    world = dist.get_world_size()
    n = torch.tensor([t.shape[0]], device=t.device, dtype=torch.int64)
    sizes = [torch.zeros_like(n) for _ in range(world)]
    dist.all_gather(sizes, n)
    sizes = [int(s) for s in sizes]
    m = max(sizes)
    pad = torch.zeros((m,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[:t.shape[0]] = t
    out = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(out, pad)
    return torch.cat([o[:s] for o, s in zip(out, sizes)])
