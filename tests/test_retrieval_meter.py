"""The product's IndexBasedMeter / ranx meters (torchok_b200/metrics) against the reference's known answers.

CPU part (`not gpu`): the host logic (query / gallery bookkeeping, self-hit removal, gains, metric formulas, group
averaging, rank sharding) with the nearest-neighbour SEARCH stubbed by the oracle's brute force — the search itself is
CUDA only and is covered by the gpu-marked tests below, which run the real kernels on the same golden vectors and on
random data against oracle/retrieval.py.
"""
import json
import os

import numpy as np
import pytest
import torch

import torchok_b200 as tb
from oracle import retrieval as orc
from torchok_b200.metrics import index_base_metric as ibm

G = json.load(open(os.path.join(os.path.dirname(__file__), 'golden', 'retrieval_kat.json')))
V = torch.tensor(G['vectors'], dtype=torch.float32)
METERS = {'precision': 'PrecisionAtKMeter', 'recall': 'RecallAtKMeter',
          'average_precision': 'MeanAveragePrecisionAtKMeter', 'ndcg': 'NDCGAtKMeter', 'hit_rate': 'HitAtKMeter'}


def oracle_search(self, queries, gallery, k):
    s, i = orc.flat_search(gallery.cpu().numpy(), queries.cpu().numpy(), k, self.metric_distance)
    return torch.from_numpy(s).to(queries.device), torch.from_numpy(i.astype(np.int64)).to(queries.device)


def run_meter(metric, dataset_type, device, scores_key='scores', **params):
    out = {}
    for k in range(1, G['max_k'] + 1):
        m = tb.METRICS.get(METERS[metric])(dataset_type=dataset_type, k=k, **params)
        for i in range(len(V)):  # BATCH_SIZE = 1 as in the reference's context.py
            if dataset_type == 'classification':
                m.update(vectors=V[i:i + 1].to(device), group_labels=torch.tensor(G['targets'][i:i + 1]).to(device))
            else:
                m.update(vectors=V[i:i + 1].to(device), group_labels=torch.tensor(G['group_labels'][i:i + 1]).to(device),
                         query_idxs=torch.tensor(G['queries_idx'][i:i + 1]).to(device),
                         scores=torch.tensor(G[scores_key][i:i + 1]).to(device))
        out[k] = m.compute()
    return out


CASES = [('classification', m, 'classification_answers', 'scores', dict(normalize_vectors=True))
         for m in ('precision', 'recall', 'average_precision')] + \
        [('representation', m, 'representation_answers', 'scores', {})
         for m in ('precision', 'recall', 'average_precision', 'ndcg')] + \
        [('representation', m, 'representation_query_as_relevant_answers', 'scores_query_as_relevant',
          dict(normalize_vectors=True)) for m in ('precision', 'recall')]


@pytest.mark.parametrize('dataset,metric,answers,scores_key,params', CASES)
def test_host_logic_reproduces_reference_answers(monkeypatch, dataset, metric, answers, scores_key, params):
    monkeypatch.setattr(ibm.IndexBasedMeter, '_search', oracle_search)
    got = run_meter(metric, dataset, 'cpu', scores_key, **params)
    for k, v in got.items():
        np.testing.assert_almost_equal(v, G[answers][metric][str(k)], decimal=6)


def test_error_behaviour_matches_reference():
    with pytest.raises(ValueError):
        tb.METRICS.get('HitAtKMeter')(dataset_type='detection')
    with pytest.raises(ValueError):
        tb.METRICS.get('HitAtKMeter')(dataset_type='classification', metric_distance='cosine')
    m = tb.METRICS.get('HitAtKMeter')(dataset_type='classification')
    with pytest.raises(ValueError, match='group_labels must be not None'):
        m.update(vectors=V)
    m = tb.METRICS.get('HitAtKMeter')(dataset_type='representation')
    with pytest.raises(ValueError, match='scores must be not None'):
        m.update(vectors=V, query_idxs=torch.zeros(9))
    # no CPU fallback for the search itself
    with pytest.raises(RuntimeError, match='CUDA'):
        ibm.search_topk(V, V, 2)


def _gloo_worker(rank, world, port, ret):
    import torch.distributed as dist
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world)
    ibm.IndexBasedMeter._search = oracle_search
    res = {}
    for k in (1, 3):
        m = tb.METRICS.get('PrecisionAtKMeter')(dataset_type='classification', k=k, normalize_vectors=True)
        for i in range(rank, len(V), world):  # every rank saw a different part of the validation set
            m.update(vectors=V[i:i + 1], group_labels=torch.tensor(G['targets'][i:i + 1]))
        res[k] = m.compute()
    ret[rank] = res
    dist.destroy_process_group()


def test_rank_sharded_compute_gloo_world2():
    """N>1 path on CPU: rows gathered from both ranks, each rank scores its slice of the queries, partial sums are
    all-reduced.  (Gathered row ORDER differs from the single-process order; the metric is order-invariant.)"""
    import torch.multiprocessing as mp
    ctx = mp.get_context('spawn')
    ret = ctx.Manager().dict()
    port = 29500 + os.getpid() % 1000
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, ret)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    for r in range(2):
        for k in (1, 3):
            np.testing.assert_almost_equal(ret[r][k], G['classification_answers']['precision'][str(k)], decimal=6)


# ------------------------------------------------------------------------------------------------------ GPU
@pytest.mark.gpu
@pytest.mark.parametrize('dataset,metric,answers,scores_key,params', CASES)
def test_gpu_meter_reproduces_reference_answers(dataset, metric, answers, scores_key, params):
    got = run_meter(metric, dataset, 'cuda', scores_key, **params)
    for k, v in got.items():
        np.testing.assert_almost_equal(v, G[answers][metric][str(k)], decimal=6)


@pytest.mark.gpu
@pytest.mark.parametrize('n,d,k,metric', [(9, 4, 8, 'IP'), (300, 32, 5, 'IP'), (1000, 128, 10, 'IP'),
                                          (4097, 512, 4, 'IP'), (777, 72, 6, 'L2'), (129, 8, 28, 'IP')])
def test_search_topk_indices_are_bit_exact_vs_oracle(n, d, k, metric):
    """index work: the returned neighbour indices must equal the brute-force oracle's exactly."""
    rng = np.random.default_rng(n + d)
    v = rng.standard_normal((n, d)).astype(np.float32)
    if metric == 'IP':
        v /= np.linalg.norm(v, axis=1, keepdims=True)
    s_ref, i_ref = orc.flat_search(v, v, min(k, n), metric)
    t = torch.from_numpy(v).cuda()
    s, i = ibm.search_topk(t, t, min(k, n), metric)
    s, i = s.cpu().numpy(), i.cpu().numpy()
    # exact index equality wherever the oracle's neighbouring scores are separated by more than fp32 noise
    gap_ok = np.ones_like(i_ref, dtype=bool)
    gap = np.abs(np.diff(s_ref, axis=1))
    tie = gap < 1e-6 * np.maximum(1.0, np.abs(s_ref[:, 1:]))
    gap_ok[:, 1:] &= ~tie
    gap_ok[:, :-1] &= ~tie
    assert (i[gap_ok] == i_ref[gap_ok]).all()
    assert gap_ok.mean() > 0.99
    np.testing.assert_allclose(s, s_ref, rtol=1e-4, atol=1e-5)


@pytest.mark.gpu
def test_search_topk_edge_cases():
    t = torch.randn(5, 16).cuda()
    s, i = ibm.search_topk(t, t[:2], 4)           # fewer gallery rows than k: faiss-style -1 padding
    assert (i[:, 2:] == -1).all() and torch.isinf(s[:, 2:]).all() and (i[:, :2] >= 0).all()
    big = torch.randn(3, 16).cuda()
    s, i = ibm.search_topk(big, torch.cat([big, big]), 2)   # exact duplicates: lower index first
    assert (i[:, 0] == torch.arange(3).cuda()).all() and (i[:, 1] == torch.arange(3).cuda() + 3).all()
    with pytest.raises(ValueError):
        ibm.search_topk(t, t, 0)


@pytest.mark.gpu
def test_large_search_self_retrieval_property():
    """Size-independent property at a size the oracle cannot brute-force quickly: with unit-norm rows every vector's
    nearest neighbour is itself, and a planted near-duplicate is its second."""
    n, d = 50000, 512
    g = torch.Generator(device='cuda').manual_seed(1)
    v = torch.randn(n, d, device='cuda', generator=g)
    v[1::2] = v[0::2] + 0.01 * torch.randn(n // 2, d, device='cuda', generator=g)
    v = ibm.normalize_rows(v)
    s, i = ibm.search_topk(v, v, 2)
    ar = torch.arange(n, device='cuda')
    assert (i[:, 0] == ar).all()
    assert (i[:, 1] == (ar ^ 1)).all()
