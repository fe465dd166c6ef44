"""Regenerates tests/golden/retrieval_kat.json from the reference's own known-answer test data.

Run in the build container (where /root/reference exists):  python tests/golden/make_retrieval_golden.py
It imports /root/reference/tests/base_tests/metrics/representation/data.py (pure data: tensors + dicts of exact
fractions; the module only needs torch) and writes the vectors / labels / score matrices / expected metric values as
JSON, so the tests never need /root/reference at run time.  MAX_K = 6, BATCH_SIZE = 1 come from context.py:12-14.
"""
import importlib.util
import json
import os

SRC = '/root/reference/tests/base_tests/metrics/representation/data.py'
HERE = os.path.dirname(os.path.abspath(__file__))


def main():
    spec = importlib.util.spec_from_file_location('ref_representation_data', SRC)
    d = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(d)

    def ans(table):
        return {name: {str(k): float(v) for k, v in by_k.items()} for name, by_k in table.items()}

    out = {
        'source': 'eora-ai/torchok tests/base_tests/metrics/representation/data.py (+ context.py MAX_K=6, BATCH_SIZE=1)',
        'max_k': 6,
        'vectors': d.VECTORS.tolist(),
        'targets': d.TARGETS.tolist(),
        'group_labels': d.GROUP_LABELS.tolist(),
        'queries_idx': d.QUERIES_IDX.tolist(),
        'scores': d.SCORES.tolist(),
        'scores_query_as_relevant': d.SCORES_QUERY_AS_RELEVANT.tolist(),
        'classification_answers': ans(d.CLASSIFICATION_ANSWERS),
        'representation_answers': ans(d.REPRESENTATION_ANSWERS),
        'representation_query_as_relevant_answers': ans(d.REPRESENTATION_QUERY_AS_RELEVANT_ANSWERS),
        # hand-derived neighbour lists from the comments of data.py:65-100 (classification dataset, cosine)
        'classification_closest': {
            '0': [5, 6, 1, 3, 2, 4, 7, 8], '1': [0, 6, 7, 3, 4, 5, 2, 8], '2': [5, 0, 8, 6, 3, 4, 1, 7],
            '3': [6, 5, 0, 1, 8, 2, 7, 4], '4': [7, 1, 0, 5, 2, 6, 3, 8], '5': [0, 6, 3, 2, 1, 8, 4, 7],
            '6': [3, 5, 0, 1, 2, 8, 7, 4], '7': [4, 1, 0, 6, 3, 5, 2, 8], '8': [2, 5, 6, 0, 3, 7, 1, 4]},
    }
    with open(os.path.join(HERE, 'retrieval_kat.json'), 'w') as f:
        json.dump(out, f, indent=1)
    print('wrote retrieval_kat.json')


if __name__ == '__main__':
    main()
