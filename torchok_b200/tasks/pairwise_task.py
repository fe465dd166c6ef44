"""PairwiseLearnTask (torchok/tasks/pairwise_task.py:12-107): ClassificationTask whose `forward_with_gt` returns
{'emb1', 'emb2', 'R', 'target'} for pairwise losses.

Deviation: the reference forwards `pooling_name, head_name, neck_name` POSITIONALLY into
ClassificationTask's `(neck_name, pooling_name, head_name)` slots (pairwise_task.py:41-52), so with
examples/configs/pairwise_sop.yaml it looks `Pooling` up in the NECKS registry and raises KeyError.  The intended
(keyword) wiring is implemented so that the YAML drops in unchanged.

`calc_relevance_matrix` keeps the reference's semantics (R[i, j] = 1 when samples i and j share a label) but compares
labels directly for the single-label case instead of multiplying two (B x num_classes) one-hot matrices.
"""
import torch

from ..constructor import TASKS
from .classification import ClassificationTask


@TASKS.register_class
class PairwiseLearnTask(ClassificationTask):
    def __init__(self, hparams, num_classes, backbone_name, pooling_name, head_name, neck_name=None,
                 backbone_params=None, neck_params=None, pooling_params=None, head_params=None, inputs=None):
        super().__init__(hparams, backbone_name=backbone_name, neck_name=neck_name, pooling_name=pooling_name,
                         head_name=head_name, backbone_params=backbone_params, neck_params=neck_params,
                         pooling_params=pooling_params, head_params=head_params, inputs=inputs)
        self.num_classes = num_classes

    def forward_with_gt(self, batch):
        input_data = batch.get('image')
        target = batch.get('target')
        embedding = self.forward(input_data)
        output = {'emb1': embedding, 'emb2': embedding}
        if target is not None:
            output['R'] = self.calc_relevance_matrix(target)
            output['target'] = target
        return output

    def calc_relevance_matrix(self, y):
        if y.ndim == 1:
            # the range check reads a device flag back (as F.one_hot does in the reference): not possible while the
            # step is being captured into a CUDA graph — the eager warm-up steps before the capture have run it
            capturing = y.is_cuda and torch.cuda.is_current_stream_capturing()
            if not capturing and bool(((y < 0) | (y >= self.num_classes)).any()):
                raise RuntimeError('calc_relevance_matrix: label outside [0, num_classes)')
            return (y[:, None] == y[None, :]).float()
        y = y.float()
        intersections = torch.matmul(y, y.transpose(1, 0))
        return torch.where(intersections > 0, 1., 0.)
