"""SegmentationTask (torchok/tasks/segmentation.py:13-98): backbone.forward_features -> neck -> head; same constructor
signature and `forward_with_gt` dictionary ({'prediction', 'target'})."""
import torch.nn as nn

from ..constructor import BACKBONES, HEADS, NECKS, TASKS
from ..models.base import BackboneWrapper
from .base import BaseTask


@TASKS.register_class
class SegmentationTask(BaseTask):
    def __init__(self, hparams, backbone_name, head_name, neck_name, backbone_params=None, neck_params=None,
                 head_params=None, **kwargs):
        super().__init__(hparams, **kwargs)
        self.backbone = BACKBONES.get(backbone_name)(**(backbone_params or dict()))
        self.neck = NECKS.get(neck_name)(in_channels=self.backbone.out_encoder_channels, **(neck_params or dict()))
        self.head = HEADS.get(head_name)(in_channels=self.neck.out_channels, **(head_params or dict()))

    def forward(self, x):
        return self.head(self.neck(self.backbone.forward_features(x)))

    def forward_with_gt(self, batch):
        input_data = batch.get('image')
        target = batch.get('target')
        prediction = self.head(self.neck(self.backbone.forward_features(input_data)))
        output = {'prediction': prediction}
        if target is not None:
            output['target'] = target
        return output

    def as_module(self):
        return nn.Sequential(BackboneWrapper(self.backbone), self.neck, self.head)
