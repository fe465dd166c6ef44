"""ClassificationTask: backbone -> neck -> pooling -> head, assembled from the registries
(torchok/tasks/classification.py:13-123; same constructor signature, same `forward_with_gt` dictionary)."""
import torch.nn as nn

from ..constructor import BACKBONES, HEADS, NECKS, POOLINGS, TASKS
from .base import BaseTask


@TASKS.register_class
class ClassificationTask(BaseTask):
    def __init__(self, hparams, backbone_name, neck_name=None, pooling_name=None, head_name=None,
                 backbone_params=None, neck_params=None, pooling_params=None, head_params=None, inputs=None):
        super().__init__(hparams, inputs)
        self.backbone = BACKBONES.get(backbone_name)(**(backbone_params or dict()))

        if neck_name is None:
            self.neck = nn.Identity()
            pooling_in_channels = self.backbone.out_channels
        else:
            self.neck = NECKS.get(neck_name)(in_channels=self.backbone.out_encoder_channels, **(neck_params or dict()))
            pooling_in_channels = self.neck.out_channels

        if pooling_name is None:
            self.pooling = nn.Identity()
            head_in_channels = self.backbone.out_channels
        else:
            self.pooling = POOLINGS.get(pooling_name)(in_channels=pooling_in_channels, **(pooling_params or dict()))
            head_in_channels = self.pooling.out_channels

        if head_name is None:
            self.head = nn.Identity()
        else:
            self.head = HEADS.get(head_name)(in_channels=head_in_channels, **(head_params or dict()))

    def forward(self, x):
        return self.head(self.pooling(self.neck(self.backbone(x))))

    def forward_with_gt(self, batch):
        image = batch.get('image')
        target = batch.get('target')
        features = self.neck(self.backbone(image))
        embeddings = self.pooling(features)
        prediction = self.head(embeddings, target) if not isinstance(self.head, nn.Identity) else embeddings
        output = {'embeddings': embeddings, 'prediction': prediction}
        if target is not None:
            output['target'] = target
        return output

    def as_module(self):
        return nn.Sequential(self.backbone, self.neck, self.pooling, self.head)
