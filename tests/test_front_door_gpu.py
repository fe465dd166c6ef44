"""The YAML / CLI front door on the GPU: `python -m torchok_b200 -cp … -cn …` semantics driven
in-process through `torchok_b200.__main__.entrypoint` — config file + overrides → ClassificationTask(resnet18) →
engine.StreamLoop (captured step graph, arena SGD) with the FreezeUnfreeze policy of
examples/configs/classification_cifar10.yaml, an ExponentialLR scheduler, Accuracy / F1Score meters, ModelCheckpoint,
then `resume_path`.  Data is the package's seeded SyntheticImages (no datasets on the box)."""
import os
import textwrap

import pytest
import torch

import torchok_b200 as tb
from torchok_b200.callbacks import Callback

pytestmark = pytest.mark.gpu

YAML = textwrap.dedent('''
    task:
      name: ClassificationTask
      params:
        backbone_name: resnet18
        backbone_params: {pretrained: false, in_channels: 3}
        pooling_name: Pooling
        head_name: ClassificationHead
        head_params: {num_classes: &num_classes 10}
        inputs:
          - shape: [3, &height 32, &width 32]
            dtype: &input_dtype float32
    joint_loss:
      losses:
        - name: CrossEntropyLoss
          mapping: {input: prediction, target: target}
    optimization:
      - optimizer: {name: SGD, params: {lr: 0.01, momentum: 0.9, weight_decay: 0.0001}}
        scheduler: {name: ExponentialLR, params: {gamma: 0.5}}
    data:
      TRAIN:
        - dataloader: {batch_size: 64, num_workers: 0, drop_last: true, shuffle: true}
          dataset:
            name: SyntheticImages
            params: {num_samples: 200, shape: [*height, *width, 3], num_classes: *num_classes, seed: 1, input_dtype: *input_dtype}
            transform:
              - {name: Resize, params: {height: *height, width: *width}}
              - {name: Normalize, params: {mean: [0.485, 0.456, 0.406], std: [0.229, 0.224, 0.225]}}
              - {name: ToTensorV2}
      VALID:
        - dataloader: {batch_size: 64, num_workers: 0, drop_last: false, shuffle: false}
          dataset:
            name: SyntheticImages
            params: {num_samples: 96, shape: [*height, *width, 3], num_classes: *num_classes, seed: 2, input_dtype: *input_dtype}
            transform:
              - {name: Normalize, params: {mean: [0.485, 0.456, 0.406], std: [0.229, 0.224, 0.225]}}
              - {name: ToTensorV2}
    trainer: {accelerator: gpu, max_epochs: 3, precision: 16, num_sanity_val_steps: 0, log_every_n_steps: 2}
    seed_params: {seed: 42, workers: true}
    logger: {log_dir: '${oc.env:TOK_TEST_LOGS}', experiment_name: resnet18, name: CSVLogger}
    callbacks:
      - name: ModelCheckpoint
        params: {monitor: valid/F1Score, save_top_k: 1, save_last: true, mode: max}
      - name: FreezeUnfreeze
        params:
          freeze_modules:
            - {module_name: backbone, epoch: 1}
            - {module_name: backbone, stages: 1}
            - {module_name: backbone, module_class: _BatchNorm, bn_requires_grad: false, bn_track_running_stats: false}
      - name: SnapshotForTest
      - name: TQDMProgressBar
        params: {refresh_rate: 5}
    metrics:
      - name: Accuracy
        params: {task: multiclass, num_classes: *num_classes}
        mapping: {preds: prediction, target: target}
      - name: F1Score
        params: {task: multiclass, num_classes: *num_classes, average: macro}
        mapping: {preds: prediction, target: target}
''')

WATCHED = ('backbone.conv1.weight', 'backbone.layer1.0.conv1.weight', 'backbone.layer4.0.conv1.weight',
           'backbone.layer4.0.bn1.weight', 'backbone.layer4.0.bn1.running_mean', 'backbone.layer4.0.bn1.running_var',
           'head.fc.weight')
SNAPSHOTS = []


class SnapshotForTest(Callback):
    def on_train_epoch_start(self, runner):
        sd = runner.task.state_dict()
        SNAPSHOTS.append({k: sd[k].detach().float().cpu().clone() for k in WATCHED})

    def teardown(self, runner):
        self.on_train_epoch_start(runner)


if 'SnapshotForTest' not in tb.CALLBACKS:
    tb.CALLBACKS.register_class(SnapshotForTest)


def test_cli_trains_freezes_thaws_checkpoints_and_resumes(tmp_path, monkeypatch):
    from torchok_b200.__main__ import entrypoint
    (tmp_path / 'configs').mkdir()
    (tmp_path / 'configs' / 'cls.yaml').write_text(YAML)
    monkeypatch.setenv('TOK_TEST_LOGS', str(tmp_path / 'logs'))
    del SNAPSHOTS[:]
    logs = entrypoint(['-cp', str(tmp_path / 'configs'), '-cn', 'cls'])
    for key in ('train/loss', 'valid/loss', 'train/Accuracy', 'valid/Accuracy', 'valid/F1Score'):
        assert key in logs and logs[key] == logs[key], key                        # present and not NaN
    assert 0.5 < logs['train/loss'] < 10 and 0.0 <= logs['valid/Accuracy'] <= 1.0
    e0, e1, e2, end = SNAPSHOTS
    same = lambda a, b, k: torch.equal(a[k], b[k])  # noqa: E731
    # epoch 0: the whole backbone is frozen, only the head trains; frozen BatchNorms do not track statistics
    for k in WATCHED[:-1]:
        assert same(e0, e1, k), k
    assert not same(e0, e1, 'head.fc.weight')
    # epochs 1-2: backbone thawed except stem + layer1 (stages: 1) and every BatchNorm (weights and statistics)
    assert not same(e1, end, 'backbone.layer4.0.conv1.weight') and not same(e1, e2, 'backbone.layer4.0.conv1.weight')
    for k in ('backbone.conv1.weight', 'backbone.layer1.0.conv1.weight', 'backbone.layer4.0.bn1.weight',
              'backbone.layer4.0.bn1.running_mean', 'backbone.layer4.0.bn1.running_var'):
        assert same(e0, end, k), k
    out = tmp_path / 'logs' / 'resnet18'
    header = open(out / 'metrics.csv').readline()
    assert 'valid/F1Score' in header and 'lr' in header
    ckpts = sorted(os.listdir(out / 'checkpoints'))
    assert 'last.ckpt' in ckpts and len(ckpts) == 2
    last = torch.load(out / 'checkpoints' / 'last.ckpt', weights_only=False)
    assert last['epoch'] == 2 and last['global_step'] == 9
    assert 'buf' in last['optimizer_states'][0]['state'] and last['optimizer_states'][0]['step'] == 9
    assert torch.equal(last['state_dict']['head.fc.weight'].float(), end['head.fc.weight'])

    # resume: one more epoch from last.ckpt, lr continues the schedule (0.01 * 0.5**3 during epoch 3)
    del SNAPSHOTS[:]
    monkeypatch.setenv('TOK_TEST_LOGS', str(tmp_path / 'logs2'))
    logs2 = entrypoint(['-cp', str(tmp_path / 'configs'), '-cn', 'cls.yaml', 'trainer.max_epochs=4',
                        f'resume_path={out / "checkpoints" / "last.ckpt"}'])
    start, stop = SNAPSHOTS
    assert torch.equal(start['head.fc.weight'], end['head.fc.weight'])             # weights came from the checkpoint
    assert not torch.equal(start['head.fc.weight'], stop['head.fc.weight'])
    assert logs2['lr'] == pytest.approx(0.01 * 0.5 ** 3) or logs2['lr'] == pytest.approx(0.01 * 0.5 ** 4)
    assert logs2['train/loss'] == logs2['train/loss']
    resumed = torch.load(tmp_path / 'logs2' / 'resnet18' / 'checkpoints' / 'last.ckpt', weights_only=False)
    assert resumed['epoch'] == 3 and resumed['global_step'] == 12

    # test mode: metrics over the TEST loaders from a checkpoint given as task.load_checkpoint
    monkeypatch.setenv('TOK_TEST_LOGS', str(tmp_path / 'logs3'))
    cfg = tb.load_config(str(tmp_path / 'configs' / 'cls.yaml'))
    cfg['data']['TEST'] = cfg['data']['VALID']
    cfg['task']['load_checkpoint'] = {'base_ckpt_path': str(out / 'checkpoints' / 'last.ckpt')}
    cfg['callbacks'] = []
    from torchok_b200.runner import Runner
    tlogs = Runner(cfg).run('test')
    assert set(tlogs) == {'test/Accuracy', 'test/F1Score'}
    # same weights, same 96 images: equal up to one near-tie argmax flip between two builds of the task
    assert float(tlogs['test/Accuracy']) == pytest.approx(logs['valid/Accuracy'], abs=1.5 / 96)
