"""Runner: what `python -m torchok_b200 -cp … -cn …` drives.

The reference builds a pytorch_lightning.Trainer from the config (torchok/constructor/runner.py:7-19) and calls
`trainer.fit / test / predict(model, ckpt_path=config.resume_path)` (torchok/__main__.py:38-50); Lightning then runs
`training_step`, the optimizer, schedulers, callbacks, loggers and checkpoints.  Here the same config drives

    task = TASKS.get(cfg.task.name)(cfg, **cfg.task.params)              # unchanged plug-in boundary
    loop = engine.StreamLoop(task, cfg.optimization[0].optimizer)         # captured CUDA-graph step, arena optimizer
    for epoch: for batch: loop.train_step(batch)                          # no host sync inside an epoch

with the host-side pieces Lightning provided restated around it: `task.load_checkpoint` on start, `resume_path`
(model + optimizer + scheduler + callbacks + epoch), per-epoch validation with `MetricsManager`, lr schedulers
(optim.LrDriver), callbacks (callbacks.FreezeUnfreeze / ModelCheckpoint / EarlyStopping), CSV (+ TensorBoard when the
config names it) logging, `limit_*_batches`, `max_epochs` / `max_steps`, `check_val_every_n_epoch`,
`num_sanity_val_steps` (ignored), `log_every_n_steps`.  `accumulate_grad_batches`, `gradient_clip_val`, `sync_batchnorm`
and `overfit_batches` raise NotImplementedError when set to a non-neutral value.  Trainer keys about devices /
precision / strategy are accepted:
the device is `cuda:LOCAL_RANK`, compute is bf16 with fp32 masters, and data parallelism follows torch.distributed's
environment (one process per GPU, launched by torchrun) — not the trainer block.

There is no CPU path: the step engine refuses to start without the CUDA extension (engine.StreamLoop → _lib.lib()).
"""
import csv
import os
import random
import time

import numpy as np
import torch
import torch.distributed as dist

from .callbacks import Callback  # noqa: F401  (registers the callbacks)
from .constructor import CALLBACKS, TASKS
from .constructor.config import Config, load_config
from .constructor.load import load_checkpoint
from .constructor.paramwise import paramwise_multipliers
from .data import create_dataloaders
from .metrics.metrics_manager import MetricsManager, Phase
from .optim import LrDriver

MODES = ('train', 'test', 'predict')


def seed_everything(seed=None, workers=False):
    """pytorch_lightning.seed_everything as used in torchok/__main__.py:33-34."""
    if seed is None:
        return None
    seed = int(seed)
    os.environ['PL_GLOBAL_SEED'] = str(seed)
    os.environ['PL_SEED_WORKERS'] = str(int(bool(workers)))
    random.seed(seed)
    np.random.seed(seed % (2 ** 32))
    torch.manual_seed(seed)
    return seed


def _limit(n_batches, limit):
    """Lightning's limit_*_batches: int = that many batches, float = that fraction, None = all."""
    if limit is None:
        return n_batches
    if isinstance(limit, float) and limit <= 1.0 and not float(limit).is_integer():
        return int(n_batches * limit)
    if isinstance(limit, float) and limit == 1.0:
        return n_batches
    return min(n_batches, int(limit))


class ScalarLog:
    """metrics.csv (step, epoch, one column per key — the layout of Lightning's CSVLogger) and, when the config's
    logger is TensorBoardLogger and tensorboard is importable, the same scalars as TB events."""

    def __init__(self, output_dir, tensorboard=False, enabled=True):
        self.output_dir, self.enabled, self.rows, self.keys = output_dir, enabled, [], []
        self.tb = None
        if enabled:
            os.makedirs(output_dir, exist_ok=True)
            if tensorboard:
                try:
                    from torch.utils.tensorboard import SummaryWriter
                    self.tb = SummaryWriter(log_dir=output_dir)
                except Exception:   # tensorboard not installed: CSV only
                    self.tb = None

    def log(self, scalars, step, epoch):
        if not self.enabled or not scalars:
            return
        row = {'step': step, 'epoch': epoch}
        for k, v in scalars.items():
            row[k] = float(v)
            if k not in self.keys:
                self.keys.append(k)
            if self.tb is not None:
                self.tb.add_scalar(k, float(v), step)
        self.rows.append(row)
        with open(os.path.join(self.output_dir, 'metrics.csv'), 'w', newline='') as f:
            w = csv.DictWriter(f, fieldnames=['step', 'epoch'] + self.keys)
            w.writeheader()
            w.writerows(self.rows)

    def close(self):
        if self.tb is not None:
            self.tb.close()


class Runner:
    def __init__(self, config, overrides=None, loop_factory=None, device=None):
        """`config`: path / dict / Config.  `loop_factory(task, optimizer_cfg)` replaces engine.StreamLoop (the tests of
        this host logic pass a stand-in; the product path never does)."""
        self.cfg = config if isinstance(config, Config) and not overrides else load_config(config, overrides)
        cfg = self.cfg
        self.trainer = Config.wrap(dict(cfg.get('trainer') or {}))
        # trainer keys that would change the arithmetic of a step are refused rather than silently ignored
        for key, neutral in (('accumulate_grad_batches', (None, 1)), ('gradient_clip_val', (None, 0, 0.0)),
                             ('sync_batchnorm', (None, False)), ('overfit_batches', (None, 0, 0.0))):
            if self.trainer.get(key) not in neutral:
                raise NotImplementedError(f'trainer.{key}={self.trainer.get(key)!r}: not built in the stream loop '
                                          f'(one optimizer step per batch, unclipped gradients, per-GPU BatchNorm '
                                          f'statistics)')
        seed_everything(**dict(cfg.get('seed_params') or {}))
        self.rank = dist.get_rank() if dist.is_available() and dist.is_initialized() else 0
        self.world = dist.get_world_size() if dist.is_available() and dist.is_initialized() else 1
        self._loop_factory = loop_factory
        if device is None:
            device = torch.device('cuda', int(os.environ.get('LOCAL_RANK', 0))) if loop_factory is None else 'cpu'
        self.device = torch.device(device)
        if loop_factory is None:
            from ._lib import lib
            lib()                                   # fail now, loudly, if libtokb200.so is missing
            if not torch.cuda.is_available():
                raise RuntimeError('torchok_b200 needs a CUDA device (sm_100a): there is no CPU path')
            torch.cuda.set_device(self.device)
        self.task = TASKS.get(cfg.task.name)(cfg, **dict(cfg.task.get('params') or {})).to(self.device)
        self.metrics_manager = MetricsManager(cfg.get('metrics') or [])
        self.callbacks = [CALLBACKS.get(c['name'])(**dict(c.get('params') or {})) for c in (cfg.get('callbacks') or [])]
        lg = cfg.get('logger')
        if lg:
            parts = [lg['log_dir'], lg.get('experiment_name', 'default')] + ([lg['timestamp']] if lg.get('timestamp') else [])
            self.output_dir = os.path.join(*[str(p) for p in parts])
        else:
            self.output_dir = os.path.join(os.getcwd(), 'torchok_b200_logs')
        self.logger = ScalarLog(self.output_dir, tensorboard=bool(lg) and lg.get('name') == 'TensorBoardLogger',
                                enabled=self.rank == 0)
        self.loop = self.scheduler = None
        self.current_epoch = self.global_step = 0
        self.should_stop = False
        self.has_validation = bool((cfg.get('data') or {}).get('VALID'))
        self.logged = {}                             # most recent value of every logged key (callback monitors)

    # ------------------------------------------------------------------------------------------------ construction
    def _build_loop(self):
        opt_cfgs = self.cfg.get('optimization') or []
        if len(opt_cfgs) != 1:
            raise NotImplementedError(f'{len(opt_cfgs)} optimization entries: the stream loop drives exactly one '
                                      f'optimizer over the whole task')
        opt = opt_cfgs[0]
        if self._loop_factory is not None:
            self.loop = self._loop_factory(self.task, opt['optimizer'])
        else:
            from .engine import StreamLoop
            self.loop = StreamLoop(self.task, optimizer=opt['optimizer'])
        self._base_mults = paramwise_multipliers(self.task, dict(opt['optimizer'].get('paramwise_cfg') or {}))
        sch = opt.get('scheduler')
        if sch:
            self.scheduler = LrDriver(self.loop.optimizer, sch['name'], sch.get('params'), sch.get('pl_params'))

    def frozen_set_changed(self):
        """Called by FreezeUnfreeze: frozen parameters get zero lr / decay multipliers, the step graph is dropped."""
        if self.loop is None:
            return
        mults = {p: (m if p.requires_grad else (0.0, 0.0)) for p, m in self._base_mults.items()}
        self.loop.optimizer.set_param_multipliers(mults)
        self.loop.reset_graph()

    # ------------------------------------------------------------------------------------------------ checkpoints
    def save_checkpoint(self, path, weights_only=False):
        # the optimizer state is sharded over the ranks under the peer-fused gradient exchange: gathering it is collective
        opt_state = self.loop.optimizer.state_dict(self.task) if (not weights_only and self.loop is not None) else None
        if self.rank != 0:
            return
        ckpt = {'state_dict': {k: v.detach().cpu() for k, v in self.task.state_dict().items()},
                'epoch': self.current_epoch, 'global_step': self.global_step,
                'torchok_b200': True, 'hyper_parameters': self.cfg.to_dict()}
        if not weights_only and self.loop is not None:
            ckpt['optimizer_states'] = [opt_state]
            ckpt['lr_schedulers'] = [self.scheduler.state_dict()] if self.scheduler else []
            ckpt['callbacks'] = {self._callback_key(i, c): c.state_dict() for i, c in enumerate(self.callbacks)}
        os.makedirs(os.path.dirname(os.path.abspath(path)), exist_ok=True)
        tmp = f'{path}.tmp'
        torch.save(ckpt, tmp)
        os.replace(tmp, path)

    def _callback_key(self, index, c):
        ident = {k: getattr(c, k) for k in ('monitor', 'mode', 'every_n_epochs', 'save_top_k') if hasattr(c, k)}
        same = [x for x in self.callbacks if type(x) is type(c)]
        suffix = f'#{same.index(c)}' if len(same) > 1 else ''
        return f'{type(c).__name__}{ident}{suffix}'

    def remove_checkpoint(self, path):
        if self.rank == 0 and os.path.exists(path):
            os.remove(path)

    def _resume(self, path, restore_training_state=True):
        ckpt = torch.load(path, map_location='cpu', weights_only=False)
        self.task.load_state_dict(ckpt['state_dict'] if 'state_dict' in ckpt else ckpt)
        if self.loop is not None:
            self.loop.arena.refresh_shadow()
        if not restore_training_state:
            return
        if ckpt.get('optimizer_states') and self.loop is not None and ckpt.get('torchok_b200'):
            self.loop.optimizer.load_state_dict(ckpt['optimizer_states'][0], self.task)
        if ckpt.get('lr_schedulers') and self.scheduler is not None and ckpt.get('torchok_b200'):
            self.scheduler.load_state_dict(ckpt['lr_schedulers'][0])
        saved = ckpt.get('callbacks') or {}
        for i, c in enumerate(self.callbacks):
            # keyed like Lightning's state_key: class name + the parameters that tell two instances apart (two
            # ModelCheckpoints monitoring different metrics); older checkpoints used the bare class name
            state = saved.get(self._callback_key(i, c), saved.get(type(c).__name__))
            if state and ckpt.get('torchok_b200'):
                c.load_state_dict(state)
        self.current_epoch = int(ckpt.get('epoch', -1)) + 1       # the stored epoch was completed
        self.global_step = int(ckpt.get('global_step', 0))

    # ------------------------------------------------------------------------------------------------------- loops
    def _to_device(self, batch):
        return {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v) for k, v in batch.items()}

    def _log(self, scalars):
        scalars = {k: float(v) for k, v in scalars.items()}
        self.logged.update(scalars)
        self.logger.log(scalars, self.global_step, self.current_epoch)

    def _evaluate(self, phase, loaders, limit):
        """validation_step / test_step over every loader: metrics + (VALID only) mean loss; tasks/base.py:135-161."""
        self.task.eval()
        logs = {}
        with torch.no_grad():
            for dl_idx, loader in enumerate(loaders):
                sums, count = {}, 0
                n = _limit(len(loader), limit)
                for i, batch in enumerate(loader):
                    if i >= n:
                        break
                    batch = self._to_device(batch)
                    if phase is Phase.VALID:
                        losses, output = self.task.validation_step(batch, i, dl_idx)
                        for k, v in losses.items():
                            sums[k] = sums.get(k, 0.0) + v.detach().float()
                        count += 1
                    else:
                        output = self.task.forward_with_gt(batch)
                    self.metrics_manager.update(phase, dl_idx, **output)
                # with several loaders Lightning suffixes every logged key with its dataloader index
                suffix = f'/dataloader_idx_{dl_idx}' if len(loaders) > 1 else ''
                for k, v in sums.items():
                    logs[f'{phase.value}/{k}{suffix}'] = self._mean_over_ranks(v / max(count, 1))
        logs.update(self.metrics_manager.on_epoch_end(phase))
        return logs

    def _mean_over_ranks(self, value):
        """The reference all-gathers every step's loss dict and logs the mean (tasks/base.py:163-173); one scalar
        all-reduce per epoch gives the same epoch mean."""
        value = value if torch.is_tensor(value) else torch.tensor(float(value), device=self.device)
        if self.world > 1:
            value = value.clone()
            dist.all_reduce(value)
            value = value / self.world
        return float(value)

    def fit(self):
        cfg, tr = self.cfg, self.trainer
        lc = cfg.task.get('load_checkpoint')
        if lc:
            load_checkpoint(self.task, **dict(lc))                          # BaseTask.on_fit_start
        self._build_loop()
        for c in self.callbacks:
            c.setup(self)
        if cfg.get('resume_path'):
            self._resume(cfg['resume_path'])
        train_loaders = create_dataloaders(cfg.data, 'TRAIN')
        if len(train_loaders) != 1:
            raise ValueError(f'TRAIN phase needs exactly one dataloader, got {len(train_loaders)}')
        train_loader = train_loaders[0]
        val_loaders = create_dataloaders(cfg.data, 'VALID')
        for entry in (cfg.data.get('VALID') or []):
            if entry and (entry.get('dataloader') or {}).get('drop_last', False):
                raise ValueError('DataLoader parameters `drop_last` must be False in valid phase.')
        max_epochs = tr.get('max_epochs')
        max_steps = tr.get('max_steps', -1) or -1
        if max_epochs is None:
            max_epochs = 1000 if max_steps == -1 else 10 ** 9
        every = tr.get('check_val_every_n_epoch', 1) or 1
        log_every = tr.get('log_every_n_steps') or 50
        train_metrics = len(self.metrics_manager.phase2metrics[Phase.TRAIN.name]) > 0
        t0 = time.time()
        while self.current_epoch < max_epochs and not self.should_stop:
            for c in self.callbacks:
                c.on_train_epoch_start(self)
            if hasattr(train_loader.sampler, 'set_epoch'):
                train_loader.sampler.set_epoch(self.current_epoch)
            n = _limit(len(train_loader), tr.get('limit_train_batches'))
            loss_sum, steps, tag_sums = None, 0, {}
            for i, batch in enumerate(train_loader):
                if i >= n:
                    break
                loss = self.loop.train_step(batch)                          # device scalar, no sync
                loss_sum = loss.float().clone() if loss_sum is None else loss_sum + loss
                for tag, value in getattr(self.loop, 'tagged', {}).items():       # train/<tag>, tasks/base.py:163-173
                    tag_sums[tag] = value.float().clone() if tag not in tag_sums else tag_sums[tag] + value
                steps += 1
                self.global_step += 1
                if train_metrics:
                    self.metrics_manager.update(Phase.TRAIN, **self.task.last_output)
                if self.scheduler:
                    self.scheduler.step_end()
                if self.global_step % log_every == 0:
                    self._log({'loss': float(loss), 'lr': self.loop.optimizer.lr})
                if max_steps != -1 and self.global_step >= max_steps:
                    self.should_stop = True
                    break
            logs = {'train/loss': self._mean_over_ranks(loss_sum / max(steps, 1))} if steps else {}
            logs.update({f'train/{t}': self._mean_over_ranks(v / max(steps, 1)) for t, v in tag_sums.items()})
            logs.update(self.metrics_manager.on_epoch_end(Phase.TRAIN))
            logs['step'] = float(self.current_epoch)
            self._log(logs)
            # epoch-interval schedulers tick before validation so that a checkpoint written at on_validation_end holds
            # the lr of the NEXT epoch (what resuming needs); ReduceLROnPlateau waits for the validation metrics
            if self.scheduler and not self.scheduler.on_plateau:
                self.scheduler.epoch_end(dict(self.logged))
            for c in self.callbacks:
                c.on_train_epoch_end(self, dict(self.logged))
            if val_loaders and (self.current_epoch + 1) % every == 0:
                vlogs = self._evaluate(Phase.VALID, val_loaders, tr.get('limit_val_batches'))
                self._log(vlogs)
                for c in self.callbacks:
                    c.on_validation_end(self, dict(self.logged))
            if self.scheduler and self.scheduler.on_plateau:
                self.scheduler.epoch_end(dict(self.logged))
            if self.rank == 0:
                shown = {k: round(v, 5) for k, v in self.logged.items() if k != 'step'}
                print(f'[torchok_b200] epoch {self.current_epoch} step {self.global_step} '
                      f'{time.time() - t0:.1f}s {shown}', flush=True)
            self.current_epoch += 1
        for c in self.callbacks:
            c.teardown(self)
        self.logger.close()
        return dict(self.logged)

    def _start_inference(self):
        lc = self.cfg.task.get('load_checkpoint')
        if lc:
            load_checkpoint(self.task, **dict(lc))                          # on_test_start / on_predict_start
        if self.cfg.get('resume_path'):
            self._resume(self.cfg['resume_path'], restore_training_state=False)

    def test(self):
        self._start_inference()
        loaders = create_dataloaders(self.cfg.data, 'TEST')
        logs = self._evaluate(Phase.TEST, loaders, self.trainer.get('limit_test_batches'))
        self._log(logs)
        self.logger.close()
        return logs

    def predict(self):
        self._start_inference()
        loaders = create_dataloaders(self.cfg.data, 'PREDICT')
        self.task.eval()
        outputs = []
        with torch.no_grad():
            for loader in loaders:
                n = _limit(len(loader), self.trainer.get('limit_predict_batches'))
                for i, batch in enumerate(loader):
                    if i >= n:
                        break
                    out = self.task.predict_step(self._to_device(batch), i)
                    outputs.append({k: (v.detach().cpu() if torch.is_tensor(v) else v) for k, v in out.items()})
        return outputs

    def run(self, mode='train'):
        if mode == 'train':
            return self.fit()
        if mode == 'test':
            return self.test()
        if mode == 'predict':
            return self.predict()
        raise ValueError(f'Main function error. Entrypoint with name <{mode}> does not support, please use '
                         f'the following entrypoints - [train, test, predict].')
