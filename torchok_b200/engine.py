"""The CUDA-stream step loop that replaces Lightning's fit loop for the hot path.

Reference behaviour replaced (torchok/tasks/base.py:125-133 `training_step`, :163-173 per-step all_gather, Lightning's
automatic optimisation + DDP): per step  zero_grad -> forward_with_gt -> JointLoss -> backward -> (gradient
all-reduce) -> optimizer.step.  Here:

  * `ParamArena` re-homes every trainable parameter into ONE flat fp32 buffer (memory order preserved, so conv
    weights stay [K][R][S][C]), with a matching flat fp32 gradient arena the wgrad kernels accumulate into and a flat
    bf16 shadow arena the conv/linear kernels read.  state_dict keys / shapes are unchanged.
  * `ArenaSGD` / `ArenaAdam` run torch.optim.SGD / Adam(W) arithmetic over the whole arena in one kernel that also
    refreshes the bf16 shadow and clears the consumed gradient (tok_sgd_step_dev / tok_adam_step_dev) — one parameter
    group with shared hyper-parameters, which is what the reference's Constructor builds from the example YAMLs
    (constructor.py:151-156).
  * data parallelism = one process per GPU; gradients are summed with ONE NCCL all-reduce per gradient bucket
    (contiguous slices of the gradient arena) issued on a communication stream as soon as the backward pass has
    produced every gradient of the bucket; BatchNorm statistics stay local (sync_batchnorm default False,
    config_structure.py:170).  The reference's per-step loss all_gather (base.py:170) is dropped: the loss stays on
    the device and is reduced only when read.
  * the whole step (forward, loss, backward, all-reduce, optimizer) is captured once into a CUDA graph and replayed.
"""
import math
import os

import torch
import torch.distributed as dist
import torch.nn as nn

from . import kernels as K
from ._lib import lib, tokPeerArenas

F32, BF16 = torch.float32, torch.bfloat16
_ALIGN = 64  # elements; keeps every bf16 shadow 128-byte aligned for TMA


def _round_up(n, a):
    return (n + a - 1) // a * a


def plan_buckets(numels, bucket_mb=32.0, align=_ALIGN):
    """Cut the flat gradient arena into all-reduce buckets: contiguous slices that start and end at parameter boundaries
    (every parameter occupies `round_up(numel, align)` elements), closed as soon as they reach `bucket_mb` MiB of fp32.
    Returns [(begin, end, [parameter indices])]; the slices tile [0, total) without gaps or overlap.  Pure host logic
    (covered by the gloo world-size-2 test on CPU)."""
    limit = max(1, int(bucket_mb * (1 << 20) / 4))
    buckets, begin, members, off = [], 0, [], 0
    for i, n in enumerate(numels):
        end = off + _round_up(n, align)
        members.append(i)
        if end - begin >= limit or i == len(numels) - 1:
            buckets.append((begin, end, members))
            begin, members = end, []
        off = end
    return buckets


class ParamArena:
    def __init__(self, module, bucket_mb=32.0, alloc=None):
        """`alloc(numel, dtype) -> zero-filled flat CUDA tensor` lets the caller place the three arenas in memory it
        owns (PeerExchange: CUDA-IPC allocations the other ranks can map); default: torch's caching allocator."""
        params, seen = [], set()
        for p in module.parameters():
            if p.requires_grad and id(p) not in seen:
                seen.add(id(p))
                params.append(p)
        if not params:
            raise ValueError('ParamArena: module has no trainable parameters')
        dev = params[0].device
        K.require_cuda(params[0], 'parameter')
        offs, total = [], 0
        for p in params:
            if p.dtype != F32:
                raise TypeError('ParamArena expects fp32 master parameters')
            if not torch.ops.aten.is_non_overlapping_and_dense(p):
                p.data = p.data.contiguous()
            offs.append(total)
            total += _round_up(p.numel(), _ALIGN)
        self.params, self.offsets, self.numel = params, offs, total
        if alloc is None:
            def alloc(n, dtype):
                return torch.zeros(n, dtype=dtype, device=dev)
        self.master = alloc(total, F32)
        self.grad = alloc(total, F32)
        self.shadow = alloc(total, BF16)
        for p, off in zip(params, offs):
            shape, stride = tuple(p.shape), tuple(p.stride())
            view = torch.as_strided(self.master, shape, stride, off)
            view.copy_(p.data)
            p.data = view
            p.grad = torch.as_strided(self.grad, shape, stride, off)
            p._tok_shadow = torch.as_strided(self.shadow, shape, stride, off)
        self.refresh_shadow()
        # gradient buckets: contiguous arena slices, cut at parameter boundaries
        self.buckets = plan_buckets([p.numel() for p in params], bucket_mb)
        for b, (_, _, members) in enumerate(self.buckets):
            for i in members:
                params[i]._tok_bucket = (self, b)
        self._pending = [len(m) for _, _, m in self.buckets]
        self._launched = [False] * len(self.buckets)
        self.reducer = None

    def refresh_shadow(self):
        K.cast_bf16(self.master, self.shadow)

    def zero_grad(self):
        self.grad.zero_()

    # -- bucket readiness (called from the backward Functions through kernels.grad_ready) --------------------------
    def begin_step(self):
        self._pending = [len(m) for _, _, m in self.buckets]
        self._launched = [False] * len(self.buckets)

    def ready(self, b):
        self._pending[b] -= 1
        if self._pending[b] == 0 and self.reducer is not None and not self._launched[b]:
            self._launched[b] = True
            K.join_side()          # weight gradients launched on the side stream belong to this bucket
            self.reducer.launch(b)

    def finish(self):
        """Reduce whatever has not been reduced yet (parameters whose backward did not announce itself)."""
        K.join_side()
        if self.reducer is not None:
            for b in range(len(self.buckets)):
                if not self._launched[b]:
                    self._launched[b] = True
                    self.reducer.launch(b)
            self.reducer.join()


class BucketAllReduce:
    """One NCCL all-reduce (sum) per gradient bucket, on a side stream, overlapped with the rest of backward."""

    def __init__(self, arena, group=None):
        self.arena, self.group = arena, group
        self.world = dist.get_world_size(group)
        self.comm_stream = torch.cuda.Stream()
        self._events = []
        arena.reducer = self

    def launch(self, b):
        begin, end, _ = self.arena.buckets[b]
        cur = torch.cuda.current_stream()
        ev = torch.cuda.Event()
        ev.record(cur)
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            dist.all_reduce(self.arena.grad[begin:end], op=dist.ReduceOp.SUM, group=self.group)

    def join(self):
        torch.cuda.current_stream().wait_stream(self.comm_stream)


class _DevMem:
    """A raw device allocation seen through the CUDA array interface (zero-copy torch.as_tensor)."""

    def __init__(self, ptr, nbytes):
        self.ptr, self.nbytes = ptr, nbytes
        self.__cuda_array_interface__ = {'shape': (nbytes,), 'typestr': '|u1', 'data': (ptr, False), 'version': 2}


class PeerExchange:
    """Gradient exchange over NVLink peer memory fused with the optimizer step (tok_peer_step, csrc/tok_comm.cu): per
    gradient bucket ONE kernel per rank does reduce-scatter (peer loads) -> SGD / Adam(W) on the owned slice ->
    all-gather of the new fp32 master + bf16 shadow values (peer stores).  Replaces DDP's all-reduce + optimizer.step
    (`trainer.strategy: ddp`, torchok/constructor/config_structure.py:137-140).  torch.distributed is used only to
    exchange the 64-byte IPC handles.  Plain kernel launches: the step stays CUDA-graph capturable at any world size."""

    def __init__(self, device, group=None):
        import ctypes as C
        self.C, self.L = C, lib()
        self.group, self.device = group, device
        self.world, self.rank = dist.get_world_size(group), dist.get_rank(group)
        if self.world > 8:
            raise RuntimeError('PeerExchange covers one NVSwitch domain (<= 8 ranks)')
        self._local, self._opened = [], []
        self.comm_stream = torch.cuda.Stream()
        self._first = True
        self.table = None

    def _alloc_raw(self, nbytes):
        C = self.C
        ptr, handle = C.c_void_p(), C.create_string_buffer(64)
        self.L.tok_ipc_alloc(C.c_size_t(nbytes), C.byref(ptr), handle)
        self._local.append(ptr.value)
        return ptr.value, handle.raw

    def alloc(self, numel, dtype):
        """Arena allocator handed to ParamArena: IPC-exportable, zero-filled."""
        nbytes = _round_up(numel * torch.empty((), dtype=dtype).element_size(), 256)
        ptr, handle = self._alloc_raw(nbytes)
        t = torch.as_tensor(_DevMem(ptr, nbytes), device=self.device).view(dtype)[:numel]
        assert t.data_ptr() == ptr
        self._handles = getattr(self, '_handles', []) + [handle]
        return t

    def connect(self, arena):
        """Exchange the handles of (master, grad, shadow, flags) and build the pointer table of tok_peer_step."""
        C = self.C
        fptr, fh = self._alloc_raw(int(self.L.tok_peer_flag_bytes()))
        mine = self._handles[-3:] + [fh]
        ptrs = [arena.master.data_ptr(), arena.grad.data_ptr(), arena.shadow.data_ptr(), fptr]
        everyone = [None] * self.world
        dist.all_gather_object(everyone, mine, group=self.group)
        tab = tokPeerArenas()
        tab.world, tab.rank = self.world, self.rank
        for r in range(self.world):
            for j, field in enumerate(('master', 'grad', 'shadow', 'flags')):
                if r == self.rank:
                    val = ptrs[j]
                else:
                    p = C.c_void_p()
                    self.L.tok_ipc_open(everyone[r][j], C.byref(p))
                    self._opened.append(p.value)
                    val = p.value
                getattr(tab, field)[r] = val
        self.table, self.arena = tab, arena
        arena.reducer = self
        dist.barrier(group=self.group)

    # -- reducer interface used by ParamArena.ready / finish ---------------------------------------------------------
    def begin_step(self):
        self._first = True

    def launch(self, b):
        begin, end, _ = self.arena.buckets[b]
        ev = torch.cuda.Event()
        ev.record(torch.cuda.current_stream())
        self.comm_stream.wait_event(ev)
        with torch.cuda.stream(self.comm_stream):
            self.optimizer.peer_step(self.table, b, begin, end, self._first)
        self._first = False

    def join(self):
        torch.cuda.current_stream().wait_stream(self.comm_stream)

    def close(self):
        torch.cuda.synchronize()
        if dist.is_initialized():
            dist.barrier(group=self.group)
        for p in self._opened:
            self.L.tok_ipc_close(p)
        self._opened = []


class _ArenaOptimizer:
    def __init__(self, arena, lr):
        self.arena = arena
        dev = arena.master.device
        self.lr_dev = torch.tensor([float(lr)], dtype=F32, device=dev)
        self.step_dev = torch.zeros(1, dtype=torch.int32, device=dev)
        self._lr = float(lr)
        self.grad_scale = 1.0

    def set_param_multipliers(self, mults):
        """`mults`: {parameter: (lr_mult, decay_mult)} from constructor.paramwise.paramwise_multipliers.  Builds the
        per-parameter segment table the step kernels look multipliers up in (one segment per arena parameter)."""
        a = self.arena
        lr_m = [float(mults.get(p, (1., 1.))[0]) for p in a.params]
        wd_m = [float(mults.get(p, (1., 1.))[1]) for p in a.params]
        if all(v == 1. for v in lr_m) and all(v == 1. for v in wd_m):
            self.segs = None
            return
        dev = a.master.device
        self.segs = (torch.tensor(a.offsets, dtype=torch.int32, device=dev),
                     torch.tensor(lr_m, dtype=F32, device=dev), torch.tensor(wd_m, dtype=F32, device=dev))
        # per-parameter step counts (torch.optim.Adam's state['step']): a parameter that has been optimised keeps its
        # count while frozen; one that never was starts at 0, so its bias correction begins at 1 when it is thawed
        if getattr(self, 'seg_steps', None) is None or self.seg_steps.numel() != len(a.params):
            self.seg_steps = self.step_dev.expand(len(a.params)).clone()

    def _seg_args(self):
        if getattr(self, 'segs', None) is None:
            return None, None, None, 0
        b, l, w = self.segs
        return K._p(b), K._p(l), K._p(w), int(b.numel())

    def _seg_args5(self):
        b, l, w, n = self._seg_args()
        return b, l, w, (K._p(self.seg_steps) if n else None), n

    def _full_state(self, flat):
        """Under the peer-fused exchange every rank holds the momentum / moment values of ITS slice of each bucket only
        (ZeRO-1): assemble the full buffer with one all-reduce.  COLLECTIVE — every rank must call state_dict()."""
        peer = getattr(a := self.arena, 'reducer', None)
        if not isinstance(peer, PeerExchange):
            return flat
        full = torch.zeros_like(flat)
        for begin, end, _ in a.buckets:
            per = ((end - begin) // 4 + peer.world - 1) // peer.world * 4
            lo = begin + per * peer.rank
            hi = min(end, lo + per)
            if hi > lo:
                full[lo:hi] = flat[lo:hi]
        dist.all_reduce(full, group=peer.group)
        return full

    @property
    def lr(self):
        return self._lr

    @lr.setter
    def lr(self, value):
        """Schedulers write here; the value reaches the (possibly graph-captured) kernel through device memory."""
        self._lr = float(value)
        self.lr_dev.fill_(self._lr)

    @property
    def param_groups(self):  # enough of the torch.optim surface for lr schedulers that only touch 'lr'
        return [{'lr': self._lr, 'params': self.arena.params}]

    def zero_grad(self, set_to_none=False):
        self.arena.zero_grad()

    # -- checkpointing: state is stored per parameter NAME so that it survives a different arena layout -------------
    _STATE_BUFFERS = ('buf', 'exp_avg', 'exp_avg_sq')

    def state_dict(self, module):
        """{'step', 'lr', 'state': {buffer: {parameter name: tensor}}} for the parameters of `module` in the arena."""
        a = self.arena
        names = {id(p): n for n, p in module.named_parameters()}
        state = {}
        for key in self._STATE_BUFFERS:
            flat = getattr(self, key, None)
            if flat is None:
                continue
            flat = self._full_state(flat)
            state[key] = {names[id(p)]: flat[off:off + p.numel()].detach().cpu().clone()
                          for p, off in zip(a.params, a.offsets) if id(p) in names}
        out = {'step': int(self.step_dev.item()), 'lr': self._lr, 'state': state}
        if getattr(self, 'segs', None) is not None and getattr(self, 'seg_steps', None) is not None:
            steps = self.seg_steps.cpu().tolist()
            out['param_steps'] = {names[id(p)]: int(t) for p, t in zip(a.params, steps) if id(p) in names}
        return out

    def load_state_dict(self, state, module):
        a = self.arena
        names = {id(p): n for n, p in module.named_parameters()}
        for key, per_name in state.get('state', {}).items():
            flat = getattr(self, key, None)
            if flat is None:
                continue
            for p, off in zip(a.params, a.offsets):
                src = per_name.get(names.get(id(p)))
                if src is not None:
                    if src.numel() != p.numel():
                        raise ValueError(f'optimizer state {key} of {names[id(p)]}: {src.numel()} values for a '
                                         f'parameter of {p.numel()}')
                    flat[off:off + p.numel()].copy_(src.reshape(-1))
        self.step_dev.fill_(int(state.get('step', 0)))
        self.lr = state.get('lr', self._lr)
        per = state.get('param_steps')
        if getattr(self, 'seg_steps', None) is not None:
            vals = [int((per or {}).get(names.get(id(p)), state.get('step', 0))) for p in a.params]
            self.seg_steps.copy_(torch.tensor(vals, dtype=torch.int32))


class ArenaSGD(_ArenaOptimizer):
    """torch.optim.SGD (registered in torchok/optim/optimizers/__init__.py:9-19) over the flat arena."""

    def __init__(self, arena, lr, momentum=0.0, dampening=0.0, weight_decay=0.0, nesterov=False, **other):
        super().__init__(arena, lr)
        _check_optimizer_kwargs('SGD', other)
        self.momentum, self.dampening, self.weight_decay, self.nesterov = momentum, dampening, weight_decay, nesterov
        self.buf = torch.zeros_like(arena.master) if momentum != 0 else None

    def step(self):
        a = self.arena
        lib().tok_sgd_step_dev_groups(a.numel, K._p(a.master), K._p(a.grad), K._p(self.buf), K._p(a.shadow),
                                      K._p(self.lr_dev), K._p(self.step_dev), self.momentum, self.weight_decay,
                                      self.dampening, int(self.nesterov), self.grad_scale, 1, *self._seg_args(), K._st())

    def peer_step(self, table, bucket, begin, end, first):
        import ctypes as C
        lib().tok_peer_step(C.byref(table), bucket, begin, end, 0, K._p(self.buf), None, K._p(self.lr_dev),
                            K._p(self.step_dev), self.momentum, self.weight_decay, self.dampening, 0.0,
                            int(self.nesterov), self.grad_scale, *self._seg_args5(), int(first), K._st())


class ArenaAdam(_ArenaOptimizer):
    """torch.optim.Adam / AdamW (amsgrad off) over the flat arena."""

    def __init__(self, arena, lr=1e-3, betas=(0.9, 0.999), eps=1e-8, weight_decay=0.0, decoupled=False, **other):
        super().__init__(arena, lr)
        _check_optimizer_kwargs('Adam', other)
        self.betas, self.eps, self.weight_decay, self.decoupled = tuple(betas), eps, weight_decay, decoupled
        self.exp_avg = torch.zeros_like(arena.master)
        self.exp_avg_sq = torch.zeros_like(arena.master)

    def step(self):
        a = self.arena
        lib().tok_adam_step_dev_groups(a.numel, K._p(a.master), K._p(a.grad), K._p(self.exp_avg),
                                       K._p(self.exp_avg_sq), K._p(a.shadow), K._p(self.lr_dev), K._p(self.step_dev),
                                       self.betas[0], self.betas[1], self.eps, self.weight_decay, int(self.decoupled),
                                       self.grad_scale, 1, *self._seg_args5(), K._st())

    def peer_step(self, table, bucket, begin, end, first):
        import ctypes as C
        lib().tok_peer_step(C.byref(table), bucket, begin, end, 1, K._p(self.exp_avg), K._p(self.exp_avg_sq),
                            K._p(self.lr_dev), K._p(self.step_dev), self.betas[0], self.betas[1], self.eps,
                            self.weight_decay, int(self.decoupled), self.grad_scale, *self._seg_args5(), int(first),
                            K._st())


_NOOP_OPT_KEYS = ('foreach', 'fused', 'capturable', 'differentiable')


def _check_optimizer_kwargs(name, other):
    """torch.optim keys the arena kernels do not implement must not be dropped silently (they change the arithmetic)."""
    for k, v in other.items():
        if k in _NOOP_OPT_KEYS or (k in ('amsgrad', 'maximize') and not v):
            continue
        raise NotImplementedError(f'{name}: optimizer parameter {k}={v!r} is not implemented by the arena step kernels')


def build_optimizer(arena, name, params, module=None, paramwise_cfg=None):
    """`paramwise_cfg` (optimization[i].optimizer.paramwise_cfg, torchok/constructor/constructor.py:145-156) needs the
    `module` whose tree the rules are evaluated on (the task)."""
    params = dict(params or {})
    if name == 'SGD':
        opt = ArenaSGD(arena, **params)
    elif name == 'Adam':
        opt = ArenaAdam(arena, decoupled=False, **params)
    elif name == 'AdamW':
        params.setdefault('weight_decay', 1e-2)
        opt = ArenaAdam(arena, decoupled=True, **params)
    else:
        raise NotImplementedError(f'optimizer {name}: the arena step kernels cover SGD, Adam and AdamW')
    if paramwise_cfg:
        if module is None:
            raise ValueError('paramwise_cfg needs the module tree it is evaluated on')
        from .constructor.paramwise import paramwise_multipliers
        opt.set_param_multipliers(paramwise_multipliers(module, dict(paramwise_cfg)))
    return opt


class StreamLoop:
    """Drives `task.training_step` as a captured CUDA graph.

        loop = StreamLoop(task, optimizer=dict(name='SGD', params=dict(lr=0.1, momentum=0.9, weight_decay=1e-4)))
        loss = loop.train_step({'image': pinned_cpu_or_cuda_tensor, 'target': ...})   # device scalar

    With torch.distributed initialised (NCCL), gradients are averaged over ranks bucket by bucket.
    """

    def __init__(self, task, optimizer=None, use_graph=True, bucket_mb=32.0, warmup=3, distributed=True):
        self.task = task
        self.device = next(task.parameters()).device
        K.require_cuda(next(task.parameters()), 'task')
        lib()  # fail now, loudly, if the extension is missing
        if optimizer is None:
            opt_cfg = (task.hparams.get('optimization') or [None])[0]
            if opt_cfg is None:
                raise ValueError('StreamLoop needs an optimizer (argument or hparams.optimization[0].optimizer)')
            optimizer = opt_cfg['optimizer']
        self.world = dist.get_world_size() if distributed and dist.is_available() and dist.is_initialized() else 1
        # gradient exchange: 'none' (one GPU), 'peer-fused' (NVLink peer memory + optimizer in one kernel per bucket,
        # the default for 2..8 ranks), 'nccl' (torch.distributed all-reduce per bucket, then the arena optimizer;
        # TOK_DDP=nccl or when CUDA IPC / peer access is unavailable)
        self.exchange, self.peer = 'none', None
        if self.world > 1:
            self.exchange = 'nccl'
            if os.environ.get('TOK_DDP', 'peer') == 'peer' and self.world <= 8 and dist.get_backend() == 'nccl':
                try:
                    self.peer = PeerExchange(self.device)
                    self.arena = ParamArena(task, bucket_mb=bucket_mb, alloc=self.peer.alloc)
                    self.peer.connect(self.arena)
                    self.exchange = 'peer-fused'
                except Exception as e:   # no IPC in this container / no peer access: say so and use NCCL
                    import warnings
                    warnings.warn(f'torchok_b200: peer-memory gradient exchange unavailable ({type(e).__name__}: {e}); '
                                  f'using the NCCL all-reduce path')
                    if self.peer is not None and getattr(self, 'arena', None) is not None:
                        raise   # arenas already re-homed into IPC memory: do not continue half-configured
                    self.peer = None
        if self.peer is None:
            self.arena = ParamArena(task, bucket_mb=bucket_mb)
        self.optimizer = build_optimizer(self.arena, optimizer['name'], optimizer.get('params'), task,
                                         optimizer.get('paramwise_cfg'))
        if self.world > 1:
            if self.peer is not None:
                self.peer.optimizer = self.optimizer
            else:
                BucketAllReduce(self.arena)
            self.optimizer.grad_scale = 1.0 / self.world
            with torch.no_grad():  # replicas start from rank 0's weights (DDP's initial broadcast)
                dist.broadcast(self.arena.master, 0)
                for b in task.buffers():
                    if b.is_cuda and b.is_floating_point():
                        dist.broadcast(b, 0)
            self.arena.refresh_shadow()
        self.use_graph = use_graph and os.environ.get('TOK_NO_GRAPH', '0') != '1'
        if self.exchange == 'nccl' and os.environ.get('TOK_GRAPH_DDP', '0') != '1':
            # Capturing torch.distributed's NCCL all-reduce needs capture_error_mode='thread_local' and
            # TORCH_NCCL_ASYNC_ERROR_HANDLING=0; validated at 2 ranks only, so it stays opt-in (TOK_GRAPH_DDP=1).  The
            # default peer-fused exchange is plain kernel launches and is always captured.
            self.use_graph = False
        self.warmup = warmup
        self.graph = None
        self.static = None
        self.loss = None
        self.tagged = {}      # tagged JointLoss values of the last step (device scalars), logged as train/<tag>
        self.outputs = None
        self.steps = 0
        self.stream = torch.cuda.Stream()
        self._bns = [m for m in task.modules() if isinstance(m, nn.BatchNorm2d) and hasattr(m, '_pending_batches')]

    # ------------------------------------------------------------------------------------------------------------
    def _eager_step(self, batch):
        self.arena.begin_step()
        if self.peer is not None:
            self.peer.begin_step()
        out = self.task.training_step(batch)
        K.chain_flush()   # close the BatchNorm sum chain inside the step: a captured step must not depend on what ran before it
        out['loss'].backward()
        self.arena.finish()
        if self.peer is None:      # peer-fused: the optimizer ran inside the exchange kernels, bucket by bucket
            self.optimizer.step()
        return out

    def close(self):
        """Release the peer mappings (all ranks together); the loop must not be used afterwards."""
        if self.peer is not None:
            self.graph = None
            self.peer.close()

    def _stage(self, batch):
        """Copy a batch (pinned host or device tensors) into the static device buffers the graph reads."""
        if self.static is None:
            self.static = {k: torch.empty(v.shape, dtype=v.dtype, device=self.device)
                           for k, v in batch.items() if torch.is_tensor(v)}
        for k, dst in self.static.items():
            src = batch[k]
            if src.data_ptr() != dst.data_ptr():
                dst.copy_(src, non_blocking=True)
        return self.static

    def train_step(self, batch):
        self.task.train()
        odd_shape = self.static is not None and any(
            torch.is_tensor(batch.get(k)) and batch[k].shape != v.shape for k, v in self.static.items())
        if not self.use_graph or odd_shape:   # e.g. the short last batch of an epoch: run it outside the graph
            dev_batch = {k: (v.to(self.device, non_blocking=True) if torch.is_tensor(v) else v)
                         for k, v in batch.items()}
            out = self._eager_step(dev_batch)
            self.steps += 1
            self.loss = out['loss'].detach()
            self.tagged = {k: v.detach() for k, v in out.items() if k != 'loss' and torch.is_tensor(v)}
            return self.loss
        cur = torch.cuda.current_stream()
        if self.graph is None:
            # Warm-up steps (needed before capture: lazy kernel attributes, allocator pools, NCCL channels) run on
            # a snapshot of the training state which is restored afterwards, so the first replay below IS step 1.
            static = self._stage(batch)
            snap = self._snapshot()
            self.stream.wait_stream(cur)
            with torch.cuda.stream(self.stream):
                for _ in range(self.warmup):
                    self._eager_step(static)
            cur.wait_stream(self.stream)
            torch.cuda.synchronize()
            self._restore(snap)
            g = torch.cuda.CUDAGraph()
            # with NCCL in the step, other threads (the process-group watchdog) touch the CUDA API during capture
            mode = 'thread_local' if self.world > 1 else 'global'
            with torch.cuda.graph(g, stream=self.stream, capture_error_mode=mode):
                out = self._eager_step(static)
                self._graph_loss = out['loss'].detach()
                self._graph_tagged = {k: v.detach() for k, v in out.items() if k != 'loss' and torch.is_tensor(v)}
                self._graph_output = getattr(self.task, 'last_output', None)   # static forward outputs (metrics)
            self.graph = g
            self._restore(snap)
            del snap
        self._stage(batch)
        self.graph.replay()
        self.steps += 1
        for m in self._bns:
            if m.track_running_stats:
                m._pending_batches += 1
        self.loss = self._graph_loss
        self.tagged = self._graph_tagged
        if self._graph_output is not None:
            self.task.last_output = self._graph_output
        return self.loss

    def reset_graph(self):
        """Drop the captured step; the next `train_step` warms up and captures again.  Needed whenever what a step
        launches changes: the frozen set (callbacks.FreezeUnfreeze), the multiplier table, the batch shape."""
        self.graph = None
        self.static = None

    def _state_tensors(self):
        opt = self.optimizer
        ts = [self.arena.master, self.arena.grad, opt.step_dev]
        ts += [t for t in (getattr(opt, 'buf', None), getattr(opt, 'exp_avg', None), getattr(opt, 'exp_avg_sq', None),
                           getattr(opt, 'seg_steps', None)) if t is not None]
        ts += [b for b in self.task.buffers() if b.is_cuda]
        return ts

    def _snapshot(self):
        return [t.clone() for t in self._state_tensors()], [m._pending_batches for m in self._bns]

    def _restore(self, snap):
        tensors, pending = snap
        with torch.no_grad():
            for dst, src in zip(self._state_tensors(), tensors):
                dst.copy_(src)
        for m, n in zip(self._bns, pending):
            m._pending_batches = n
        self.arena.refresh_shadow()
        torch.cuda.synchronize()

    def set_lr(self, lr):
        self.optimizer.lr = lr
