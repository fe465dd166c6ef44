"""GPU parity of the step-loop plumbing (SURVEY §8 row a4 and "next" row N2): the flat-arena optimizers against
torch.optim (the reference's optimizers, torchok/optim/optimizers/__init__.py:9-19) including `paramwise_cfg` param
groups (torchok/constructor/constructor.py:145-251), and CUDA-graph replay against the eager loop."""
import copy

import pytest
import torch
import torch.nn as nn

pytestmark = pytest.mark.gpu


def _toy():
    torch.manual_seed(3)
    return nn.Sequential(nn.Conv2d(8, 16, 3, bias=True), nn.BatchNorm2d(16), nn.Conv2d(16, 16, 3, groups=16),
                         nn.LayerNorm(16), nn.Linear(16, 24)).cuda()


PARAMWISE = dict(bias_lr_mult=2.0, bias_decay_mult=0.0, norm_decay_mult=0.5, dwconv_decay_mult=0.25,
                 custom_keys={'4.weight': dict(lr_mult=0.1, decay_mult=3.0)})


@pytest.mark.parametrize('name,params', [('SGD', dict(lr=0.1, momentum=0.9, weight_decay=1e-2)),
                                         ('SGD', dict(lr=0.05, momentum=0.9, weight_decay=1e-2, nesterov=True)),
                                         ('Adam', dict(lr=1e-2, weight_decay=1e-2)),
                                         ('AdamW', dict(lr=1e-2, weight_decay=5e-2))])
@pytest.mark.parametrize('paramwise', [None, PARAMWISE])
def test_arena_optimizer_matches_torch_optim(name, params, paramwise):
    from torchok_b200 import engine
    from torchok_b200.constructor.paramwise import paramwise_multipliers
    m = _toy()
    ref = copy.deepcopy(m)
    arena = engine.ParamArena(m)
    opt = engine.build_optimizer(arena, name, params, m, paramwise)
    mults = paramwise_multipliers(ref, paramwise)
    groups = []
    for p in ref.parameters():
        lm, dm = mults[p]
        groups.append({'params': [p], 'lr': params['lr'] * lm, 'weight_decay': params['weight_decay'] * dm})
    topt = getattr(torch.optim, name)(groups, **params)
    gen = torch.Generator(device='cuda').manual_seed(5)
    for step in range(4):
        for p, q in zip(m.parameters(), ref.parameters()):
            g = torch.randn(p.shape, device='cuda', generator=gen) * (1 + step)
            p.grad.copy_(g)          # arena views: the kernels read the flat gradient buffer
            q.grad = g.clone()
        opt.step()
        topt.step()
        if step == 1:                # a scheduler writes the base lr; every group scales with it
            opt.lr = params['lr'] * 0.5
            for gr, p in zip(topt.param_groups, ref.parameters()):
                gr['lr'] = params['lr'] * 0.5 * mults[p][0]
    for (n_, p), q in zip(m.named_parameters(), ref.parameters()):
        assert torch.allclose(p, q, rtol=2e-5, atol=2e-6), (n_, float((p - q).abs().max()))
        assert torch.equal(p._tok_shadow, p.detach().to(torch.bfloat16)), n_      # bf16 shadow refreshed by the step
        assert float(p.grad.abs().max()) == 0.0                                    # consumed gradient cleared


def test_stream_loop_graph_replay_matches_eager_loop():
    """StreamLoop: the captured-graph path (warm-up on a snapshot, capture, replay) trains like the eager loop."""
    import torchok_b200 as tb
    from torchok_b200.engine import StreamLoop
    cfg = {'task': {'name': 'ClassificationTask', 'params': {
        'backbone_name': 'resnet18', 'backbone_params': {'pretrained': False, 'in_channels': 3},
        'pooling_name': 'Pooling', 'head_name': 'ClassificationHead', 'head_params': {'num_classes': 10}}},
        'joint_loss': {'losses': [{'name': 'CrossEntropyLoss', 'mapping': {'input': 'prediction', 'target': 'target'}}]},
        'optimization': [{'optimizer': {'name': 'SGD', 'params': {'lr': 0.05, 'momentum': 0.9, 'weight_decay': 1e-4},
                                        'paramwise_cfg': {'bias_decay_mult': 0.0, 'norm_decay_mult': 0.0}}}]}
    torch.manual_seed(1)
    x = torch.randn(32, 3, 32, 32, device='cuda')
    y = torch.randint(0, 10, (32,), device='cuda')
    losses = {}
    for graph in (False, True):
        torch.manual_seed(7)
        c = tb.load_config(cfg)
        task = tb.TASKS.get('ClassificationTask')(c, **c.task.params).cuda()
        loop = StreamLoop(task, use_graph=graph)
        assert loop.optimizer.segs is not None       # paramwise_cfg reached the step kernel
        losses[graph] = [float(loop.train_step({'image': x, 'target': y})) for _ in range(5)]
    print(losses)
    assert losses[True][0] == pytest.approx(losses[False][0], rel=1e-3)   # step 1 starts from the same weights
    for a, b in zip(losses[True], losses[False]):
        assert a == pytest.approx(b, rel=5e-2)
    assert losses[False][-1] < losses[False][0]      # it trains
