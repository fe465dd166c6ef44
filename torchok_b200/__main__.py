"""`python -m torchok_b200 -cp <config dir> -cn <config name> [key=value ...]` — the reference's front door
(torchok/__main__.py:13-55: hydra `--config-path/-cp`, `--config-name/-cn`, dotted overrides, `mode` = train | test |
predict) without hydra.  Under torchrun (one process per GPU) the NCCL process group is created here from the
environment; data parallelism then follows engine.StreamLoop (one all-reduce per gradient bucket)."""
import argparse
import os
import sys

import torch


def find_config(config_path, config_name):
    """hydra resolves a relative -cp against the directory of the calling module (README: `-cp ../examples/configs`);
    the current directory is tried first, then the package directory."""
    name = config_name if config_name.endswith(('.yaml', '.yml')) else None
    roots = [config_path] if os.path.isabs(config_path) else \
        [os.path.join(os.getcwd(), config_path), os.path.join(os.path.dirname(os.path.abspath(__file__)), config_path)]
    tried = []
    for root in roots:
        for cand in ([name] if name else [config_name + '.yaml', config_name + '.yml']):
            path = os.path.normpath(os.path.join(root, cand))
            tried.append(path)
            if os.path.isfile(path):
                return path
    raise FileNotFoundError(f"Cannot find primary config '{config_name}'. Tried: {', '.join(tried)}")


def parse_args(argv=None):
    ap = argparse.ArgumentParser(prog='python -m torchok_b200', description=__doc__,
                                 formatter_class=argparse.RawDescriptionHelpFormatter)
    ap.add_argument('-cp', '--config-path', required=True, help='directory holding the YAML config')
    ap.add_argument('-cn', '--config-name', required=True, help='config file name (with or without .yaml)')
    ap.add_argument('overrides', nargs='*', help='hydra-style overrides: a.b.c=value, +new.key=value, mode=test')
    return ap.parse_args(argv)


def entrypoint(argv=None):
    args = parse_args(argv)
    bad = [o for o in args.overrides if '=' not in o]
    if bad:
        raise SystemExit(f'overrides must look like key=value, got {bad}')
    mode = 'train'
    overrides = []
    for o in args.overrides:
        key, _, value = o.lstrip('+').partition('=')
        if key in ('mode', 'entrypoint'):
            mode = value
        else:
            overrides.append(o)
    from .constructor.config import load_config
    from .runner import Runner
    cfg = load_config(find_config(args.config_path, args.config_name), overrides)
    if cfg.get('mode'):
        mode = cfg.pop('mode')
    if int(os.environ.get('WORLD_SIZE', '1')) > 1 and not torch.distributed.is_initialized():
        os.environ.setdefault('NCCL_IB_DISABLE', '1')      # NVLink / NVSwitch only
        torch.cuda.set_device(int(os.environ.get('LOCAL_RANK', 0)))
        torch.distributed.init_process_group('nccl')
    torch.set_float32_matmul_precision('highest')          # __main__.py:36 (only host-side torch ops are affected)
    result = Runner(cfg).run(mode)
    if torch.distributed.is_initialized():
        torch.distributed.barrier()
        torch.distributed.destroy_process_group()
    return result


if __name__ == '__main__':
    entrypoint(sys.argv[1:])
