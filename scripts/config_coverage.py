"""Which names of the reference's example configs resolve in this package's registries.

    python scripts/config_coverage.py /root/reference/examples/configs        # prints a markdown table

Every YAML is loaded with torchok_b200.load_config (anchors, ${oc.env:…}, ${now:…}, ${a.b}) and each `name:` is
looked up in the registry its position in the config implies.  Used to keep DESIGN §7's drop-in table honest."""
import glob
import os
import sys

os.environ.setdefault('HOME', '/root')
import torchok_b200 as tb  # noqa: E402


def names_of(cfg):
    out = []
    t = cfg.get('task') or {}
    out.append(('TASKS', t.get('name')))
    p = t.get('params') or {}
    for key, reg in (('backbone_name', 'BACKBONES'), ('neck_name', 'NECKS'), ('pooling_name', 'POOLINGS'),
                     ('head_name', 'HEADS')):
        if p.get(key):
            out.append((reg, p[key]))
    for loss in ((cfg.get('joint_loss') or {}).get('losses') or []):
        out.append(('LOSSES', loss['name']))
    for o in cfg.get('optimization') or []:
        out.append(('OPTIMIZERS', o['optimizer']['name']))
        if o.get('scheduler'):
            out.append(('SCHEDULERS', o['scheduler']['name']))
    for m in cfg.get('metrics') or []:
        out.append(('METRICS', m['name']))
    for c in cfg.get('callbacks') or []:
        out.append(('CALLBACKS', c['name']))
    for phase, entries in (cfg.get('data') or {}).items():
        for e in entries or []:
            if not e:
                continue
            out.append(('DATASETS', e['dataset']['name']))
            for t in (e['dataset'].get('transform') or []) + (e['dataset'].get('augment') or []):
                out.append(('TRANSFORMS', t['name']))
            if e.get('sampler'):
                out.append(('SAMPLERS', e['sampler']['name']))
    return out


def main(root):
    print('| config | loads | resolved | missing (registry: name) |')
    print('|---|---|---|---|')
    for path in sorted(glob.glob(os.path.join(root, '**', '*.y*ml'), recursive=True)):
        rel = os.path.relpath(path, root)
        try:
            cfg = tb.load_config(path)
        except Exception as e:  # noqa: BLE001
            print(f'| {rel} | no: {type(e).__name__}: {e} | | |')
            continue
        seen, ok, missing = set(), 0, []
        for reg, name in names_of(cfg):
            if (reg, name) in seen or name is None:
                continue
            seen.add((reg, name))
            if name in getattr(tb, reg):
                ok += 1
            else:
                missing.append(f'{reg}: {name}')
        print(f'| {rel} | yes | {ok}/{len(seen)} | {", ".join(missing) or "—"} |')


if __name__ == '__main__':
    main(sys.argv[1] if len(sys.argv) > 1 else '/root/reference/examples/configs')
