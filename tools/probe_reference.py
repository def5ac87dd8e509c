#!/usr/bin/env python3
"""probe_reference.py -- is a real MuJoCo (the reference's physics backend) reachable on THIS machine?

SURVEY.md section 8c's probe order for the un-vendored dependency of /root/reference/src/cassiemujoco.c:521-555
(dlopen of libmujoco210.so / libmujoco210nogl.so):

  1. baseline/_ref/                (a driver-provided install of the reference)
  2. $HOME/.mujoco/mujoco210/      (the location the reference's Makefile:5 and loader :184-185, 536-539 use)
  3. python: import mujoco / mujoco_py / dm_control     (any version; >= 2.1.2 changes defaults, sanity check only)
  4. anything else on the box that looks like MuJoCo: shared objects, headers, wheels in the offline wheelhouse, pip metadata

Prints one JSON document and, when gpurun_out/ exists (GPU box), also writes gpurun_out/probe_reference.json so that the log of the
GPU-box run can be committed under profiles/.  tests/test_mujoco_parity.py uses `find_mujoco()`.

TEST / MEASUREMENT INFRASTRUCTURE: nothing in the product imports this file.
"""
import glob
import importlib
import json
import os
import platform
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _try_import(name):
    try:
        m = importlib.import_module(name)
        return {'found': True, 'version': getattr(m, '__version__', None), 'file': getattr(m, '__file__', None)}
    except Exception as ex:   # ModuleNotFoundError, or a wheel that cannot load its shared objects
        return {'found': False, 'error': '%s: %s' % (type(ex).__name__, ex)}


def _glob_many(patterns, limit=40):
    out = []
    for p in patterns:
        try:
            out += glob.glob(p, recursive=True)
        except Exception:
            pass
        if len(out) >= limit:
            break
    return sorted(set(out))[:limit]


def find_mujoco():
    """-> dict with every probe's outcome and 'usable': None or a short description of what could serve as the physics reference."""
    home = os.path.expanduser('~')
    rep = {'host': platform.node(), 'python': sys.version.split()[0], 'cwd': os.getcwd(), 'probes': {}}
    pr = rep['probes']
    # 1. baseline/_ref
    ref_dir = os.path.join(REPO, 'baseline', '_ref')
    pr['baseline/_ref'] = {'exists': os.path.isdir(ref_dir), 'entries': sorted(os.listdir(ref_dir))[:20] if os.path.isdir(ref_dir) else []}
    # 2. ~/.mujoco/mujoco210 (and any sibling)
    mj_home = os.path.join(home, '.mujoco')
    pr['~/.mujoco'] = {'exists': os.path.isdir(mj_home), 'entries': sorted(os.listdir(mj_home)) if os.path.isdir(mj_home) else [],
                       'libmujoco210': _glob_many([os.path.join(mj_home, 'mujoco210', 'bin', 'libmujoco210*.so')])}
    for var in ('MUJOCO_PY_MUJOCO_PATH', 'MUJOCO_PATH', 'MUJOCO_GL', 'LD_LIBRARY_PATH'):
        pr.setdefault('env', {})[var] = os.environ.get(var)
    # 3. python bindings
    pr['python'] = {n: _try_import(n) for n in ('mujoco', 'mujoco_py', 'dm_control', 'mujoco_mjx', 'mujoco_warp', 'gymnasium', 'gym', 'brax', 'robosuite')}
    # 4. anything else
    site = [p for p in sys.path if p.endswith('site-packages')]
    pr['site_packages_matches'] = _glob_many([os.path.join(s, '*ujoco*') for s in site] + [os.path.join(s, '*mjx*') for s in site])
    pr['wheelhouse_matches'] = _glob_many(['/opt/wheelhouse/*ujoco*', '/opt/wheelhouse/*mjx*', '/opt/wheelhouse/*dm_control*', '/opt/wheelhouse/*gym*'])
    pr['shared_objects'] = _glob_many(['/usr/lib/**/libmujoco*', '/usr/local/lib/**/libmujoco*', '/opt/**/libmujoco*', os.path.join(home, '**', 'libmujoco*')])
    pr['headers'] = _glob_many(['/usr/include/**/mujoco.h', '/usr/local/include/**/mujoco.h', '/opt/**/mujoco.h', os.path.join(home, '**', 'mujoco.h')])
    try:
        out = subprocess.run([sys.executable, '-m', 'pip', 'list', '--format=freeze'], capture_output=True, text=True, timeout=120).stdout
        pr['pip_matches'] = [ln for ln in out.splitlines() if any(k in ln.lower() for k in ('mujoco', 'mjx', 'dm-control', 'dm_control', 'gymnasium', 'brax'))]
    except Exception as ex:
        pr['pip_matches'] = 'pip unavailable: %r' % (ex,)
    try:
        out = subprocess.run(['ldconfig', '-p'], capture_output=True, text=True, timeout=30).stdout
        pr['ldconfig_matches'] = [ln.strip() for ln in out.splitlines() if 'mujoco' in ln.lower() or 'glfw' in ln.lower()]
    except Exception as ex:
        pr['ldconfig_matches'] = 'ldconfig unavailable: %r' % (ex,)
    pr['reference_checkout'] = {'path': os.environ.get('CASSIE_REFERENCE', '/root/reference'),
                                'model_xml_present': os.path.exists(os.path.join(os.environ.get('CASSIE_REFERENCE', '/root/reference'), 'model', 'cassie.xml'))}
    usable = None
    if pr['~/.mujoco']['libmujoco210']:
        usable = 'libmujoco210 at ' + pr['~/.mujoco']['libmujoco210'][0]
    elif pr['python']['mujoco']['found']:
        usable = 'python mujoco ' + str(pr['python']['mujoco']['version'])
    elif pr['python']['mujoco_py']['found']:
        usable = 'python mujoco_py ' + str(pr['python']['mujoco_py']['version'])
    elif pr['python']['dm_control']['found']:
        usable = 'dm_control ' + str(pr['python']['dm_control']['version'])
    elif pr['shared_objects']:
        usable = 'shared object ' + pr['shared_objects'][0]
    rep['usable'] = usable
    rep['conclusion'] = ('a MuJoCo is reachable: ' + usable) if usable else \
        'no MuJoCo of any version is reachable on this machine: physics parity stays unpinned (DESIGN.md section 3)'
    return rep


def main():
    rep = find_mujoco()
    txt = json.dumps(rep, indent=1, sort_keys=True)
    print(txt)
    out_dir = os.path.join(REPO, 'gpurun_out')
    if os.path.isdir(out_dir):
        tag = 'gpu_box' if not os.path.isdir('/root/reference') else 'build_container'
        with open(os.path.join(out_dir, 'probe_reference_%s.json' % tag), 'w') as f:
            f.write(txt + '\n')


if __name__ == '__main__':
    main()
