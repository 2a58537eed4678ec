#!/usr/bin/env python
"""Install the UNMODIFIED reference (a pure-Python repo with no setup.py / pyproject, so `pip install --target` has
nothing to build) into git-ignored baseline/_ref/ so that it travels to the GPU box with the gpurun snapshot:

    baseline/_ref/src/*.py     byte-for-byte copies of /root/reference/src/*.py
    baseline/_ref/MANIFEST     source path, commit and sha256 of every file

baseline/_ref/ is listed in .gitignore (never committed) and NOT in .gpurunignore.  Nothing in the product imports it;
only tests/, bench.py (--impl reference / --impl torch-gpu / gpu_incumbent) and __graft_entry__ use it, through
oracle/refload.py.  Run automatically by __graft_entry__.build() when /root/reference exists.
"""
import hashlib
import json
import os
import shutil
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.environ.get('KBNET_REF', '/root/reference')
DST = os.path.join(ROOT, 'baseline', '_ref')


def install(verbose=True):
    src_dir = os.path.join(SRC, 'src')
    if not os.path.isdir(src_dir):
        if verbose:
            print('install_reference: %s not present (GPU box?) -- keeping whatever is in %s' % (src_dir, DST))
        return os.path.isdir(os.path.join(DST, 'src'))
    out = os.path.join(DST, 'src')
    os.makedirs(out, exist_ok=True)
    manifest = {'source': SRC, 'files': {}}
    try:
        manifest['commit'] = json.load(open(os.path.join(SRC, '.SUBMODULES.json'))).get('commit')
    except Exception:
        manifest['commit'] = None
    for f in sorted(os.listdir(src_dir)):
        if f.endswith('.py'):
            shutil.copyfile(os.path.join(src_dir, f), os.path.join(out, f))
            manifest['files'][f] = hashlib.sha256(open(os.path.join(out, f), 'rb').read()).hexdigest()
    json.dump(manifest, open(os.path.join(DST, 'MANIFEST'), 'w'), indent=1, sort_keys=True)
    if verbose:
        print('install_reference: %d files -> %s' % (len(manifest['files']), out))
    return True


if __name__ == '__main__':
    sys.exit(0 if install() else 1)
