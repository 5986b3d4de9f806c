#!/usr/bin/env python
"""Snapshot the reference's Python sources into ``baseline/_ref`` (the reference arm of bench.py and the live drop-in
tests run on the GPU box, where /root/reference does not exist).

The reference is pure Python with no setup.py / pyproject, so ``pip install --target baseline/_ref /root/reference``
has nothing to build; this recipe is the equivalent: a verbatim, unmodified copy of the importable packages
(CaSE, common, GTTP, Masque, GLKS + the root Utils.py), byte for byte, with a manifest of sha256 digests so the copy can
be checked against the source tree.  ``baseline/_ref`` is listed in .gitignore (reference sources never enter this
repository's history) but not in .gpurunignore (it travels to the GPU box like the built .so).

    python baseline/make_ref.py [--src /root/reference]
"""
import argparse
import hashlib
import json
import os
import shutil
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
DST = os.path.join(HERE, '_ref')
PACKAGES = ('CaSE', 'common', 'GTTP', 'Masque', 'GLKS')
ROOT_FILES = ('Utils.py',)


def snapshot(src='/root/reference', dst=DST, quiet=False):
    """-> number of files copied (0 when ``src`` is absent: the GPU box only uses the prebuilt snapshot)."""
    if not os.path.isdir(src):
        return 0
    manifest = {}
    if os.path.isdir(dst):
        shutil.rmtree(dst)
    os.makedirs(dst)
    files = [(f, f) for f in ROOT_FILES if os.path.isfile(os.path.join(src, f))]
    for pkg in PACKAGES:
        for name in sorted(os.listdir(os.path.join(src, pkg))):
            if name.endswith('.py'):
                files.append((os.path.join(pkg, name), os.path.join(pkg, name)))
    for rel, out in files:
        os.makedirs(os.path.dirname(os.path.join(dst, out)) or dst, exist_ok=True)
        with open(os.path.join(src, rel), 'rb') as f:
            blob = f.read()
        with open(os.path.join(dst, out), 'wb') as f:
            f.write(blob)
        manifest[out] = hashlib.sha256(blob).hexdigest()
    with open(os.path.join(dst, 'MANIFEST.json'), 'w') as f:
        json.dump(dict(source=src, files=manifest), f, indent=1, sort_keys=True)
    if not quiet:
        print(f'baseline/_ref: {len(manifest)} reference files snapshotted from {src}')
    return len(manifest)


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--src', default='/root/reference')
    a = ap.parse_args()
    n = snapshot(a.src)
    sys.exit(0 if n else 1)
