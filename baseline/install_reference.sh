#!/bin/sh
# Install the unmodified reference (pure Python) into the git-ignored baseline/_ref/ for bench.py's CPU arm.
# /root/reference is read-only and setuptools writes build/ + egg-info into the source tree: install from a copy.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/ref"
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/ref"
# setuptools' find_packages() skips the reference's namespace packages (directories without __init__.py:
# promptttspp/models, promptttspp/modules/nnsvs); copy them verbatim so the installed tree equals the source tree.
for d in models modules/nnsvs; do
  [ -d "$TMP/ref/promptttspp/$d" ] && cp -r "$TMP/ref/promptttspp/$d" "$HERE/_ref/promptttspp/$d"
done
rm -rf "$TMP"
