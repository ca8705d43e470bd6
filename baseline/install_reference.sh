#!/bin/bash
# Offline install of the UNMODIFIED reference package (pure Python + Numba)
# into baseline/_ref (git-ignored; travels to the GPU box with the snapshot).
#
# The reference's build backend is poetry (pyproject.toml: poetry.masonry.api),
# which this image does not have, so `pip install /root/reference` fails with
# "No module named 'poetry'".  The package itself is a plain source tree
# (packages = [{include = "phd_qmclib", from = "src"}]): this script installs
# it from a scratch copy under /tmp whose ONLY change is a setuptools
# [build-system]/[project] header in place of the poetry one -- no file under
# src/ is touched.  Dependencies are not resolved (--no-deps): numba, numpy,
# scipy, attrs, mpmath come from the image; the missing ones (h5py, dask,
# colorlog, ...) are stubbed at import time by oracle/refshim.py.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
TMP="$(mktemp -d /tmp/refinstall.XXXXXX)"
cp -r "$SRC"/. "$TMP"/
cat > "$TMP/pyproject.toml" <<'TOML'
[build-system]
requires = ["setuptools"]
build-backend = "setuptools.build_meta"

[project]
name = "phd-qmclib"
version = "0.17.0"

[tool.setuptools.packages.find]
where = ["src"]
TOML
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps \
    --find-links /opt/wheelhouse --target "$HERE/_ref" "$TMP"
rm -rf "$TMP"
python - <<PY
import os, hashlib
ref = os.path.join("$SRC", "src", "phd_qmclib")
dst = os.path.join("$HERE", "_ref", "phd_qmclib")
bad = 0
for d, _, fs in os.walk(ref):
    for f in fs:
        if not f.endswith('.py'):
            continue
        a = os.path.join(d, f)
        b = os.path.join(dst, os.path.relpath(a, ref))
        if not os.path.exists(b) or open(a, 'rb').read() != open(b, 'rb').read():
            bad += 1
            print('DIFFERS', b)
print('installed; files differing from the reference tree:', bad)
PY
