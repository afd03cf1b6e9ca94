#!/bin/bash
# Install the UNMODIFIED reference (WSIMOD, /root/reference) into baseline/_ref (git-ignored, shipped to
# the GPU box by gpurun) for bench.py's `--impl reference` arm and cpu_baseline leg.
#
# The contract's recipe is tried first:
#   python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --target baseline/_ref <src>
# In this image it stops at the build backend (pyproject.toml: hatchling + hatch-vcs, neither in
# /opt/wheelhouse, no network).  WSIMOD is pure Python, so what its wheel would contain is the `wsimod/`
# package directory; the fallback lays exactly that down (plus the _version.py hatch-vcs would write)
# without touching a single source line.  Runtime dependencies (PyYAML, tqdm, dill, pandas, pyarrow) are
# part of the image.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
SRC="${1:-/root/reference}"
DST="$HERE/_ref"
[ -d "$SRC/wsimod" ] || { echo "no reference at $SRC"; exit 1; }
rm -rf "$DST"; mkdir -p "$DST"
TMP="$(mktemp -d)"; cp -r "$SRC" "$TMP/src"   # /root/reference is read-only
if python -m pip install --quiet --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
      --target "$DST" "$TMP/src" >"$HERE/_ref_pip.log" 2>&1; then
  echo "pip install ok" > "$DST/INSTALL_METHOD"
else
  cp -r "$SRC/wsimod" "$DST/wsimod"
  find "$DST" -name "__pycache__" -type d -prune -exec rm -rf {} +
  echo '__version__ = version = "0.0.0+reference"' > "$DST/wsimod/_version.py"
  echo "package directory laid down (pip: build backend hatchling unavailable offline, see baseline/_ref_pip.log)" > "$DST/INSTALL_METHOD"
fi
rm -rf "$TMP"
PYTHONDONTWRITEBYTECODE=1 PYTHONPATH="$DST" python -c "import wsimod, wsimod.orchestration.model as m; print('reference importable from', wsimod.__file__)"
cat "$DST/INSTALL_METHOD"
