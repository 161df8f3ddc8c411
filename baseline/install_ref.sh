#!/bin/bash
# Installs the UNMODIFIED reference (NikNazarov/TorchPIV) into baseline/_ref/ with pip, as the bench contract
# prescribes.  baseline/_ref/ is git-ignored (no reference source enters the history) but not gpurun-ignored,
# so the installed package travels to the GPU box, where `bench.py --impl reference` and the
# `torch_cuda_reference` leg of `bench.py` import it through oracle/ref_loader.py.
#
#   bash baseline/install_ref.sh [/root/reference]
#
# --no-deps: the reference pins numpy==1.26.3 / imageio / PyQt5 / matplotlib / watchdog, none of which its
# PIV path needs; the backend runs unmodified on the image's numpy / torch / scipy / cv2.
set -euo pipefail
SRC="${1:-/root/reference}"
HERE="$(cd "$(dirname "$0")" && pwd)"
TMP="$(mktemp -d)"
cp -r "$SRC" "$TMP/ref"            # the source tree is read-only; setuptools writes build/ and *.egg-info
rm -rf "$HERE/_ref"
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse \
    --target "$HERE/_ref" "$TMP/ref"
rm -rf "$TMP"
ls "$HERE/_ref/torchPIV/PIVbackend.py"
