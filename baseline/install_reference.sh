#!/bin/sh
# Installs the UNMODIFIED reference (mystic) into baseline/_ref (git-ignored, travels to the GPU box)
# so that bench.py's CPU legs can time the reference's own DifferentialEvolutionSolver2 there.
# /root/reference is read-only and the build writes into its source tree, hence the copy; klepto is
# absent from this image (oracle/_shim/klepto stands in: two argument-count helpers), hence --no-deps.
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
rm -rf /tmp/mystic_refcopy "$HERE/_ref"
cp -r /root/reference /tmp/mystic_refcopy
python -m pip install --no-index --no-build-isolation --find-links /opt/wheelhouse --no-deps \
    --target "$HERE/_ref" /tmp/mystic_refcopy
rm -rf /tmp/mystic_refcopy
