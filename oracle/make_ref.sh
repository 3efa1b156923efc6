#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Stages the UNMODIFIED reference (pure Python: it has no build step) under oracle/_ref/ so
# that it travels to the GPU box with the snapshot (oracle/_ref/ is git-ignored, not gpurun-ignored; /root/reference
# itself does not exist there).  Nothing is edited: the files are byte-for-byte copies, checked with cmp below.
# Consumers: oracle/ref_import.py (tests that run the reference's own train() / evaluate_vae_reconstruction() on top of
# the engine's overlay, also under `-m gpu`), bench.py --impl reference / cpu_baseline (kind: "reference").
# The product (simulgen_vae_b200/) never reads oracle/_ref.
set -euo pipefail
SRC="${SIMULGEN_REFERENCE_SRC:-/root/reference}"
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
DST="$HERE/_ref"
if [ ! -f "$SRC/modules/VAE_network.py" ]; then
    echo "make_ref: no reference checkout at $SRC - keeping whatever is in $DST" >&2
    exit 0
fi
rm -rf "$DST"
mkdir -p "$DST/modules" "$DST/input_data"
cp "$SRC"/modules/*.py "$DST/modules/"
cp "$SRC/SimulGen-VAE.py" "$SRC/preset.txt" "$DST/"
cp "$SRC/input_data/condition.txt" "$DST/input_data/"
for f in "$DST"/modules/*.py; do cmp -s "$f" "$SRC/modules/$(basename "$f")"; done
( cd "$SRC" && sha256sum modules/*.py SimulGen-VAE.py preset.txt input_data/condition.txt ) > "$DST/SHA256SUMS"
echo "make_ref: staged $(ls "$DST/modules" | wc -l) module files from $SRC into $DST"
