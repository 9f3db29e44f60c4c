#!/bin/bash
# TEST INFRASTRUCTURE — stages the files of the UNMODIFIED reference that its step()/reset() path imports into
# oracle/_ref/ (git-ignored, never part of the history; it travels to the GPU box with the gpurun snapshot like a
# built .so), so that `bench.py --impl reference` and oracle/ref_subproc_bench.py can time the reference itself on
# the box's host cores.  Nothing is modified; the package never reads this directory.
#   oracle/make_ref.sh [/root/reference]
set -e
SRC=${1:-/root/reference}
DST="$(cd "$(dirname "$0")" && pwd)/_ref"
[ -f "$SRC/src/envs_train/env_super.py" ] || { echo "no reference checkout at $SRC"; exit 1; }
rm -rf "$DST"
mkdir -p "$DST/src/envs_train" "$DST/utils"
cp "$SRC"/src/envs_train/env_super.py "$SRC"/src/envs_train/env_1_sort.py "$SRC"/src/envs_train/env_2_press.py \
   "$SRC"/src/envs_train/env_monolith.py "$DST/src/envs_train/"
cp "$SRC"/utils/input_generator.py "$SRC"/utils/plotting.py "$DST/utils/"
cp "$SRC"/config.yml "$DST/"
( cd "$SRC" && sha256sum src/envs_train/*.py utils/input_generator.py utils/plotting.py config.yml ) > "$DST/SHA256SUMS"
echo "staged $(wc -l < "$DST/SHA256SUMS") reference files into $DST"
