#!/bin/bash
# builds experimental variants of libpyflyt_b200.so into pyflyt_b200/lib/variants/<tag>/ (timed by tools/time_variants.py)
# usage: tools/build_variants.sh tag1="-DFOO=1" tag2="-DBAR=0 -DBAZ"
set -e
cd "$(dirname "$0")/../pyflyt_b200/csrc"
build() { tag=$1; shift; mkdir -p ../lib/variants/$tag; make -s -j4 OUT=../lib/variants/$tag VARIANT="$*" >/dev/null; echo "built $tag: $*"; }
for spec in "$@"; do tag=${spec%%=*}; flags=${spec#*=}; build $tag "$flags" & done
wait
