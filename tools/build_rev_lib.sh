#!/bin/bash
# Builds the library of a git revision next to the working tree's one, for same-box A/B runs (M5_LIB_PATH=<out>):
#   tools/build_rev_lib.sh <rev> [name]   ->   mars5-tts_b200/lib/variants/libmars5_b200_<name>.so
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REV=${1:-HEAD}; NAME=${2:-$REV}
TMP=$(mktemp -d)
(cd $ROOT && git archive $REV mars5-tts_b200/csrc include | tar -x -C $TMP)
make -s -j8 -C $TMP/mars5-tts_b200/csrc > /dev/null
mkdir -p $ROOT/mars5-tts_b200/lib/variants
cp $TMP/mars5-tts_b200/lib/libmars5_b200.so $ROOT/mars5-tts_b200/lib/variants/libmars5_b200_$NAME.so
rm -rf $TMP
ls -la $ROOT/mars5-tts_b200/lib/variants/libmars5_b200_$NAME.so
