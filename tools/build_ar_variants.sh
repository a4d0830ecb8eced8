#!/bin/bash
# A/B builds of the AR decode kernel for tools/ar_decode_bench.py (M5_LIB_PATH=...):
#   v0 = csrc of a git revision (default HEAD), v1 = working tree, v2 = working tree with the -D flags of $AD_V2_FLAGS
set -e
ROOT=$(cd "$(dirname "$0")/.." && pwd)
REV=${1:-HEAD}
CS=$ROOT/mars5-tts_b200/csrc; BD=$ROOT/mars5-tts_b200/build; OUT=$ROOT/mars5-tts_b200/lib/variants
NVCC=/usr/local/cuda/bin/nvcc
FLAGS="-O3 -std=c++17 -lineinfo -Xcompiler -fPIC -gencode arch=compute_100a,code=sm_100a --expt-relaxed-constexpr"
make -C $CS > /dev/null
mkdir -p $OUT $BD/v0src $BD/v0 $BD/v2
others=$(ls $BD/*.o | grep -v -e '/ar.o' -e '/ar_decode.o' -e '/sampler.o')
# v0
(cd $ROOT && git archive $REV mars5-tts_b200/csrc include | tar -x -C $BD/v0src)
for f in ar ar_decode sampler; do $NVCC $FLAGS -I$BD/v0src/mars5-tts_b200/csrc -c $BD/v0src/mars5-tts_b200/csrc/$f.cu -o $BD/v0/$f.o; done
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libmars5_b200_v0.so $others $BD/v0/ar.o $BD/v0/ar_decode.o $BD/v0/sampler.o -lcudart
# v1
cp $ROOT/mars5-tts_b200/lib/libmars5_b200.so $OUT/libmars5_b200_v1.so
# v2
$NVCC $FLAGS $AD_V2_FLAGS -c $CS/ar_decode.cu -o $BD/v2/ar_decode.o
$NVCC -shared -gencode arch=compute_100a,code=sm_100a -o $OUT/libmars5_b200_v2.so $others $BD/ar.o $BD/v2/ar_decode.o $BD/sampler.o -lcudart
ls -la $OUT
