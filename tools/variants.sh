#!/bin/bash
# Development aid: build the library once per set of -D flags, side by side, for A/B runs on the GPU box.
#   tools/variants.sh name1 "-DFOO" name2 "-DFOO -DBAR" ...   ->  gmerlin-avdecoder_b200/lib/variant_<name>.so
# Use with RTJPEG_B200_LIBFILE=$PWD/gmerlin-avdecoder_b200/lib/variant_<name>.so
set -e
cd "$(dirname "$0")/.."
while [ $# -ge 2 ]; do
    name=$1; flags=$2; shift 2
    rm -rf gmerlin-avdecoder_b200/build
    make -s -C gmerlin-avdecoder_b200 EXTRA_NVFLAGS="$flags" > /dev/null
    cp gmerlin-avdecoder_b200/lib/librtjpeg_b200.so gmerlin-avdecoder_b200/lib/variant_$name.so
    echo "built variant_$name ($flags)"
done
rm -rf gmerlin-avdecoder_b200/build
make -s -C gmerlin-avdecoder_b200 > /dev/null
