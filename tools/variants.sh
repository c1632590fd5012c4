#!/bin/bash
# builds variant copies of the library for A/B timing on the GPU: tools/variants.sh name "EXTRA flags" ...
# e.g. tools/variants.sh lb256x4 "-DOFFTB_LB_512=256,4"
set -e
cd "$(dirname "$0")/../offt_b200/csrc"
mkdir -p ../lib/variants
while [ $# -ge 2 ]; do
  name=$1; extra=$2; shift 2
  make -j"$(nproc)" BUILD=build_$name LIBNAME=variants/lib_$name.so EXTRA="$extra" > /tmp/variant_$name.log 2>&1 || { tail -20 /tmp/variant_$name.log; exit 1; }
  echo "built variants/lib_$name.so ($extra)"
done
