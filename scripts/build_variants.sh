#!/bin/bash
# Development: build libvlsa_b200 variants (P in {4,12} only) into vlsa_b200/lib/variants/ for scripts/dev_variants.py.
# usage: build_variants.sh name1:"-DFLAG ..." name2:"..."
set -e
cd "$(dirname "$0")/.."
mkdir -p vlsa_b200/lib/variants
for spec in "$@"; do
  name="${spec%%:*}"; flags="${spec#*:}"
  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -shared -DVLSA_DEV_FEWP $flags \
       -o vlsa_b200/lib/variants/$name.so vlsa_b200/csrc/api.cu &
done
wait
ls -la vlsa_b200/lib/variants/
