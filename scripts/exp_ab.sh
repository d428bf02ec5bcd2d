#!/bin/bash
# A/B build variants of libwam.so with -D toggles and time the bench kernel for each.
set -e
cd "$(dirname "$0")/.."
mkdir -p build/ab
for v in "$@"; do
  name=$(echo "$v" | tr ' =' '__' | tr -d '-')
  [ -z "$name" ] && name=base
  nvcc $v -gencode arch=compute_100a,code=sm_100a -lineinfo -O3 -std=c++17 -Xcompiler -fPIC -shared -o build/ab/libwam_$name.so webaudio-modem_b200/csrc/wam_api.cu
done
