#!/bin/bash
# build named variants of libfrp_b200.so into tools/variants/<name>.so (git-ignored; shipped to the
# GPU box by gpurun).  usage: tools/build_variants.sh name1="-DFLAG=1 -DX=2" name2="..."
set -e
cd "$(dirname "$0")/.."
mkdir -p tools/variants
for spec in "$@"; do
  name="${spec%%=*}"; flags="${spec#*=}"
  FRP_NVCC_FLAGS="$flags" python face-recognition-platform_b200/build.py >/dev/null
  cp face-recognition-platform_b200/frp_b200/_lib/libfrp_b200.so tools/variants/$name.so
  echo "built $name ($flags)"
done
FRP_NVCC_FLAGS="" python face-recognition-platform_b200/build.py >/dev/null
echo "restored default build"
