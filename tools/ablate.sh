#!/bin/bash
# Side builds of the library with one stage of classify_bulk_kernel compiled out (-DIRP_ABLATE=mask), for TIMING ONLY:
# what each stage costs inside the shipped kernel is the difference to the full build.  Results of these libraries
# are wrong by construction; they live under build/ (git-ignored) and are never loaded by the package (IRP_LIB_PATH points the
# bench at them).  Usage: tools/ablate.sh build | tools/ablate.sh run
cd "$(dirname "$0")/.."
CS=image-restoration-platform_b200/csrc
if [ "$1" = build ]; then
  mkdir -p build/ablate
  for m in 1 2 4 8 16 6; do
    /usr/local/cuda/bin/nvcc -O3 -std=c++17 -lineinfo -gencode arch=compute_100a,code=sm_100a -Xcompiler -fPIC -DIRP_ABLATE=$m -shared \
      -o build/ablate/libirp_ablate_$m.so $CS/irp_lib.cu -lcudart &
  done
  wait
else
  for m in 0 1 2 4 8 16 6; do
    lib=build/ablate/libirp_ablate_$m.so
    [ $m = 0 ] && lib=image-restoration-platform_b200/libirp_b200.so
    IRP_LIB_PATH=$PWD/$lib python bench.py --steps 30 --warmup 3 --no-e2e --no-cpu --no-jpeg 2>/dev/null | python -c "
import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('ablate $m', d['roofline']['other_kernel'])"
  done
fi
