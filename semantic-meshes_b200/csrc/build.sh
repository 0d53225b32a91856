#!/bin/bash
# Builds semantic-meshes_b200/semantic_meshes/libsmesh_b200.so (sm_100a only; nvcc cross-compiles without a GPU).
set -e
HERE="$(cd "$(dirname "$0")" && pwd)"
OUT="$HERE/../semantic_meshes/libsmesh_b200.so"
NVCC="${NVCC:-nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
COMMON="-std=c++17 -O3 -lineinfo $ARCH -Xcompiler -fPIC ${SMESH_NVCC_EXTRA:-}"
mkdir -p "$HERE/build"
# stale objects must never be linked if a compile fails
rm -f "$HERE/build/smesh_api.o" "$HERE/build/smesh_raster.o" "$HERE/build/smesh_fuse.o" "$HERE/build/smesh_pipeline.o"
pids=()
$NVCC $COMMON -c "$HERE/smesh_api.cu" -o "$HERE/build/smesh_api.o" &
pids+=($!)
# the rasterizer's arithmetic contract is explicit intrinsics; -fmad=false guards anything written as plain operators
$NVCC $COMMON -fmad=false -Xcompiler -fopenmp -c "$HERE/smesh_raster.cu" -o "$HERE/build/smesh_raster.o" &
pids+=($!)
$NVCC $COMMON -c "$HERE/smesh_fuse.cu" -o "$HERE/build/smesh_fuse.o" &
pids+=($!)
$NVCC $COMMON -c "$HERE/smesh_pipeline.cu" -o "$HERE/build/smesh_pipeline.o" &
pids+=($!)
for pid in "${pids[@]}"; do
  wait "$pid" || { echo "build.sh: a compile job failed" >&2; exit 1; }
done
$NVCC $ARCH -shared -o "$OUT" "$HERE/build/smesh_api.o" "$HERE/build/smesh_raster.o" "$HERE/build/smesh_fuse.o" "$HERE/build/smesh_pipeline.o" -lcudart -lgomp
echo "built $OUT"
