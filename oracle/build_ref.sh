#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds the UNMODIFIED reference extensions
# (pointnet2_cuda, iou3d_cuda, roipool3d_cuda) from the sources where they lie
# under /root/reference into oracle/_ref/ (git-ignored, but it travels to the
# GPU box with gpurun).  No reference source is copied into this repository:
# nvcc/g++ are pointed at the read-only tree, only objects and .so files are
# written here.  Recipe = SURVEY.md Appendix A (THC shim + AT_CHECK alias).
#
# Usage: oracle/build_ref.sh [reference_root]      (default /root/reference)
set -euo pipefail
REF=${1:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
OBJ=$OUT/obj
if [ ! -d "$REF/pointnet2_lib/pointnet2/src" ]; then
  echo "build_ref: $REF not present; keeping prebuilt oracle/_ref" >&2
  exit 0
fi
mkdir -p "$OBJ" "$OUT/shim/THC"

# --- shims (ours, not reference code) -------------------------------------
cat > "$OUT/shim/THC/THC.h" <<'EOF'
#pragma once
#include <ATen/cuda/CUDAContext.h>
#include <c10/cuda/CUDAStream.h>
struct THCState;
static inline cudaStream_t THCState_getCurrentStream(THCState *) {
  return at::cuda::getCurrentCUDAStream().stream();
}
EOF
echo 'struct THCState; THCState *state = nullptr;' > "$OUT/state.cpp"

PY=${PYTHON:-python}
TORCH_INC=$($PY - <<'EOF'
import torch.utils.cpp_extension as c, sysconfig
print(" ".join("-I" + p for p in c.include_paths() + [sysconfig.get_paths()["include"]]))
EOF
)
TORCH_LIB=$($PY -c 'import torch, os; print(os.path.join(os.path.dirname(torch.__file__), "lib"))')
CUDA=${CUDA_HOME:-/usr/local/cuda}
INC="$TORCH_INC -I$CUDA/include -I$OUT/shim"
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVCC="$CUDA/bin/nvcc -O2 -std=c++17 -Xcompiler -fPIC $ARCH"
CXX="g++ -std=c++17 -O2 -fPIC -DAT_CHECK=TORCH_CHECK -Wno-deprecated-declarations"

P2=$REF/pointnet2_lib/pointnet2/src
IOU=$REF/lib/utils/iou3d/src
ROI=$REF/lib/utils/roipool3d/src

pids=()
for f in sampling_gpu ball_query_gpu group_points_gpu interpolate_gpu; do
  [ -f "$OBJ/$f.o" ] || { $NVCC $INC -I$P2 -c "$P2/$f.cu" -o "$OBJ/$f.o" & pids+=($!); }
done
[ -f "$OBJ/iou3d_kernel.o" ]     || { $NVCC -c "$IOU/iou3d_kernel.cu" -o "$OBJ/iou3d_kernel.o" & pids+=($!); }
[ -f "$OBJ/roipool3d_kernel.o" ] || { $NVCC -c "$ROI/roipool3d_kernel.cu" -o "$OBJ/roipool3d_kernel.o" & pids+=($!); }
for f in sampling ball_query group_points interpolate; do
  [ -f "$OBJ/$f.o" ] || { $CXX $INC -I$P2 -c "$P2/$f.cpp" -o "$OBJ/$f.o" & pids+=($!); }
done
[ -f "$OBJ/pointnet2_api.o" ] || { $CXX $INC -I$P2 -DTORCH_EXTENSION_NAME=pointnet2_cuda -c "$P2/pointnet2_api.cpp" -o "$OBJ/pointnet2_api.o" & pids+=($!); }
[ -f "$OBJ/iou3d.o" ]         || { $CXX $INC -DTORCH_EXTENSION_NAME=iou3d_cuda -c "$IOU/iou3d.cpp" -o "$OBJ/iou3d.o" & pids+=($!); }
[ -f "$OBJ/roipool3d.o" ]     || { $CXX $INC -DTORCH_EXTENSION_NAME=roipool3d_cuda -c "$ROI/roipool3d.cpp" -o "$OBJ/roipool3d.o" & pids+=($!); }
[ -f "$OBJ/state.o" ]         || { $CXX -c "$OUT/state.cpp" -o "$OBJ/state.o" & pids+=($!); }
for p in "${pids[@]}"; do wait "$p"; done

LIBS="-L$TORCH_LIB -L$CUDA/lib64 -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -ltorch_python -lcudart -Wl,-rpath,$TORCH_LIB"
g++ -shared "$OBJ"/{pointnet2_api,sampling,ball_query,group_points,interpolate,sampling_gpu,ball_query_gpu,group_points_gpu,interpolate_gpu,state}.o $LIBS -o "$OUT/pointnet2_cuda.so"
g++ -shared "$OBJ"/{iou3d,iou3d_kernel}.o $LIBS -o "$OUT/iou3d_cuda.so"
g++ -shared "$OBJ"/{roipool3d,roipool3d_kernel}.o $LIBS -o "$OUT/roipool3d_cuda.so"
echo "build_ref: wrote $OUT/{pointnet2_cuda,iou3d_cuda,roipool3d_cuda}.so"

# --- the reference's Python op wrappers as one archive next to the extensions they bind (see oracle/stage_refpy.py)
$PY "$HERE/stage_refpy.py" "$REF" "$OUT/refpy.zip"
