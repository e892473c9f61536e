#!/bin/bash
# TEST INFRASTRUCTURE ONLY.  Compiles the reference's own GPU implementation of the AEV stage (cuAEV:
# /root/reference/torchani/csrc/{aev.cu,cuaev.cpp}) for sm_100 FROM THE SOURCES WHERE THEY LIE into
# oracle/_ref/cuaev.so (git-ignored, travels to the GPU box).  Only tests/ load it
# (tests/test_gpu_reference_cuaev.py: parity of our AEV kernels with the reference's CUDA kernels and their
# timing on the same box).  No reference source is copied.  Recipe: SURVEY.md 2.2 (setup.py:43-80).
set -e
REF=${REF:-/root/reference}
SRC=$REF/torchani/csrc
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
[ -f "$SRC/aev.cu" ] || { echo "no reference sources under $REF: nothing to build"; exit 0; }
mkdir -p "$OUT"
STAMP="$OUT/cuaev.stamp"
FP="$(cat "$SRC/aev.cu" "$SRC/cuaev.cpp" "$SRC/aev.h" "$SRC/cuaev_cub.cuh" "$SRC/cell_list.cpp" "$0" | sha256sum | cut -d' ' -f1)"
if [ -f "$OUT/cuaev.so" ] && [ -f "$OUT/cell_list.so" ] && [ -f "$STAMP" ] && [ "$(cat "$STAMP")" = "$FP" ]; then echo "$OUT/cuaev.so up to date"; exit 0; fi
PY=${PYTHON:-python}
TORCH="$($PY -c 'import torch, os; print(os.path.dirname(torch.__file__))')"
PYINC="$($PY -c 'import sysconfig; print(sysconfig.get_paths()["include"])')"
INC="-I$TORCH/include -I$TORCH/include/torch/csrc/api/include -I/usr/local/cuda/include -I$PYINC -I$SRC"
TMP="$(mktemp -d)"
nvcc -c "$SRC/aev.cu" -o "$TMP/aev.o" $INC -std=c++17 --expt-extended-lambda --expt-relaxed-constexpr \
  -DCUB_NS_QUALIFIER=::cuaev::cub "-DCUB_NS_PREFIX=namespace cuaev {" "-DCUB_NS_POSTFIX=}" \
  -DTORCHANI_OPT -use_fast_math -gencode=arch=compute_100,code=sm_100 -Xcompiler -fPIC -D_GLIBCXX_USE_CXX11_ABI=1
g++ -c "$SRC/cuaev.cpp" -o "$TMP/cuaev.o" $INC -std=c++17 -fPIC -O2 -D_GLIBCXX_USE_CXX11_ABI=1
g++ -shared "$TMP/aev.o" "$TMP/cuaev.o" -o "$OUT/cuaev.so" -L"$TORCH/lib" -L/usr/local/cuda/lib64 \
  -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -lcudart -Wl,-rpath,"$TORCH/lib"
# the reference's C++/ATen cell list (csrc/cell_list.cpp:342-363, torch.ops.cell_list.cell_list): the
# neighbour stage of its GPU path
g++ -shared "$SRC/cell_list.cpp" -o "$OUT/cell_list.so" $INC -std=c++17 -fPIC -O2 -fopenmp -D_GLIBCXX_USE_CXX11_ABI=1 \
  -L"$TORCH/lib" -lc10 -ltorch_cpu -ltorch -Wl,-rpath,"$TORCH/lib"
rm -rf "$TMP"
echo "$FP" > "$STAMP"
echo "built $OUT/cuaev.so"
