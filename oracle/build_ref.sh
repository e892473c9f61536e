#!/bin/bash
# TEST / BASELINE INFRASTRUCTURE ONLY.  Makes the UNMODIFIED reference runnable beside the product:
#   1. compiles the reference's own native extensions FROM THE SOURCES WHERE THEY LIE
#      (/root/reference/torchani/csrc/{aev.cu,cuaev.cpp,cell_list.cpp,mnp.cpp}, recipe: setup.py:43-130)
#      for sm_100 into oracle/_ref/{cuaev,cell_list,mnp}.so, and
#   2. stages the reference's Python package (pure .py files, no csrc sources) next to them as
#      oracle/_ref/torchani/ with the three extensions where torchani/csrc/__init__.py:8-10 looks for
#      them, so that `import torchani` on the GPU box gives the real reference with strategy="cuaev",
#      neighborlist="cell_list" and .to_infer_model() available.
# oracle/_ref/ is git-ignored (never committed, no reference source enters the history) but NOT
# gpurun-ignored: it travels to the GPU box, where /root/reference does not exist.  Only tests/,
# bench.py's reference / cpu_baseline legs and tools/reference_gpu_path.py load it (oracle/ref_torchani.py).
set -e
REF=${REF:-/root/reference}
SRC=$REF/torchani/csrc
OUT="$(cd "$(dirname "$0")" && pwd)/_ref"
[ -f "$SRC/aev.cu" ] || { echo "no reference sources under $REF: nothing to build"; exit 0; }
mkdir -p "$OUT"
STAMP="$OUT/cuaev.stamp"
FP="$(cat "$SRC/aev.cu" "$SRC/cuaev.cpp" "$SRC/aev.h" "$SRC/cuaev_cub.cuh" "$SRC/cell_list.cpp" "$SRC/mnp.cpp" "$0" | sha256sum | cut -d' ' -f1)"
PY=${PYTHON:-python}
if [ -f "$OUT/cuaev.so" ] && [ -f "$OUT/cell_list.so" ] && [ -f "$OUT/mnp.so" ] && [ -f "$STAMP" ] && [ "$(cat "$STAMP")" = "$FP" ]; then
  echo "$OUT/cuaev.so up to date"
else
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
  # MNP (csrc/mnp.cpp, torch.ops.mnp.run): the multi-stream network path behind to_infer_model(use_mnp=True)
  g++ -shared "$SRC/mnp.cpp" -o "$OUT/mnp.so" $INC -std=c++17 -fPIC -O2 -fopenmp -D_GLIBCXX_USE_CXX11_ABI=1 \
    -L"$TORCH/lib" -L/usr/local/cuda/lib64 -lc10 -lc10_cuda -ltorch_cpu -ltorch_cuda -ltorch -lcudart \
    -Wl,-rpath,"$TORCH/lib" || echo "mnp.so not built (optional)"
  rm -rf "$TMP"
  echo "$FP" > "$STAMP"
  echo "built $OUT/cuaev.so"
fi
# ---- stage the Python package (every run: cheap, and keeps it in step with /root/reference)
PKG="$OUT/torchani"
rm -rf "$PKG"
cp -r "$REF/torchani" "$PKG"
chmod -R u+w "$PKG"
find "$PKG" -name '__pycache__' -type d -prune -exec rm -rf {} +
# the native sources stay where they lie: only csrc/__init__.py (the extension loader) is staged
find "$PKG/csrc" -type f ! -name '__init__.py' -delete
for so in cuaev cell_list mnp; do [ -f "$OUT/$so.so" ] && cp "$OUT/$so.so" "$PKG/$so.so"; done
echo "staged $PKG"
