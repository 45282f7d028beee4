#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (see oracle/README.md).
# Builds the UNMODIFIED reference raymarching extension (CUDA, pybind11/ATen) for sm_100a from the
# sources where they lie under /root/reference (never copied into this repo). Output: oracle/_ref/_raymarching.so
# It is the bit-exact GPU oracle ("kernel to beat") for rows a1-a10 of SURVEY.md section 8; it only RUNS on a GPU box.
# The recipe follows SURVEY.md Appendix B (--expt-relaxed-constexpr is what torch's CUDAExtension adds by default).
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${REFERENCE_ROOT:-/root/reference}/submodules/raymarching/src"
OUT="$HERE/_ref"
if [ ! -f "$REF/raymarching.cu" ]; then
  echo "[build_ref] reference sources not present ($REF); keeping prebuilt $OUT/_raymarching.so if any"; exit 0
fi
mkdir -p "$OUT"
if [ -f "$OUT/_raymarching.so" ] && [ "$OUT/_raymarching.so" -nt "$REF/raymarching.cu" ]; then
  echo "[build_ref] up to date"; exit 0
fi
PY="${PYTHON:-python}"
T="$($PY -c 'import torch,os;print(os.path.dirname(torch.__file__))')"
PYINC="$($PY -c 'import sysconfig;print(sysconfig.get_paths()["include"])')"
INC="-I$T/include -I$T/include/torch/csrc/api/include -I$PYINC -I/usr/local/cuda/include"
DEFS="-DTORCH_EXTENSION_NAME=_raymarching -DTORCH_API_INCLUDE_EXTENSION_H"
nvcc -O3 -std=c++17 --expt-relaxed-constexpr -U__CUDA_NO_HALF_OPERATORS__ -U__CUDA_NO_HALF_CONVERSIONS__ -U__CUDA_NO_HALF2_OPERATORS__ \
  -gencode arch=compute_100a,code=sm_100a $INC --compiler-options -fPIC $DEFS -c "$REF/raymarching.cu" -o "$OUT/raymarching.o" &
g++ -O3 -std=c++17 -fPIC $INC $DEFS -c "$REF/bindings.cpp" -o "$OUT/bindings.o" &
wait
g++ -shared "$OUT/raymarching.o" "$OUT/bindings.o" -L"$T/lib" -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -ltorch_cuda \
  -L/usr/local/cuda/lib64 -lcudart -Wl,-rpath,"$T/lib" -o "$OUT/_raymarching.so"
rm -f "$OUT/raymarching.o" "$OUT/bindings.o"
echo "[build_ref] built $OUT/_raymarching.so"
