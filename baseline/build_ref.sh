#!/bin/bash
# Installs the reference's deep_gemm_oss (sm_100a comparator for the block-scaled FP8 GEMM) into baseline/_ref.  See README.md.
set -e
REF=${REF:-/root/reference/3rdparty/deep_gemm_oss}
HERE=$(cd "$(dirname "$0")" && pwd)
W=$(mktemp -d)
cp -r $REF/setup.py $REF/csrc $REF/deep_gemm_oss $W/
mkdir -p $W/third-party/cutlass $W/third-party/fmt
cp -r $REF/third-party/cutlass/include $W/third-party/cutlass/
cp -r $REF/third-party/fmt/include $W/third-party/fmt/
python -m pip install --no-index --no-build-isolation --no-deps --find-links /opt/wheelhouse --target $HERE/_ref $W
TORCH=$(python -c "import torch,os;print(os.path.dirname(torch.__file__))")
PYINC=$(python -c "import sysconfig;print(sysconfig.get_paths()['include'])")
SO=$(ls $HERE/_ref/deep_gemm_cpp_oss*.so)
cd $W && g++ -O2 -std=c++17 -fPIC -shared -Wno-psabi -Wno-deprecated-declarations -D_GLIBCXX_USE_CXX11_ABI=1 \
  -DTORCH_EXTENSION_NAME=deep_gemm_cpp_oss -DTORCH_API_INCLUDE_EXTENSION_H csrc/python_api.cpp \
  -I/usr/local/cuda/include -I/usr/local/cuda/include/cccl -Ideep_gemm_oss/include -Ithird-party/cutlass/include \
  -Ithird-party/fmt/include -I$TORCH/include -I$TORCH/include/torch/csrc/api/include -I$PYINC \
  -L$TORCH/lib -lc10 -ltorch -ltorch_cpu -ltorch_python -lc10_cuda -L/usr/local/cuda/lib64 -L/usr/local/cuda/lib64/stubs \
  -lcuda -lcudart -lnvrtc -o $SO
echo "installed: $SO"
