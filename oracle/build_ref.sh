#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY (oracle).  Builds oracle/_ref/unet_ref: the reference's own
# unet.cpp/unet.hpp compiled UNCHANGED, in place, from /root/reference against the image's libtorch,
# plus oracle/ref_build/driver.cpp.  calc_losses (train.cpp:501-552) and default_feature
# (train.cpp:1054-1069) are cut out of /root/reference/train.cpp by pattern at build time into the
# git-ignored oracle/_ref/train_extract.inc (train.cpp itself needs Qt + TIPL and cannot be compiled).
# Outputs go ONLY to oracle/_ref/ (git-ignored, not gpurun-ignored, so the binary travels to the GPU box).
# No reference source is copied into tracked files.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REFERENCE_DIR:-/root/reference}"
OUT="$HERE/_ref"
PY="${PYTHON:-python}"
if [ ! -f "$REF/unet.cpp" ]; then
    echo "build_ref: $REF/unet.cpp not found (GPU box?) - keeping prebuilt $OUT/unet_ref" >&2
    exit 0
fi
mkdir -p "$OUT"
T="$($PY -c 'import torch, os; print(os.path.dirname(torch.__file__))')"
ABI="$($PY -c 'import torch; print(int(torch._C._GLIBCXX_USE_CXX11_ABI))')"

awk '
/^inline std::tuple<torch::Tensor,torch::Tensor,torch::Tensor> calc_losses\(/ {on=1}
/^std::string default_feature\(int out_count\)/ {on=1}
on {print}
on && /^}/ {on=0; print ""}
' "$REF/train.cpp" > "$OUT/train_extract.inc"
grep -q "calc_losses" "$OUT/train_extract.inc"
grep -q "default_feature" "$OUT/train_extract.inc"

STAMP="$OUT/.stamp"
NEW="$(cat "$REF/unet.cpp" "$REF/unet.hpp" "$OUT/train_extract.inc" "$HERE/ref_build/driver.cpp" "$HERE/ref_build/shim/TIPL/tipl.hpp" | sha1sum | cut -d' ' -f1)-$($PY -c 'import torch; print(torch.__version__)')"
if [ -x "$OUT/unet_ref" ] && [ -f "$STAMP" ] && [ "$(cat "$STAMP")" = "$NEW" ]; then
    echo "build_ref: up to date"
    exit 0
fi
CXXFLAGS=(-std=c++17 -O2 -D_GLIBCXX_USE_CXX11_ABI="$ABI" -I"$HERE/ref_build/shim" -I"$REF" -I"$OUT"
          -I"$T/include" -I"$T/include/torch/csrc/api/include")
g++ "${CXXFLAGS[@]}" -c "$REF/unet.cpp" -o "$OUT/unet_reference.o" &
g++ "${CXXFLAGS[@]}" -c "$HERE/ref_build/driver.cpp" -o "$OUT/driver.o" &
wait
LIBS=(-ltorch -ltorch_cpu -lc10)
if [ -f "$T/lib/libtorch_cuda.so" ]; then LIBS+=(-Wl,--no-as-needed -ltorch_cuda -lc10_cuda -Wl,--as-needed); fi
g++ "$OUT/unet_reference.o" "$OUT/driver.o" -o "$OUT/unet_ref" -L"$T/lib" -Wl,-rpath,"$T/lib" "${LIBS[@]}" -lz -lpthread
echo "$NEW" > "$STAMP"
echo "build_ref: built $OUT/unet_ref"
