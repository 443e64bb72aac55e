#!/usr/bin/env bash
# In-tree build of the C-ABI library for sm_100a (cross-compiles without a GPU).
# Output: ../lib/librawboost_b200.so  (git-ignored, travels to the GPU box with the snapshot)
set -euo pipefail
here="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
root="$(cd "$here/../.." && pwd)"
out="$here/../lib"
name="${RB_LIB_NAME:-librawboost_b200}"   # RB_LIB_NAME / RB_EXTRA_FLAGS: kernel-variant experiments
mkdir -p "$out"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
FLAGS=(-gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -std=c++17 -Xcompiler -fPIC -Xcompiler -fvisibility=hidden
       -Xptxas -v -I"$root/include" -I"$here" ${RB_EXTRA_FLAGS:-})
objs=()
for src in rb_fir_bank rb_dense rb_api rb_probe rb_devplan rb_multiview; do
  extra=()
  # the FP32 peak probe wants its multiplier in a uniform register (FFMA2 R, R, UR, R: two register-file operands instead of
  # three); ptxas does that at the lower register-usage levels only (rb_probe.cu)
  [ "$src" = rb_probe ] && extra=(-Xptxas --register-usage-level=3)
  "$NVCC" "${FLAGS[@]}" "${extra[@]}" -c "$here/$src.cu" -o "$out/$name.$src.o" 2> "$out/$name.$src.ptxas.log" || { cat "$out/$name.$src.ptxas.log" >&2; exit 1; }
  objs+=("$out/$name.$src.o")
done
# host-only planner: plain g++ (function multiversioning for the MT19937 block)
CUDA_INC="$(dirname "$(dirname "$NVCC")")/include"
"${CXX:-g++}" -O3 -std=c++17 -fPIC -fvisibility=hidden -I"$root/include" -I"$CUDA_INC" -c "$here/rb_planner.cpp" -o "$out/$name.rb_planner.o"
objs+=("$out/$name.rb_planner.o")
"$NVCC" -gencode arch=compute_100a,code=sm_100a -shared -o "$out/$name.so" "${objs[@]}" -Xcompiler -fPIC
# Post-link: operand-reuse flags / yield hints of the FFMA2 streams of fir_bank_kernel (control bits only, see sass_reuse_patch.py).
# A failure leaves the unpatched library in place (same results, ~3-5 % slower FIR kernels) and says so.
if [ -z "${RB_NO_SASS_PATCH:-}" ]; then
  if CUOBJDUMP="$(dirname "$NVCC")/cuobjdump" python3 "$here/sass_reuse_patch.py" "$out/$name.so" "$out/$name.so.patched"; then
    mv "$out/$name.so.patched" "$out/$name.so"
  else
    rm -f "$out/$name.so.patched"; echo "WARNING: sass_reuse_patch.py failed; $name.so is left as ptxas scheduled it" >&2
  fi
fi
grep -h -E "Used [0-9]+ registers|spill" "$out"/$name.*.ptxas.log | sort | uniq -c | sort -rn | head -20
echo "built $out/$name.so"
