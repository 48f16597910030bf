#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY -- builds the UNMODIFIED CompV reference (base+gpu+core, AVX2/SSE intrinsics,
# asm disabled because yasm is absent) from the sources where they lie under /root/reference into
# oracle/_ref/libcompv_ref.so, plus our extern "C" shim (oracle/ref_shim.cxx) as oracle/_ref/libcompv_refshim.so.
# Recipe follows SURVEY.md Appendix A (mirrors common.cmake:192-225 per-file ISA flags). No reference source is copied.
# Nothing in the product path (compv_b200/) may load these libraries; only tests/, smoke() and bench.py's CPU legs do.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
R="${COMPV_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
OBJ="$OUT/obj"
if [ ! -d "$R/base" ]; then
  echo "[build_ref] $R not present: keeping prebuilt $OUT (GPU box mode)"; exit 0
fi
mkdir -p "$OBJ"
INC="-I$R/base/include -I$R/gpu/include -I$R/core/include -I$R/thirdparties/include/common -I$R/base"
COMMON="-std=c++11 -O2 -fPIC -DCOMPV_ASM=0 -flax-vector-conversions -fvisibility=hidden -w \
 -include limits -include cstddef -include cstdint $INC"
isa() { case "$(basename "$1")" in
  *_intrin_*sse2.cxx) echo "-msse2";; *_intrin_*ssse3.cxx) echo "-mssse3";; *_intrin_*sse41.cxx) echo "-msse4.1";;
  *_intrin_*sse42.cxx) echo "-msse4.2";; *_intrin_*avx2.cxx) echo "-mavx2 -mfma -D__FMA3__";;
  *_intrin_fma3_*avx.cxx) echo "-mavx -mfma -D__FMA3__";; *_intrin_*avx.cxx) echo "-mavx";; esac; }
list="$OBJ/cmds.txt"; : > "$list"
for pkg in base gpu core; do
  up=$(echo $pkg | tr a-z A-Z)
  while IFS= read -r f; do
    case "$f" in */intrin/arm/*|*/android/*|*/vs_android/*|*/compv_base_ml_knn.cxx) continue;; esac
    o="$OBJ/$(echo "${f#$R/}" | tr '/' '_').o"
    if [ ! -f "$o" ] || [ "$f" -nt "$o" ]; then
      echo "g++ $COMMON -DCOMPV_${up}_EXPORTS $(isa "$f") -c '$f' -o '$o' || echo 'FAILED $f' >&2" >> "$list"
    fi
  done < <(find "$R/$pkg" -name '*.cxx' -o -name '*.cpp' | sort)
done
n=$(wc -l < "$list"); echo "[build_ref] compiling $n reference translation units"
if [ "$n" -gt 0 ]; then xargs -P "$(nproc)" -I{} bash -c {} < "$list"; fi
g++ -shared -o "$OUT/libcompv_ref.so" "$OBJ"/*.o -ldl -lpthread
echo "[build_ref] linked $OUT/libcompv_ref.so"
if [ -f "$HERE/ref_shim.cxx" ]; then
  g++ -std=c++11 -O2 -fPIC -shared -w -include limits -include cstdint -DCOMPV_ASM=0 $INC \
     "$HERE/ref_shim.cxx" -o "$OUT/libcompv_refshim.so" -L"$OUT" -lcompv_ref -ldl -lpthread -Wl,-rpath,'$ORIGIN'
  echo "[build_ref] linked $OUT/libcompv_refshim.so"
fi

# The real CompVFeature::addFactory adapter (integration/compv_b200_plugin.cxx): needs the reference headers AND the built product library
PLUGIN_SRC="$HERE/../integration/compv_b200_plugin.cxx"
B200_LIB="$HERE/../compv_b200/lib"
if [ -f "$PLUGIN_SRC" ] && [ -f "$B200_LIB/libcompv_b200.so" ]; then
  g++ -std=c++11 -O2 -fPIC -shared -w -include limits -include cstdint -include cmath -DCOMPV_ASM=0 $INC -I"$HERE/../include" \
     "$PLUGIN_SRC" -o "$OUT/libcompv_b200_plugin.so" -L"$OUT" -lcompv_ref -L"$B200_LIB" -lcompv_b200 -ldl -lpthread \
     -Wl,-rpath,'$ORIGIN' -Wl,-rpath,'$ORIGIN/../../compv_b200/lib'
  echo "[build_ref] linked $OUT/libcompv_b200_plugin.so"
fi
