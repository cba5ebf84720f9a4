#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.  Builds oracle/_ref/libref_pt*.so from the reference sources WHERE THEY
# LIE under /root/reference (nothing is copied into the repo; oracle/_ref/ is git-ignored but does
# travel to the GPU box).  Three variants of the same unity build (ref_driver.cu + pathtrace.cu):
#   libref_pt.so        nvcc defaults (-fmad=true), i.e. how the reference's CMake builds it
#   libref_pt_nofma.so  -fmad=false, the association-order-only variant (DESIGN.md "strict math")
#   libref_pt_sort.so   SORT_MATERIAL flipped to true (pathtrace.cu:21) through a sed'ed temp copy
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/_ref
SRC=$REF/Inference/src
INC="-I$SRC -I$REF/Inference/external/include"
[ -f "$SRC/pathtrace.cu" ] || { echo "reference not present at $REF - keeping prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$OUT"
TMP=$(mktemp -d)
trap 'rm -rf "$TMP"' EXIT
ARCH="-gencode arch=compute_100a,code=sm_100a"
NVCC="nvcc -std=c++14 -O2 -DNDEBUG $ARCH -Xcompiler -fPIC,-fopenmp -diag-suppress 20012,20011,20014,550,177,2361 -w"
# Scene::loadObj (scene.cpp:206-320) is declared int but falls off its end without a return; g++ >= 8
# compiles that to a trap (ud2), so every MESH scene would abort.  Build scene.cpp from a temp copy whose
# only change is a `return 1;` before that function's closing brace (the file's last line).
sed '$ s/^}[[:space:]]*$/\treturn 1;\n}/' "$SRC/scene.cpp" > "$TMP/scene.cpp"
[ "$(diff "$SRC/scene.cpp" "$TMP/scene.cpp" | grep -c '^[<>]')" = "1" ] || { echo "scene.cpp patch did not apply as expected" >&2; exit 1; }
g++ -O2 -std=c++14 -fPIC -fpermissive -w $INC -I/usr/local/cuda/include -c "$TMP/scene.cpp" -o "$TMP/scene.o"
g++ -O2 -std=c++14 -fPIC -fpermissive -w $INC -I/usr/local/cuda/include -c "$SRC/utilities.cpp" -o "$TMP/utilities.o"
g++ -O2 -std=c++14 -fPIC -fpermissive -w $INC -I/usr/local/cuda/include -c "$HERE/ref_glue.cpp" -o "$TMP/glue.o"
HOSTOBJ="$TMP/scene.o $TMP/utilities.o $TMP/glue.o"
build() { # name, extra nvcc flags, include dir that holds pathtrace.cu
  $NVCC $2 -I"$3" $INC -c "$HERE/ref_driver.cu" -o "$TMP/$1.o"
  nvcc -shared $ARCH -Xcompiler -fopenmp -o "$OUT/$1.so" "$TMP/$1.o" $HOSTOBJ -lgomp
  echo "built $OUT/$1.so"
}
build libref_pt "" "$SRC" &
build libref_pt_nofma "-fmad=false" "$SRC" &
mkdir -p "$TMP/sort"
sed 's/^#define SORT_MATERIAL false/#define SORT_MATERIAL true/' "$SRC/pathtrace.cu" > "$TMP/sort/pathtrace.cu"
grep -q '^#define SORT_MATERIAL true' "$TMP/sort/pathtrace.cu"
build libref_pt_sort "" "$TMP/sort" &
wait
