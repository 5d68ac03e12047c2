#!/usr/bin/env bash
# build_adapter.sh — link the reference's own objects (built by build_ref.sh), the adapter header of this repository
# (include/smilei_b200_operators.hpp) and the product library into oracle/_ref/libsmilei_adapter.so, the library
# tests/test_gpu_parity.py::test_adapter_executes_through_reference_vtable loads on the GPU box.
# TEST INFRASTRUCTURE ONLY; no reference source is copied, the compiler reads the headers in place.
set -euo pipefail
REF=${SMILEI_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
OUT=$HERE/../_ref
OBJ=$OUT/obj_libsmilei_ref.so
if [ ! -d "$REF/src" ] || [ ! -d "$OBJ" ]; then
    echo "build_adapter.sh: reference sources or objects not present - keeping prebuilt $OUT/libsmilei_adapter.so if any" >&2
    exit 0
fi
INC="-I$HERE/stubs -I$HERE -I$ROOT/include"
for d in "$REF"/src/*/; do INC="$INC -I$d"; done
PYINC=$(python3-config --includes)
CXXFLAGS="-std=c++14 -fPIC -fopenmp -D_OMP -DOMPI_SKIP_MPICXX -w -O2 -ffp-contract=off $INC $PYINC"
A=$OUT/obj_adapter
mkdir -p "$A"
g++ $CXXFLAGS -c "$HERE/adapter_harness.cpp" -o "$A/adapter_harness.o"
OBJS=$(ls "$OBJ"/*.o | grep -v -e ref_harness.o -e ref_creator_harness.o -e unresolved_stubs.o)
LIBDIR=$ROOT/smilei_b200/csrc
LINK="-L$LIBDIR -lsmilei_b200 -Wl,-rpath,\$ORIGIN/../../smilei_b200/csrc -lm -ldl -lpthread -lrt"
# first link: what is still undefined goes to the trapping stub (same mechanism as build_ref.sh)
g++ -shared -fopenmp -o "$OUT/libsmilei_adapter.so.tmp" "$A/adapter_harness.o" $OBJS $LINK
python3 "$HERE/gen_stubs.py" "$OUT/libsmilei_adapter.so.tmp" "$A/unresolved_stubs.s" "$LIBDIR/libsmilei_b200.so"
gcc -c "$A/unresolved_stubs.s" -o "$A/unresolved_stubs.o"
g++ -shared -fopenmp -Wl,-z,defs -o "$OUT/libsmilei_adapter.so" "$A/adapter_harness.o" "$A/unresolved_stubs.o" $OBJS $LINK
rm -f "$OUT/libsmilei_adapter.so.tmp"
echo "build_adapter: wrote $OUT/libsmilei_adapter.so"
