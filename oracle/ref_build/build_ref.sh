#!/usr/bin/env bash
# build_ref.sh — compile the REFERENCE's own hot-path translation units, from where they
# lie under /root/reference/src, into oracle/_ref/libsmilei_ref.so.
#
# TEST INFRASTRUCTURE ONLY.  No reference source is copied into this repository: the
# compiler reads the files in place.  The full `smilei` binary is unbuildable here (no MPI,
# no HDF5, generated Python headers), but the operator classes on the hot path compile
# stand-alone once <mpi.h>/<hdf5.h> are replaced by the declaration-only stand-ins in
# oracle/ref_build/stubs/.  Symbols those objects reference but the hot path never calls
# (diagnostics, I/O, MPI wrappers ...) are resolved to a trapping stub generated below.
#
# Flags: -O2 -ffp-contract=off => plain IEEE-754 double arithmetic, no FMA contraction, i.e.
# the arithmetic the source text states.  REF_OPT overrides (e.g. "-O3 -march=native" for
# the timed CPU baseline library libsmilei_ref_fast.so).
set -euo pipefail
REF=${SMILEI_REFERENCE:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$HERE/../_ref
NAME=${REF_NAME:-libsmilei_ref.so}
OPT=${REF_OPT:--O2 -ffp-contract=off}
if [ ! -d "$REF/src" ]; then
    echo "build_ref.sh: $REF/src not present - keeping prebuilt $OUT/$NAME if any" >&2
    exit 0
fi
mkdir -p "$OUT/obj_$NAME"
OBJ=$OUT/obj_$NAME
INC="-I$HERE/stubs"
for d in "$REF"/src/*/; do INC="$INC -I$d"; done
PYINC=$(python3-config --includes)
CXXFLAGS="-std=c++14 -fPIC -fopenmp -D_OMP -DOMPI_SKIP_MPICXX -w $OPT $INC $PYINC"

SRCS="
Interpolator/Interpolator3D.cpp
Interpolator/Interpolator3D2Order.cpp
Interpolator/Interpolator3D4Order.cpp
Pusher/Pusher.cpp
Pusher/PusherBoris.cpp
Pusher/PusherVay.cpp
Pusher/PusherHigueraCary.cpp
Projector/Projector.cpp
Projector/Projector3D.cpp
Projector/Projector3D2Order.cpp
Projector/Projector3D4Order.cpp
Interpolator/Interpolator3D2OrderV.cpp
Projector/Projector3D2OrderV.cpp
ElectroMagnSolver/MA_Solver3D_norm.cpp
ElectroMagnSolver/MF_Solver3D_Yee.cpp
ElectroMagn/ElectroMagn3D.cpp
Field/Field.cpp
Field/Field3D.cpp
Tools/gpu.cpp
SmileiMPI/AsyncMPIbuffers.cpp
Tools/Tools.cpp
Particles/Particles.cpp
Species/SpeciesV.cpp
ParticleBC/BoundaryConditionType.cpp
ElectroMagnBC/ElectroMagnBC.cpp
ElectroMagnBC/ElectroMagnBC3D.cpp
ElectroMagnBC/ElectroMagnBC3D_SM.cpp
Field/Field2D.cpp
Particles/ParticleCreator.cpp
DomainDecomposition/Hilbert_functions.cpp
"
pids=()
for s in $SRCS; do
    o=$OBJ/$(echo "$s" | tr '/' '_' | sed 's/\.cpp$/.o/')
    if [ ! -f "$o" ] || [ "$REF/src/$s" -nt "$o" ]; then
        g++ $CXXFLAGS -c "$REF/src/$s" -o "$o" &
        pids+=($!)
    fi
done
g++ $CXXFLAGS -c "$HERE/ref_harness.cpp" -o "$OBJ/ref_harness.o" &
pids+=($!)
g++ $CXXFLAGS -c "$HERE/ref_creator_harness.cpp" -o "$OBJ/ref_creator_harness.o" &
pids+=($!)
for p in "${pids[@]}"; do wait "$p" || { echo "build_ref: compile failed" >&2; exit 1; }; done

# First link: discover what is still undefined.
rm -f "$OBJ/unresolved_stubs.o"
g++ -shared -fopenmp -o "$OUT/$NAME.tmp" "$OBJ"/*.o
python3 "$HERE/gen_stubs.py" "$OUT/$NAME.tmp" "$OBJ/unresolved_stubs.s"
gcc -c "$OBJ/unresolved_stubs.s" -o "$OBJ/unresolved_stubs.o"
g++ -shared -fopenmp -Wl,-z,defs -o "$OUT/$NAME" "$OBJ"/*.o -lm
rm -f "$OUT/$NAME.tmp"
echo "build_ref: wrote $OUT/$NAME"
