#!/usr/bin/env bash
# TEST / BASELINE INFRASTRUCTURE.  Compiles the reference's own G-PT integrator with everything it runs on (see the
# comment above _ref/libref_mitsuba.so in oracle/Makefile) from the sources where they lie under $REF into one shared
# library.  Usage: build_ref_mitsuba.sh <out.so> "<optimisation flags>" "<plugins>" "<core sources>"
#   parity build   : -O2 -ffp-contract=off            (same unfused operation order as the restatement and the GPU tracer)
#   baseline build : -O3 -march=x86-64-v3             (the CPU baseline bench.py times; BASELINE.md §3 asks for -O3 -march=native,
#                                                      x86-64-v3 = AVX2+FMA is the part of it that is safe on every GPU box's host)
# Every compile is checked: a failing translation unit fails the build (no stale objects are linked).
set -euo pipefail
OUT=$1; OPT=$2; PLUGINS=$3; CORE=$4
REF=${REF:-/root/reference}
CXX=${CXX:-/usr/bin/g++}
PS=$REF/src/integrators/poisson_solver
TMP=$(dirname "$OUT")/.mitsuba_objs_$(basename "$OUT" .so)
rm -rf "$TMP"; mkdir -p "$TMP"
trap 'rm -rf "$TMP"' EXIT
FLAGS="-std=gnu++11 $OPT -fpermissive -w -fPIC -include unistd.h -include cassert -Irefstubs -I$REF/include -DDOUBLE_PRECISION -DSPECTRUM_SAMPLES=3 -DMTS_NO_STATISTICS"
PS_FLAGS="-std=c++14 $OPT -fopenmp -fpermissive -w -fPIC -Ishim -I$PS -include shim/compat.h"
JOBS=${JOBS:-$(nproc)}
pids=()
run() { "$@" & pids+=($!); if [ ${#pids[@]} -ge "$JOBS" ]; then wait "${pids[0]}"; pids=("${pids[@]:1}"); fi; }
drain() { for p in "${pids[@]}"; do wait "$p"; done; pids=(); }
for b in $PLUGINS; do n=$(basename "$b")
    run $CXX $FLAGS -DCreateInstance=CreateInstance_$n -DGetDescription=GetDescription_$n -c "$REF/src/$b.cpp" -o "$TMP/plugin_$n.o"
done
for c in $CORE; do run $CXX $FLAGS -c "$REF/src/$c.cpp" -o "$TMP/$(basename "$c").o"; done
# gpt.cpp jumps over initialisations with `goto half_vector_shift_failed` (MSVC accepts that, ISO C++ does not): wrap the region
# in do { ... } while (0) and turn the gotos into `break` (no loop or switch lies in between: same control flow)
sed -e 's|// Deny shifts between Dirac and non-Dirac BSDFs.|do {|' -e 's|^half_vector_shift_failed:|} while (0);|' \
    -e 's|goto half_vector_shift_failed;|break;|' "$REF/src/integrators/gpt/gpt.cpp" > "$TMP/gpt_iso.cpp"
run $CXX $FLAGS -Ishim -include shim/compat.h -I"$REF/src/integrators/gpt" -DCreateInstance=CreateInstance_gpt -DGetDescription=GetDescription_gpt \
    -c "$TMP/gpt_iso.cpp" -o "$TMP/plugin_gpt.o"
run $CXX $FLAGS -DCreateInstance=CreateInstance_gdb200_counter -DGetDescription=GetDescription_gdb200_counter \
    -c ../gradientdomain-mitsuba_b200/plugin/samplers/gdb200_counter.cpp -o "$TMP/plugin_gdb200_counter.o"
for f in Solver Backend BackendOpenMP Defs; do run $CXX $PS_FLAGS -c "$PS/$f.cpp" -o "$TMP/ps_$f.o"; done
run $CXX $FLAGS -c refstubs/ref_support.cpp -o "$TMP/ref_support.o"
run $CXX $FLAGS -c ref_mitsuba_shim.cpp -o "$TMP/ref_mitsuba_shim.o"
run $CXX $FLAGS -Ishim -include shim/compat.h -c ref_gpt_shim.cpp -o "$TMP/ref_gpt_shim.o"
drain
# Statistics::m_instance before any static StatsCounter
$CXX -shared -fopenmp -Wl,-z,defs -o "$OUT" "$TMP/statistics.o" $(ls "$TMP"/*.o | grep -v /statistics.o) -lz -lpthread
