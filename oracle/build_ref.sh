#!/usr/bin/env bash
# TEST INFRASTRUCTURE ONLY.
# Compiles the UNMODIFIED reference sources where they lie under $REF (default
# /root/reference) plus the extern "C" gateways of this directory into
# oracle/_ref/libref.so.  Nothing from the reference is copied into the repo;
# oracle/_ref/ is git-ignored (but travels to the GPU box with the snapshot).
# Flags follow the reference's own build lines (python/CMakeLists.txt:33,
# octave/compile_mex.m:41): -std=c++11 -O3 -fopenmp -DNDEBUG.
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${REF:-/root/reference}"
OUT="$HERE/_ref"
if [ ! -d "$REF/src" ]; then
    echo "build_ref.sh: $REF/src not present - keeping any prebuilt $OUT/libref.so" >&2
    exit 0
fi
mkdir -p "$OUT/obj"
CXXFLAGS="-std=c++11 -O3 -fopenmp -DNDEBUG -fPIC -w"
pids=()
objs=()
for f in PFDR_graph_quadratic_d1_l1 PFDR_graph_quadratic_d1_bounds PFDR_graph_loss_d1_simplex \
         proj_simplex_metric operator_norm_matrix CP_PFDR_graph_quadratic_d1_l1 \
         CP_PFDR_graph_quadratic_d1_l1_duplex CP_PFDR_graph_quadratic_d1_bounds \
         CP_PFDR_graph_loss_d1_simplex graph maxflow; do
    g++ $CXXFLAGS -c "$REF/src/$f.cpp" -o "$OUT/obj/$f.o" & pids+=($!)
    objs+=("$OUT/obj/$f.o")
done
for w in ref_wrap_l1 ref_wrap_bounds ref_wrap_misc; do
    g++ $CXXFLAGS -I"$REF/include" -c "$HERE/$w.cpp" -o "$OUT/obj/$w.o" & pids+=($!)
    objs+=("$OUT/obj/$w.o")
done
for p in "${pids[@]}"; do wait "$p"; done
g++ -shared -fopenmp -Wl,-Bsymbolic-functions -o "$OUT/libref.so" "${objs[@]}"
echo "built $OUT/libref.so"

# Drop-in check: the reference's UNMODIFIED cut-pursuit objects (CP_*.cpp, graph.cpp, maxflow.cpp) linked
# against the B200 shim (cp_pfdr_graph_d1_b200/lib/libpfdr_b200_shim.so) instead of the reference's PFDR objects.
#   libcp_b200.so          every hot-path routine from libpfdr_b200 (incl. operator_norm_matrix)
#   libcp_b200_refnorm.so  the reference's own operator_norm_matrix kept (isolates the PFDR solvers)
SHIMDIR="$HERE/../cp_pfdr_graph_d1_b200/lib"
if [ -f "$SHIMDIR/libpfdr_b200_shim.so" ]; then
    cp_objs=()
    for f in CP_PFDR_graph_quadratic_d1_l1 CP_PFDR_graph_quadratic_d1_l1_duplex CP_PFDR_graph_quadratic_d1_bounds \
             CP_PFDR_graph_loss_d1_simplex graph maxflow ref_wrap_l1 ref_wrap_bounds ref_wrap_misc; do
        cp_objs+=("$OUT/obj/$f.o")
    done
    # the product's own gateway for the Python module (cp_pfdr_graph_d1_b200/libCP.py) rides in the same library
    g++ $CXXFLAGS -I"$REF/include" -c "$HERE/../cp_pfdr_graph_d1_b200/csrc/cp_gateway.cpp" -o "$OUT/obj/cp_gateway.o"
    cp_objs+=("$OUT/obj/cp_gateway.o")
    LINK=(-shared -fopenmp -Wl,-Bsymbolic-functions -L"$SHIMDIR" -lpfdr_b200_shim -lpfdr_b200
          -Wl,-rpath,'$ORIGIN/../../cp_pfdr_graph_d1_b200/lib')
    g++ -o "$OUT/libcp_b200.so" "${cp_objs[@]}" "${LINK[@]}"
    g++ -o "$OUT/libcp_b200_refnorm.so" "${cp_objs[@]}" "$OUT/obj/operator_norm_matrix.o" "${LINK[@]}"
    echo "built $OUT/libcp_b200.so $OUT/libcp_b200_refnorm.so"
fi
