#!/usr/bin/env bash
# Builds the UNMODIFIED reference (chellmuth/pathed + its vendored Embree 3.6.0) headless,
# straight from the sources where they lie under /root/reference, into oracle/_ref/.
# Test infrastructure only: the product never links or executes anything built here.
#
# No reference build system is run: Embree's cmake is replaced by the explicit file lists
# below (transcribed from ext/embree/kernels/CMakeLists.txt:36-180 and common/*/CMakeLists.txt),
# its generated kernels/config.h by oracle/ref/embree_config.h.  Sources are reached through a
# symlink farm (cp -rs) so that "../config.h" and "random_generator.h" resolve to our shims
# without copying any reference source text.  The only generated source is scene.cpp with the
# two designated initialisers GCC 13 rejects rewritten as constructor calls (same order).
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
OUT=$(cd "$HERE/.." && pwd)/_ref
B=$OUT/build
JOBS=${JOBS:-$(nproc)}
[ -d "$REF/src" ] || { echo "reference not mounted at $REF; keeping prebuilt $OUT"; exit 0; }
mkdir -p "$B/obj"

# ---------------------------------------------------------------- symlink farms
if [ ! -d "$B/embree" ]; then
  mkdir -p "$B/embree"
  cp -rs "$REF/ext/embree/kernels" "$B/embree/kernels"
  cp -rs "$REF/ext/embree/common" "$B/embree/common"
  cp -rs "$REF/ext/embree/include" "$B/embree/include"
  rm -f "$B/embree/kernels/config.h" "$B/embree/kernels/hash.h"
  cp "$HERE/embree_config.h" "$B/embree/kernels/config.h"
  echo '#define RTC_HASH "pathed-b200-oracle"' > "$B/embree/kernels/hash.h"
fi
cp "$HERE/rtcore_config.h" "$B/embree/include/embree3/rtcore_config.h"
rm -rf "$B/pathed"; mkdir -p "$B/pathed"
cp -rs "$REF/include" "$B/pathed/include"
cp -rs "$REF/src" "$B/pathed/src"
rm -f "$B/pathed/include/random_generator.h" "$B/pathed/src/random_generator.cpp" "$B/pathed/src/scene.cpp"
cp "$HERE/random_generator.h" "$B/pathed/include/random_generator.h"
cp "$HERE/random_generator.cpp" "$B/pathed/src/random_generator.cpp"
python3 "$HERE/patch_scene.py" "$REF/src/scene.cpp" "$B/pathed/src/scene.cpp"

E=$B/embree
COMMON="-std=c++11 -O3 -DNDEBUG -fPIC -fno-strict-aliasing -fno-tree-vectorize -w \
 -DTASKING_INTERNAL -DEMBREE_TARGET_SSE2 -DEMBREE_TARGET_SSE42 -DEMBREE_TARGET_AVX -DEMBREE_TARGET_AVX2 \
 -I$E/include -I$E -I$E/common"
F_SSE2="-msse2 -DEMBREE_LOWEST_ISA"
F_SSE42="-msse4.2"
F_AVX="-mavx"
F_AVX2="-mf16c -mavx2 -mfma -mlzcnt -mbmi -mbmi2"

BASE="kernels/common/device.cpp kernels/common/stat.cpp kernels/common/acceln.cpp kernels/common/accelset.cpp
 kernels/common/state.cpp kernels/common/rtcore.cpp kernels/common/rtcore_builder.cpp kernels/common/scene.cpp
 kernels/common/alloc.cpp kernels/common/geometry.cpp kernels/common/scene_user_geometry.cpp
 kernels/common/scene_instance.cpp kernels/common/scene_triangle_mesh.cpp kernels/common/scene_quad_mesh.cpp
 kernels/common/scene_curves.cpp kernels/common/scene_line_segments.cpp kernels/common/scene_grid_mesh.cpp
 kernels/common/scene_points.cpp kernels/subdiv/bezier_curve.cpp kernels/subdiv/bspline_curve.cpp
 kernels/subdiv/catmullrom_curve.cpp kernels/geometry/primitive4.cpp kernels/geometry/instance_intersector.cpp
 kernels/geometry/curve_intersector_virtual.cpp kernels/builders/primrefgen.cpp kernels/bvh/bvh.cpp
 kernels/bvh/bvh_statistics.cpp kernels/bvh/bvh4_factory.cpp kernels/bvh/bvh8_factory.cpp kernels/bvh/bvh_rotate.cpp
 kernels/bvh/bvh_refit.cpp kernels/bvh/bvh_builder.cpp kernels/bvh/bvh_builder_hair.cpp kernels/bvh/bvh_builder_hair_mb.cpp
 kernels/bvh/bvh_builder_morton.cpp kernels/bvh/bvh_builder_sah.cpp kernels/bvh/bvh_builder_sah_spatial.cpp
 kernels/bvh/bvh_builder_sah_mb.cpp kernels/bvh/bvh_builder_twolevel.cpp kernels/bvh/bvh_intersector1_bvh4.cpp
 kernels/common/scene_subdiv_mesh.cpp kernels/subdiv/tessellation_cache.cpp kernels/subdiv/subdivpatch1base.cpp
 kernels/subdiv/catmullclark_coefficients.cpp kernels/geometry/grid_soa.cpp kernels/subdiv/subdivpatch1base_eval.cpp
 kernels/bvh/bvh_builder_subdiv.cpp kernels/bvh/bvh_intersector_hybrid4_bvh4.cpp kernels/bvh/bvh_intersector_stream_bvh4.cpp
 kernels/bvh/bvh_intersector_stream_filters.cpp
 common/sys/sysinfo.cpp common/sys/alloc.cpp common/sys/filename.cpp common/sys/library.cpp common/sys/thread.cpp
 common/sys/string.cpp common/sys/regression.cpp common/sys/mutex.cpp common/sys/condition.cpp common/sys/barrier.cpp
 common/simd/sse.cpp common/math/constants.cpp common/lexers/stringstream.cpp common/lexers/tokenstream.cpp
 common/tasking/taskschedulerinternal.cpp
 common/algorithms/parallel_for.cpp common/algorithms/parallel_reduce.cpp common/algorithms/parallel_prefix_sum.cpp
 common/algorithms/parallel_for_for.cpp common/algorithms/parallel_for_for_prefix_sum.cpp common/algorithms/parallel_partition.cpp
 common/algorithms/parallel_sort.cpp common/algorithms/parallel_set.cpp common/algorithms/parallel_map.cpp
 common/algorithms/parallel_filter.cpp"

ISA_ALL="kernels/geometry/instance_intersector.cpp kernels/geometry/curve_intersector_virtual.cpp kernels/bvh/bvh_intersector1_bvh4.cpp
 kernels/common/scene_subdiv_mesh.cpp kernels/geometry/grid_soa.cpp kernels/subdiv/subdivpatch1base_eval.cpp
 kernels/bvh/bvh_intersector_hybrid4_bvh4.cpp kernels/bvh/bvh_intersector_stream_bvh4.cpp kernels/bvh/bvh_intersector_stream_filters.cpp"
ISA_BUILDERS="kernels/common/scene_user_geometry.cpp kernels/common/scene_instance.cpp kernels/common/scene_triangle_mesh.cpp
 kernels/common/scene_quad_mesh.cpp kernels/common/scene_curves.cpp kernels/common/scene_line_segments.cpp
 kernels/common/scene_grid_mesh.cpp kernels/common/scene_points.cpp kernels/bvh/bvh_refit.cpp kernels/bvh/bvh_builder.cpp
 kernels/bvh/bvh_builder_hair.cpp kernels/bvh/bvh_builder_hair_mb.cpp kernels/bvh/bvh_builder_sah.cpp
 kernels/bvh/bvh_builder_sah_spatial.cpp kernels/bvh/bvh_builder_sah_mb.cpp kernels/bvh/bvh_builder_twolevel.cpp
 kernels/bvh/bvh_builder_subdiv.cpp kernels/bvh/bvh_builder_morton.cpp kernels/bvh/bvh_rotate.cpp kernels/builders/primrefgen.cpp"
ISA_WIDE="kernels/bvh/bvh_intersector1_bvh8.cpp kernels/bvh/bvh_intersector_hybrid8_bvh4.cpp kernels/bvh/bvh_intersector_hybrid4_bvh8.cpp
 kernels/bvh/bvh_intersector_hybrid8_bvh8.cpp kernels/bvh/bvh_intersector_stream_bvh8.cpp"
SSE42_FILES="$ISA_ALL"
AVX_FILES="$ISA_ALL $ISA_BUILDERS $ISA_WIDE kernels/geometry/primitive8.cpp kernels/bvh/bvh.cpp kernels/bvh/bvh_statistics.cpp"
AVX2_FILES="$ISA_ALL $ISA_BUILDERS $ISA_WIDE"

CMDS=$B/commands.txt; : > "$CMDS"
emit() { # isa flags files...
  local isa=$1 flags=$2; shift 2
  for f in "$@"; do
    local o="$B/obj/embree_${isa}_$(echo "$f" | tr '/.' '__').o"
    [ -f "$o" ] || echo "g++ $COMMON $flags -c $E/$f -o $o" >> "$CMDS"
  done
}
emit sse2 "$F_SSE2" $BASE
emit sse42 "$F_SSE42" $SSE42_FILES
emit avx "$F_AVX" $AVX_FILES
emit avx2 "$F_AVX2" $AVX2_FILES

# ---------------------------------------------------------------- pathed (no UI)
P=$B/pathed
UI="canvas gl_lines gl_points gl_scene rasterizer sample_widget screen shader pdf_widget path_visualization photon_renderer visualization ptex_local"
PFLAGS="-std=c++17 -O3 -DNDEBUG -fopenmp -fpermissive -w -fPIC -include cstdint -I$HERE/shims -I$P/include -I$REF/vendor -I$E/include -I$REF/ext/nanogui/ext/eigen"
for src in "$P"/src/*.cpp; do
  n=$(basename "$src" .cpp)
  case " $UI " in *" $n "*) continue;; esac
  o="$B/obj/pathed_$n.o"
  if [ ! -f "$o" ] || [ "$n" = scene ] || [ "$n" = random_generator ]; then
    echo "g++ $PFLAGS -c $src -o $o" >> "$CMDS"
  fi
done
echo "g++ $PFLAGS -c $HERE/shims/ptex_local_stub.cpp -o $B/obj/pathed_ptex_local_stub.o" >> "$CMDS"
for h in headless_main probe_main; do
  echo "g++ $PFLAGS -c $HERE/$h.cpp -o $B/obj/harness_$h.o" >> "$CMDS"
done
# the reference bound to libpathed_cuda: same reference objects, the Embree API served by a recording shim
REPO=$(cd "$HERE/../.." && pwd)
echo "g++ $PFLAGS -I$REPO/include -c $HERE/cuda_main.cpp -o $B/obj/harness_cuda_main.o" >> "$CMDS"
echo "g++ -std=c++17 -O2 -fPIC -w -I$E/include -c $HERE/embree_shim.cpp -o $B/obj/harness_embree_shim.o" >> "$CMDS"

echo "compiling $(wc -l < "$CMDS") translation units with $JOBS jobs"
xargs -P "$JOBS" -I{} bash -c '{} || { echo "FAILED: {}" >&2; exit 255; }' < "$CMDS"

ar rcs "$B/libembree_all.a" "$B"/obj/embree_*.o
PATHED_OBJS=$(ls "$B"/obj/pathed_*.o)
LINK="-fopenmp -Wl,--start-group $B/libembree_all.a -Wl,--end-group -lpthread -ldl"
g++ -o "$OUT/pathed_ref_headless" "$B/obj/harness_headless_main.o" $PATHED_OBJS $LINK
g++ -shared -o "$OUT/libpathed_ref_probe.so" "$B/obj/harness_probe_main.o" $PATHED_OBJS $LINK
if [ -f "$REPO/pathed_b200/libpathed_cuda.so" ]; then
  g++ -o "$OUT/pathed_ref_cuda" "$B/obj/harness_cuda_main.o" "$B/obj/harness_embree_shim.o" $PATHED_OBJS -fopenmp \
      -L"$REPO/pathed_b200" -lpathed_cuda -Wl,-rpath,'$ORIGIN/../../pathed_b200' -lpthread -ldl
  echo "built $OUT/pathed_ref_cuda"
else
  echo "pathed_b200/libpathed_cuda.so is not built yet: skipping $OUT/pathed_ref_cuda"
fi
echo "built $OUT/pathed_ref_headless and $OUT/libpathed_ref_probe.so"
