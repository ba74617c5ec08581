#!/usr/bin/env bash
# Build oracle/_ref: the reference's OWN hot-path sources, compiled where they lie.
# Test infrastructure only -- nothing under oracle/ is linked into the product.
#
# - sources are read from $GCM_REFERENCE (default /root/reference); they are NEVER copied
#   into git: the patched working copy lives under oracle/_ref/src (git-ignored)
# - 2 one-line patches for missing-return UB, needed with g++ >= 8:
#     gcm/datatypes/Node.h:136           bool addOwner -> void addOwner
#     gcm/meshtypes/TetrMesh_1stOrder.cpp:478  add `return 0;` at the end of find_border_node_normal
# - 1 tap hook after the owner search in find_nodes_on_previous_time_layer
#   (gcm/methods/TetrMethod_Plastic_1stOrder.cpp:590), compiled in ONLY with -DGCM_ORACLE_TAP
# - the reference's build system (cmake + VTK/GSL/TinyXML/MPI, all absent) is not used;
#   stubs/ provides the headers, databus_stub.cpp the one-rank DataBus, driver.cpp main().
# Outputs: oracle/_ref/gcm3d_ref (timed build, no tap), oracle/_ref/gcm3d_ref_tap and
# oracle/_ref/gcm3d_ref_kat (known-answer driver, ref_build/kat_driver.cpp).
set -euo pipefail
HERE="$(cd "$(dirname "${BASH_SOURCE[0]}")" && pwd)"
REF="${GCM_REFERENCE:-/root/reference}"
OUT="$HERE/_ref"
SRC="$OUT/src"
CXX="${CXX:-g++}"
JOBS="${JOBS:-$(nproc)}"

if [ ! -d "$REF/gcm" ]; then
	echo "build_ref.sh: $REF/gcm not found (GPU box?) -- keeping prebuilt oracle/_ref" >&2
	exit 0
fi

rm -rf "$SRC" "$OUT/obj" "$OUT/obj_tap"
mkdir -p "$SRC" "$OUT/obj" "$OUT/obj_tap"
cp -r "$REF/gcm" "$SRC/gcm"
chmod -R u+w "$SRC"

sed -i 's/bool inline addOwner (EngineOwner owner)/void inline addOwner (EngineOwner owner)/' "$SRC/gcm/datatypes/Node.h"
# end of find_border_node_normal: the line after the last throw is the closing brace at :479
python3 - "$SRC/gcm/meshtypes/TetrMesh_1stOrder.cpp" "$SRC/gcm/methods/TetrMethod_Plastic_1stOrder.cpp" <<'EOF'
import sys, re
mesh, meth = sys.argv[1], sys.argv[2]
s = open(mesh).read()
key = 'throw GCMException( GCMException::MESH_EXCEPTION, "Can not create reasonable normal for node - is the border too sharp?");\n}'
assert s.count(key) == 1
s = s.replace(key, key[:-1] + '\treturn 0;\n}')
open(mesh, 'w').write(s)
s = open(meth).read()
key = 'tmp_tetr = mesh->find_owner_tetr(cur_node, dx[0], dx[1], dx[2], debug);\n'
assert s.count(key) == 1
s = s.replace(key, key + '#ifdef GCM_ORACLE_TAP\n\t\t\tgcm_oracle_tap(cur_node, mesh, stage, i, count, alpha, dx, tmp_tetr);\n#endif\n')
s = s.replace('#include "TetrMethod_Plastic_1stOrder.h"\n',
	'#include "TetrMethod_Plastic_1stOrder.h"\n#ifdef GCM_ORACLE_TAP\nvoid gcm_oracle_tap(ElasticNode*, TetrMesh*, int, int, int, float, float*, Tetrahedron*);\n#endif\n', 1)
open(meth, 'w').write(s)
EOF

TUS="datatypes/matrixes datatypes/ElasticMatrix3D datatypes/ElasticNode datatypes/Element datatypes/Node
datatypes/Tetrahedron datatypes/Tetrahedron_1st_order datatypes/Triangle
system/quick_math system/GCMException system/Logger system/LoggerUser system/Utils system/TetrMeshSet
system/CollisionDetector system/VoidCollisionDetector system/BruteforceCollisionDetector system/Stresser
system/BorderCondition system/ContactCondition system/SPHConnector
system/forms/PulseForm system/forms/StepPulseForm system/areas/Area system/areas/BoxArea
methods/NumericalMethod methods/TetrNumericalMethod methods/TetrMethod_Plastic_1stOrder
methods/VolumeCalculator methods/BorderCalculator methods/ContactCalculator
methods/volume/SimpleVolumeCalculator
methods/border/FreeBorderCalculator methods/border/FixedBorderCalculator methods/border/ExternalForceCalculator
methods/border/ExternalVelocityCalculator methods/border/ExternalValuesCalculator
methods/contact/AdhesionContactCalculator methods/contact/FrictionContactCalculator
meshtypes/Mesh meshtypes/TetrMesh meshtypes/TetrMesh_1stOrder
rheotypes/RheologyCalculator rheotypes/VoidRheologyCalculator"

# The reference is built with plain -O3 (gcm/CMakeLists.txt:76); gnu++98 because
# Logger.cpp:22 relies on the C++98 stream->bool conversion.
REFFLAGS="-std=gnu++98 -O3 -w -fpermissive -I$HERE/ref_build/stubs -I$SRC/gcm"
OWNFLAGS="-std=gnu++11 -O2 -w -fpermissive -I$HERE/ref_build/stubs -I$SRC/gcm"

compile_one() { # $1 = tu, $2 = objdir, $3 = extra flags
	local tu="$1" obj="$2/$(echo "$1" | tr '/' '_').o"
	[ -f "$SRC/gcm/$tu.cpp" ] || return 0
	$CXX $REFFLAGS $3 -c "$SRC/gcm/$tu.cpp" -o "$obj"
}
export -f compile_one
export CXX REFFLAGS SRC
echo $TUS | tr ' ' '\n' | xargs -P "$JOBS" -I{} bash -c "compile_one {} $OUT/obj ''"
# only one TU differs in the tap build
cp "$OUT"/obj/*.o "$OUT/obj_tap/"
compile_one methods/TetrMethod_Plastic_1stOrder "$OUT/obj_tap" "-DGCM_ORACLE_TAP"

$CXX $OWNFLAGS -c "$HERE/ref_build/databus_stub.cpp" -o "$OUT/obj/databus_stub.o"
cp "$OUT/obj/databus_stub.o" "$OUT/obj_tap/"
$CXX $OWNFLAGS -c "$HERE/ref_build/driver.cpp" -o "$OUT/obj/driver.o"
$CXX $OWNFLAGS -DGCM_ORACLE_TAP -c "$HERE/ref_build/driver.cpp" -o "$OUT/obj_tap/driver.o"
$CXX -O2 "$OUT"/obj/*.o -o "$OUT/gcm3d_ref"
$CXX -O2 "$OUT"/obj_tap/*.o -o "$OUT/gcm3d_ref_tap"
# known-answer driver: the same reference objects behind kat_driver.cpp's main()
$CXX $OWNFLAGS -c "$HERE/ref_build/kat_driver.cpp" -o "$OUT/kat_driver.o"
$CXX -O2 $(ls "$OUT"/obj/*.o | grep -v '/driver.o$') "$OUT/kat_driver.o" -o "$OUT/gcm3d_ref_kat"
rm -f "$OUT/kat_driver.o"

# gcm3d_b200: the same reference classes and the same driver, with TetrMeshSet::do_next_step and
# TetrMesh_1stOrder::do_next_part_step replaced by the INTEGRATION.md shim (ref_build/shim/) and linked against
# libgcm_b200.so -- the compiled drop-in proof.  The reference's own two bodies are kept under another name
# (-D below renames them consistently in every reference TU; the shim and the driver see the real names).
LIBDIR="$HERE/../gcm-necro_b200/gcm_b200"
if [ -f "$LIBDIR/libgcm_b200.so" ]; then
	mkdir -p "$OUT/obj_b200"
	RENAME="-Ddo_next_step=do_next_step_host -Ddo_next_part_step=do_next_part_step_host"
	echo $TUS | tr ' ' '\n' | xargs -P "$JOBS" -I{} bash -c "compile_one {} $OUT/obj_b200 '$RENAME'"
	cp "$OUT/obj/databus_stub.o" "$OUT/obj/driver.o" "$OUT/obj_b200/"
	$CXX $OWNFLAGS -I"$HERE/../include" -c "$HERE/ref_build/shim/tetrmeshset_b200.cpp" -o "$OUT/obj_b200/shim.o"
	$CXX -O2 "$OUT"/obj_b200/*.o -L"$LIBDIR" -lgcm_b200 -ldl -Wl,-rpath,'$ORIGIN/../../gcm-necro_b200/gcm_b200' -o "$OUT/gcm3d_b200"
	rm -rf "$OUT/obj_b200"
	echo "built $OUT/gcm3d_b200 (reference classes + shim + libgcm_b200.so)"
else
	echo "build_ref.sh: libgcm_b200.so not built yet -- skipping gcm3d_b200" >&2
fi
rm -rf "$OUT/obj" "$OUT/obj_tap"
echo "built $OUT/gcm3d_ref and $OUT/gcm3d_ref_tap"
