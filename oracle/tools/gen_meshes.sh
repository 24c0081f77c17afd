#!/bin/bash
# TEST INFRASTRUCTURE ONLY. Regenerates tests/golden/meshes/*.f32 from the reference's res/*.obj through the
# reference's own obj_parse (build container only; needs /root/reference). Usage: oracle/tools/gen_meshes.sh
set -e
cd "$(dirname "$0")/.."
make meshtool
OUT=../tests/golden/meshes
mkdir -p $OUT
for m in cube floor ico ramp lever lever_support seesaw_support mirror_cube_collider1 mirror_cube_collider2 cylinder; do
  ./_ref/dump_mesh /root/reference/res/$m.obj $OUT/$m.f32 > /dev/null
done
ls -la $OUT
