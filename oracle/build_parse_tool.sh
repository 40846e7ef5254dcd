#!/usr/bin/env bash
# TEST INFRASTRUCTURE — compiles the reference's own scene parser (src/parsescene.cpp, src/mesh.cpp, src/imageio.cpp,
# src/bvh.cpp + headers; staged OUTSIDE the repo with the mechanical patches of oracle/stage_ref.sh plus the backslash
# includes of these three files) behind oracle/refbuild/parse_tool.cpp, which stands in for the one absent dependency
# (libassimp's Importer::ReadFile).  Only the binary lands in oracle/_ref/ (git-ignored).  This container only.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
WORK=$(mktemp -d /tmp/b200pt_parse.XXXXXX)
bash "$HERE/stage_ref.sh" "$REF" "$WORK" > /dev/null
cp "$REF/src/imageio.cpp" "$REF/src/tinyexr.h" "$REF/src/parsescene.cpp" "$REF/src/mesh.cpp" "$WORK/src/"
cd "$WORK/src"
sed -i 's#<stb\\stb_image.h>#<stb/stb_image.h>#; s#<stb\\stb_image_write.h>#<stb/stb_image_write.h>#; s#^\#define STBI_MSC_SECURE_CRT##' imageio.cpp
sed -i 's#<rapidjson\\include\\rapidjson\\\([a-z]*\).h>#<rapidjson/include/rapidjson/\1.h>#; s#<sys\\stat.h>#<sys/stat.h>#' parsescene.cpp
mkdir -p "$HERE/_ref"
g++ -O2 -w -std=c++14 -ffp-contract=off -fpermissive -I"$WORK/src" -I"$REF/include" -I/usr/local/cuda/include \
    "$HERE/refbuild/parse_tool.cpp" parsescene.cpp mesh.cpp imageio.cpp bvh.cpp -o "$HERE/_ref/parse_tool" \
    -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
rm -rf "$WORK"
echo "built $HERE/_ref/parse_tool"
