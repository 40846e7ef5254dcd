#!/usr/bin/env bash
# TEST INFRASTRUCTURE — compiles the reference's own src/imageio.cpp + src/texture.h (staged OUTSIDE the repo with the
# mechanical patches of oracle/stage_ref.sh, plus imageio.cpp's backslash includes and its MSVC-only STBI_MSC_SECURE_CRT)
# behind oracle/refbuild/imageio_tool.cpp; only the binary lands in oracle/_ref/ (git-ignored).  This container only.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
WORK=$(mktemp -d /tmp/b200pt_imageio.XXXXXX)
bash "$HERE/stage_ref.sh" "$REF" "$WORK" > /dev/null
cp "$REF/src/imageio.cpp" "$REF/src/tinyexr.h" "$WORK/src/"
sed -i 's#<stb\\stb_image.h>#<stb/stb_image.h>#; s#<stb\\stb_image_write.h>#<stb/stb_image_write.h>#; s#^\#define STBI_MSC_SECURE_CRT##' "$WORK/src/imageio.cpp"
mkdir -p "$HERE/_ref"
g++ -O2 -w -std=c++14 -ffp-contract=off -I"$WORK/src" -I"$REF/include" -I/usr/local/cuda/include \
    "$HERE/refbuild/imageio_tool.cpp" "$WORK/src/imageio.cpp" -o "$HERE/_ref/imageio_tool" \
    -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
rm -rf "$WORK"
echo "built $HERE/_ref/imageio_tool"
