#!/usr/bin/env bash
# TEST INFRASTRUCTURE — compiles oracle/refbuild/exr_tool.cpp against the reference's vendored EXR code where it lies
# (REF/src/tinyexr.h); only the binary lands in oracle/_ref/ (git-ignored).  Needs /root/reference: this container only.
set -euo pipefail
REF=${REF:-/root/reference}
HERE=$(cd "$(dirname "$0")" && pwd)
mkdir -p "$HERE/_ref"
g++ -O2 -w -std=c++11 -I"$REF/src" "$HERE/refbuild/exr_tool.cpp" -o "$HERE/_ref/exr_tool"
# ... and the image decoder it links (REF/include/stb/stb_image.h): texels as Texture::Texture would hold them
g++ -O2 -w -std=c++11 -ffp-contract=off -I"$REF/include" "$HERE/refbuild/tex_tool.cpp" -o "$HERE/_ref/tex_tool"
# ... and the GLM it vendors (REF/include/glm): the scene parser's transform arithmetic (oracle/refbuild/glm_tool.cpp)
g++ -O2 -w -std=c++11 -ffp-contract=off -I"$REF/include" "$HERE/refbuild/glm_tool.cpp" -o "$HERE/_ref/glm_tool"
echo "built $HERE/_ref/exr_tool $HERE/_ref/tex_tool $HERE/_ref/glm_tool"
