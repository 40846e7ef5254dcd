#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Stages a scratch copy of the reference's own sources (from where they lie under
# /root/reference) into a temporary directory OUTSIDE the repo and applies the mechanical, signature-only
# Linux/nvcc-12.9 patches of SURVEY.md Appendix A (MSVC-isms; zero arithmetic changes).  Nothing from the
# reference is ever written into the repo; only compiled binaries land in oracle/_ref/ (git-ignored).
#   usage: stage_ref.sh <reference_root> <work_dir>
set -euo pipefail
REF="$1"; WORK="$2"
rm -rf "$WORK"; mkdir -p "$WORK/src"
for f in area.h bbox.h bssrdf.h bvh.h bvh.cpp camera.h catmullrom.h common.h cutil_math.h imageio.h infinite.h \
         intersection.h line.h material.h medium.h mesh.h parsescene.h pathtracer.h pathtracer.cu primitive.h \
         ray.h sbvh.h scene.h sphere.h texture.h wrap.h; do
  cp "$REF/src/$f" "$WORK/src/$f"
done
cd "$WORK/src"
# A.1 backslash includes
sed -i 's#<thrust\\device_vector.h>#<thrust/device_vector.h>#; s#<glm\\glm.hpp>#<glm/glm.hpp>#; s#<glm\\gtc\\matrix_transform.hpp>#<glm/gtc/matrix_transform.hpp>#' common.h
# A.2 __debugbreak -> abort ; curand headers are unused
sed -i 's#__debugbreak();#abort();#; s#^\#include <curand.h>##; s#^\#include <curand_kernel.h>##' common.h
# A.3 rvalues bound to non-const references: add const (signature-only)
sed -i 's#Ray(float3& orig, float3& dir, Medium\* medium#Ray(const float3\& orig, const float3\& dir, Medium* medium#' ray.h
sed -i 's#BBox(float3& a, float3& b)#BBox(const float3\& a, const float3\& b)#; s#void Expand(BBox& b)#void Expand(const BBox\& b)#; s#void Expand(float3& v)#void Expand(const float3\& v)#' bbox.h
sed -i 's#operator-(float2 &a)#operator-(const float2 \&a)#; s#operator-(float3 &a)#operator-(const float3 \&a)#; s#operator-(float4 &a)#operator-(const float4 \&a)#' cutil_math.h
sed -i 's#void Pdf(Ray& ray, float3& nor, float& pdfA, float& pdfW) const#void Pdf(const Ray\& ray, const float3\& nor, float\& pdfA, float\& pdfW) const#; s#float3 Le(float3& nor, float3& dir) const#float3 Le(const float3\& nor, const float3\& dir) const#' area.h
sed -i 's#void Pdf(Ray& ray, float3& nor, float& pdfA, float& pdfW) const#void Pdf(const Ray\& ray, float3\& nor, float\& pdfA, float\& pdfW) const#; s#float3 Le(float3& dir) const#float3 Le(const float3\& dir) const#' infinite.h
sed -i 's#void PdfCamera(float3& dir,#void PdfCamera(const float3\& dir,#' camera.h
sed -i 's#void Phase(float3& in, float3& out, float& phase, float& pdf)#void Phase(const float3\& in, const float3\& out, float\& phase, float\& pdf)#; s#float getDensity(float3& p) const#float getDensity(const float3\& p) const#; s#float d(float3& p) const#float d(const float3\& p) const#' medium.h
# A.4 MSVC functional casts
sed -i 's#unsigned char(\(rgba\.[xyzw] \* 255\))#(unsigned char)(\1)#g' texture.h
# A.5 bvh.cpp:168 unsequenced ++next: make the (MSVC/g++ right-to-left) order explicit
sed -i 's#flatten(node->right, next, ++next);#{ int cur_r = next + 1; ++next; flatten(node->right, cur_r, next); }#' bvh.cpp

cd "$WORK/src"
# A.3 (cont.) dead code (SingleScatter) and the SPPM host grid (other integrator): named temporaries / const
sed -i 's#if (Intersect(Ray(pos,rdir,nullptr,kernel_epsilon), &rIsect)){#Ray rtmp_(pos,rdir,nullptr,kernel_epsilon); if (Intersect(rtmp_, \&rIsect)){#' pathtracer.cu
sed -i 's#Intersect(Ray(pos, tdir, nullptr, kernel_hdr_height), &tIsect);#Ray ttmp_(pos, tdir, nullptr, kernel_hdr_height); Intersect(ttmp_, \&tIsect);#' pathtracer.cu
sed -i 's#bool ToGrid(float3& p, BBox& bounds#bool ToGrid(const float3\& p, BBox\& bounds#' pathtracer.cu
echo "staged reference sources in $WORK/src"
