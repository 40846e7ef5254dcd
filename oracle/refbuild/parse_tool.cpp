// TEST INFRASTRUCTURE — the reference's OWN scene parser: src/parsescene.cpp, src/mesh.cpp (Mesh::processMesh, genTangent),
// src/imageio.cpp, src/texture.h compiled where they lie (staged outside the repo by oracle/build_parse_tool.sh; nothing of
// them is copied into this repository), run on a scene.json, with everything LoadScene produced dumped as raw structs.
//
// The one thing absent from this image is libassimp (only a Windows .lib ships with the reference).  Its role in the parser
// is Assimp::Importer::ReadFile -> aiScene; the stand-in below builds that aiScene from a sidecar "<mesh file>.aimesh"
// (int32 nv, nf, has_uv; float32 pos[nv][3], nrm[nv][3], uv[nv][2]; int32 idx[nf][3]) that the test writes with the
// package's own OBJ / PLY reader — one vertex per face corner, as assimp delivers without JoinIdenticalVertices.  Everything
// downstream of ReadFile — processNode / processMesh (transform by trs, normals by the inverse transpose, tangents, Triangle
// records), materials, media, lights, camera, integrator — is the reference's code.
//
//   parse_tool scene.json out.bin
//   out.bin: int32 width, height; float epsilon; Camera (104 B); int32 integrator type, maxDepth bits;
//            int32 n; Primitive[n]   (scene.primitives, in parse order, before any BVH)
//            int32 n; Material[n];   int32 n; Medium[n];   int32 n; Area[n];   int32 n_textures; {int32 w, h; uchar4[w*h]}...
//            Infinite (72 B; isvalid says whether the scene has one); if valid: float3 texels[width*height]
//            for every heterogeneous medium, in order: float density[nx*ny*nz]
#define private public
#include "parsescene.h"
#undef private
#include <assimp/Importer.hpp>
#include <assimp/scene.h>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

// ---- the stand-in for libassimp ------------------------------------------------------------------------------------
aiScene::aiScene() : mFlags(0), mRootNode(NULL), mNumMeshes(0), mMeshes(NULL), mNumMaterials(0), mMaterials(NULL),
                     mNumAnimations(0), mAnimations(NULL), mNumTextures(0), mTextures(NULL), mNumLights(0), mLights(NULL),
                     mNumCameras(0), mCameras(NULL), mPrivate(NULL) {}
aiScene::~aiScene() {}
static std::string g_err;
Assimp::Importer::Importer() : pimpl(NULL) {}
Assimp::Importer::~Importer() {}
const char* Assimp::Importer::GetErrorString() const { return g_err.c_str(); }
const aiScene* Assimp::Importer::ReadFile(const char* file, unsigned int) {
    std::string side = std::string(file) + ".aimesh";
    FILE* f = fopen(side.c_str(), "rb");
    if (!f) { g_err = "no sidecar " + side; return NULL; }
    int nv = 0, nf = 0, has_uv = 0;
    fread(&nv, 4, 1, f); fread(&nf, 4, 1, f); fread(&has_uv, 4, 1, f);
    aiMesh* m = new aiMesh();
    m->mNumVertices = nv; m->mNumFaces = nf;
    m->mVertices = new aiVector3D[nv]; m->mNormals = new aiVector3D[nv];
    std::vector<float> buf((size_t)nv * 3);
    fread(buf.data(), 4, buf.size(), f); for (int i = 0; i < nv; ++i) m->mVertices[i] = aiVector3D(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]);
    fread(buf.data(), 4, buf.size(), f); for (int i = 0; i < nv; ++i) m->mNormals[i] = aiVector3D(buf[3 * i], buf[3 * i + 1], buf[3 * i + 2]);
    std::vector<float> uv((size_t)nv * 2);
    fread(uv.data(), 4, uv.size(), f);
    if (has_uv) { m->mTextureCoords[0] = new aiVector3D[nv]; for (int i = 0; i < nv; ++i) m->mTextureCoords[0][i] = aiVector3D(uv[2 * i], uv[2 * i + 1], 0.f); }
    m->mFaces = new aiFace[nf];
    std::vector<int> idx((size_t)nf * 3);
    fread(idx.data(), 4, idx.size(), f);
    for (int i = 0; i < nf; ++i) { m->mFaces[i].mNumIndices = 3; m->mFaces[i].mIndices = new unsigned int[3]; for (int k = 0; k < 3; ++k) m->mFaces[i].mIndices[k] = idx[3 * i + k]; }
    fclose(f);
    aiScene* s = new aiScene();                 // lives as long as the process: the parser only reads it
    s->mNumMeshes = 1; s->mMeshes = new aiMesh*[1]; s->mMeshes[0] = m;
    s->mRootNode = new aiNode(); s->mRootNode->mNumMeshes = 1; s->mRootNode->mMeshes = new unsigned int[1]; s->mRootNode->mMeshes[0] = 0;
    return s;
}

template <class T> static void put_vec(FILE* f, const std::vector<T>& v) { int n = (int)v.size(); fwrite(&n, 4, 1, f); if (n) fwrite(v.data(), sizeof(T), v.size(), f); }

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: parse_tool scene.json out.bin\n"); return 1; }
    GlobalConfig config; Scene scene;
    scene.infinite.isvalid = false;
    if (!LoadScene(argv[1], config, scene)) { fprintf(stderr, "LoadScene failed\n"); return 2; }
    FILE* f = fopen(argv[2], "wb");
    fwrite(&config.width, 4, 1, f); fwrite(&config.height, 4, 1, f); fwrite(&config.epsilon, 4, 1, f);
    fwrite(&config.camera, sizeof(Camera), 1, f);
    int type = (int)scene.integrator.type; fwrite(&type, 4, 1, f); fwrite(&scene.integrator.maxDepth, 4, 1, f);
    put_vec(f, scene.primitives); put_vec(f, scene.materials); put_vec(f, scene.mediums); put_vec(f, scene.lights);
    int nt = (int)scene.textures.size(); fwrite(&nt, 4, 1, f);
    for (int i = 0; i < nt; ++i) { fwrite(&scene.textures[i].width, 4, 1, f); fwrite(&scene.textures[i].height, 4, 1, f); fwrite(scene.textures[i].data.data(), 4, scene.textures[i].data.size(), f); }
    fwrite(&scene.infinite, sizeof(Infinite), 1, f);
    if (scene.infinite.isvalid) fwrite(scene.infinite.data, sizeof(float3), (size_t)scene.infinite.width * scene.infinite.height, f);
    for (size_t i = 0; i < scene.mediums.size(); ++i)
        if (scene.mediums[i].type == MT_HETEROGENEOUS) {
            const Heterogeneous& h = scene.mediums[i].heterogeneous;
            fwrite(h.density, sizeof(float), (size_t)h.nx * h.ny * h.nz, f);
        }
    fclose(f);
    return 0;
}
