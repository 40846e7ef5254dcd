// TEST INFRASTRUCTURE — a command-line front for the reference's OWN image I/O: src/imageio.cpp and src/texture.h compiled
// where they lie (staged with the mechanical include / cast patches of oracle/stage_ref.sh by oracle/build_imageio_tool.sh;
// nothing of them is copied into this repository).
//   imageio_tool texture in.(png|jpg) out.bin    Texture::Texture(file) (src/texture.h:14-27 -> ImageIO::LoadTexture, src/imageio.cpp:11-58):
//                                                out.bin = int32 width, int32 height, uchar4[width*height]
//   imageio_tool savepng in.bin w h out.png      ImageIO::SavePng (src/imageio.cpp:61-77) of float32 RGB[w*h]
//   imageio_tool loadexr in.exr out.bin          ImageIO::LoadExr: out.bin = int32 width, int32 height, float32 RGB[width*height]
//   imageio_tool saveexr in.bin w h out.exr      ImageIO::SaveExr of float32 RGB[w*h]
// Used by oracle/make_tex_fixtures.py / tests/test_frontend_io.py to pin gpu-pathtracer_b200/imageio.py, textures.py, jpeg.py, exr.py.
#include "imageio.h"
#include "texture.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>

static std::vector<float3> read_rgb(const char* path, int w, int h) {
    std::vector<float3> v((size_t)w * h);
    FILE* f = fopen(path, "rb");
    if (!f || fread(v.data(), sizeof(float3), v.size(), f) != v.size()) { fprintf(stderr, "short input\n"); exit(2); }
    fclose(f);
    return v;
}

int main(int argc, char** argv) {
    if (argc >= 4 && !strcmp(argv[1], "texture")) {
        Texture t(argv[2]);
        FILE* f = fopen(argv[3], "wb");
        fwrite(&t.width, 4, 1, f); fwrite(&t.height, 4, 1, f); fwrite(t.data.data(), sizeof(uchar4), t.data.size(), f);
        fclose(f);
        return 0;
    }
    if (argc >= 6 && !strcmp(argv[1], "savepng")) {
        int w = atoi(argv[3]), h = atoi(argv[4]);
        std::vector<float3> v = read_rgb(argv[2], w, h);
        return ImageIO::SavePng(argv[5], w, h, v.data()) ? 0 : 2;
    }
    if (argc >= 4 && !strcmp(argv[1], "loadexr")) {
        int w = 0, h = 0; std::vector<float3> v;
        if (!ImageIO::LoadExr(argv[2], w, h, v)) return 2;
        FILE* f = fopen(argv[3], "wb");
        fwrite(&w, 4, 1, f); fwrite(&h, 4, 1, f); fwrite(v.data(), sizeof(float3), v.size(), f);
        fclose(f);
        return 0;
    }
    if (argc >= 6 && !strcmp(argv[1], "saveexr")) {
        int w = atoi(argv[3]), h = atoi(argv[4]);
        std::vector<float3> v = read_rgb(argv[2], w, h);
        return ImageIO::SaveExr(argv[5], w, h, v) ? 0 : 2;
    }
    fprintf(stderr, "usage: imageio_tool texture|savepng|loadexr|saveexr ...\n");
    return 1;
}
