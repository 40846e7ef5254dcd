// TEST INFRASTRUCTURE — a command-line front for the EXR code the reference links (its vendored src/tinyexr.h, compiled
// where it lies under /root/reference by oracle/build_exr_tool.sh; nothing of it is copied into this repository).
//   exr_tool load in.exr out.bin          LoadEXR exactly as ImageIO::LoadExr calls it (src/imageio.cpp:80-102):
//                                          out.bin = int32 width, int32 height, float32 RGBA[width*height]
//   exr_tool save out.exr w h comp half in.bin   three channels B, G, R from float32 RGB[w*h] through SaveEXRImageToFile,
//                                          the call sequence of ImageIO::SaveExr (src/imageio.cpp:104-160); comp = tinyexr
//                                          compression type (0 none, 1 rle, 2 zips, 3 zip, 4 piz), half = 1 stores HALF
// Used by oracle/make_exr_fixtures.py to pin gpu-pathtracer_b200/exr.py against the reference's reader and writer.
#define TINYEXR_IMPLEMENTATION
#include "tinyexr.h"
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>

int main(int argc, char** argv) {
    if (argc >= 4 && !strcmp(argv[1], "load")) {
        float* out = nullptr; int w = 0, h = 0; const char* err = nullptr;
        int ret = LoadEXR(&out, &w, &h, argv[2], &err);
        if (ret != TINYEXR_SUCCESS) { fprintf(stderr, "LoadEXR failed (%d): %s\n", ret, err ? err : "?"); return 2; }
        FILE* f = fopen(argv[3], "wb");
        fwrite(&w, 4, 1, f); fwrite(&h, 4, 1, f); fwrite(out, sizeof(float), (size_t)4 * w * h, f);
        fclose(f); free(out);
        return 0;
    }
    if (argc >= 8 && !strcmp(argv[1], "save")) {
        const int w = atoi(argv[3]), h = atoi(argv[4]), comp = atoi(argv[5]), half = atoi(argv[6]);
        std::vector<float> rgb((size_t)3 * w * h);
        FILE* f = fopen(argv[7], "rb");
        if (!f || fread(rgb.data(), sizeof(float), rgb.size(), f) != rgb.size()) { fprintf(stderr, "short input\n"); return 2; }
        fclose(f);
        EXRHeader header; InitEXRHeader(&header);
        EXRImage image; InitEXRImage(&image);
        image.num_channels = 3;
        std::vector<float> ch[3];
        for (int c = 0; c < 3; ++c) { ch[c].resize((size_t)w * h); for (size_t i = 0; i < (size_t)w * h; ++i) ch[c][i] = rgb[3 * i + c]; }
        float* ptr[3] = {ch[2].data(), ch[1].data(), ch[0].data()};             // B, G, R
        image.images = (unsigned char**)ptr; image.width = w; image.height = h;
        header.num_channels = 3;
        header.channels = (EXRChannelInfo*)malloc(sizeof(EXRChannelInfo) * 3);
        const char* names[3] = {"B", "G", "R"};
        for (int c = 0; c < 3; ++c) { strncpy(header.channels[c].name, names[c], 255); header.channels[c].name[strlen(names[c])] = '\0'; }
        header.pixel_types = (int*)malloc(sizeof(int) * 3);
        header.requested_pixel_types = (int*)malloc(sizeof(int) * 3);
        for (int c = 0; c < 3; ++c) { header.pixel_types[c] = TINYEXR_PIXELTYPE_FLOAT; header.requested_pixel_types[c] = half ? TINYEXR_PIXELTYPE_HALF : TINYEXR_PIXELTYPE_FLOAT; }
        header.compression_type = comp;
        const char* err = nullptr;
        int ret = SaveEXRImageToFile(&image, &header, argv[2], &err);
        if (ret != TINYEXR_SUCCESS) { fprintf(stderr, "SaveEXRImageToFile failed (%d): %s\n", ret, err ? err : "?"); return 2; }
        free(header.channels); free(header.pixel_types); free(header.requested_pixel_types);
        return 0;
    }
    fprintf(stderr, "usage: exr_tool load in.exr out.bin | exr_tool save out.exr w h comp half in.bin\n");
    return 1;
}
