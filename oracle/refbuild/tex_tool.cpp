// TEST INFRASTRUCTURE — the texels the reference would hold for an image file: the decoder it links (its vendored
// include/stb/stb_image.h, compiled where it lies by oracle/build_exr_tool.sh) followed by the conversion of
// ImageIO::LoadTexture(file, w, h, srgb = true) + Texture::Texture (src/imageio.cpp:11-58, src/texture.h:14-27) restated:
// vertical flip on load, x / 255 as x * (1.f / 255.f), powf(x, 2.2f) on r g b, (unsigned char)(v * 255) — under g++ / glibc.
//   tex_tool in.png out.bin      out.bin = int32 width, int32 height, int32 components, uchar4[width*height]
//   tex_tool in.jpg out.bin raw  out.bin = int32 width, int32 height, int32 components, the bytes stbi_load returned (flipped)
#define STB_IMAGE_IMPLEMENTATION
#include "stb/stb_image.h"
#include <cmath>
#include <cstdio>
#include <vector>

int main(int argc, char** argv) {
    if (argc < 3) { fprintf(stderr, "usage: tex_tool in.(png|jpg) out.bin\n"); return 1; }
    int w = 0, h = 0, comp = 0;
    stbi_set_flip_vertically_on_load(true);
    unsigned char* tex = stbi_load(argv[1], &w, &h, &comp, 0);
    if (!tex) { fprintf(stderr, "stbi_load failed\n"); return 2; }
    if (argc >= 4 && argv[3][0] == 'r') {
        FILE* f = fopen(argv[2], "wb");
        fwrite(&w, 4, 1, f); fwrite(&h, 4, 1, f); fwrite(&comp, 4, 1, f); fwrite(tex, 1, (size_t)w * h * comp, f);
        fclose(f);
        return 0;
    }
    std::vector<unsigned char> out((size_t)4 * w * h);
    const float inv = 1.f / 255.f;
    for (int i = 0; i < w * h; ++i) {
        float t[4] = {0.f, 0.f, 0.f, 1.f};
        if (comp == 1) { t[0] = t[1] = t[2] = tex[i] * inv; }
        else if (comp == 3) { for (int k = 0; k < 3; ++k) t[k] = tex[3 * i + k] * inv; }
        else if (comp == 4) { for (int k = 0; k < 4; ++k) t[k] = tex[4 * i + k] * inv; }
        else { fprintf(stderr, "component count %d is not handled by the reference\n", comp); return 3; }
        for (int k = 0; k < 3; ++k) t[k] = powf(t[k], 2.2f);
        for (int k = 0; k < 4; ++k) out[4 * i + k] = (unsigned char)(t[k] * 255);
    }
    FILE* f = fopen(argv[2], "wb");
    fwrite(&w, 4, 1, f); fwrite(&h, 4, 1, f); fwrite(&comp, 4, 1, f); fwrite(out.data(), 1, out.size(), f);
    fclose(f);
    stbi_image_free(tex);
    return 0;
}
