// TEST INFRASTRUCTURE — the transform arithmetic of the reference's scene parser, run through the GLM it vendors
// (REF/include/glm, compiled where it lies by oracle/build_exr_tool.sh; nothing of it is copied into this repository).
// Reads records of 32-bit words (hex) from stdin, one per line, writes the results as hex words:
//   T sx sy sz tx ty tz rx ry rz vx vy vz nx ny nz   ->  trs[16] invT[16] v'[3] n'[3]
//       trs = t * r * s exactly as src/parsescene.cpp:349-355 builds it, then Mesh::processMesh's per-vertex transform
//       (src/mesh.cpp:50-62): v' = vec3(trs * vec4(v, 1)), n' = normalize(vec3(transpose(inverse(trs)) * vec4(n, 0)))
//   R rx ry rz                                         ->  uu[3] vv[3] ww[3]   ("rotate" frame of an infinite light, src/parsescene.cpp:551-560)
//   M m0 .. m15                                        ->  uu[3] vv[3] ww[3]   ("matrix" frame, src/parsescene.cpp:563-568)
// Used by oracle/make_glm_fixtures.py to pin gpu-pathtracer_b200/scenes.py (_trs, _transform_mesh, infinite frames).
#include <glm/glm.hpp>
#include <glm/gtc/matrix_transform.hpp>
#include <cstdio>
#include <cstring>
#include <cstdint>
using namespace glm;

static float rd(const char*& p) { unsigned u; int n; sscanf(p, "%x%n", &u, &n); p += n; float f; memcpy(&f, &u, 4); return f; }
static void wr(float f) { unsigned u; memcpy(&u, &f, 4); printf("%08x ", u); }
static void wr3(const vec3& v) { wr(v.x); wr(v.y); wr(v.z); }
static void wrm(const mat4& m) { const float* f = &m[0][0]; for (int i = 0; i < 16; ++i) wr(f[i]); }

int main() {
    char line[4096];
    while (fgets(line, sizeof line, stdin)) {
        const char* p = line + 1;
        if (line[0] == 'T') {
            float a[15]; for (int i = 0; i < 15; ++i) a[i] = rd(p);
            mat4 trs, t, r, s;
            s = glm::scale(s, vec3(a[0], a[1], a[2]));
            t = glm::translate(t, vec3(a[3], a[4], a[5]));
            r = glm::rotate(r, radians(a[6]), vec3(1, 0, 0));
            r = glm::rotate(r, radians(a[7]), vec3(0, 1, 0));
            r = glm::rotate(r, radians(a[8]), vec3(0, 0, 1));
            trs = t*r*s;
            mat4 invT = transpose(inverse(trs));
            vec3 v(a[9], a[10], a[11]), n(a[12], a[13], a[14]);
            v = vec3(trs*vec4(v, 1));
            n = normalize(vec3(invT*vec4(n, 0)));
            wrm(trs); wrm(invT); wr3(v); wr3(n);
        } else if (line[0] == 'R') {
            float x = rd(p), y = rd(p), z = rd(p);
            mat4 rs;
            rs = rotate(rs, radians(x), vec3(1, 0, 0));
            rs = rotate(rs, radians(y), vec3(0, 1, 0));
            rs = rotate(rs, radians(z), vec3(0, 0, 1));
            wr3(vec3(rs * vec4(1, 0, 0, 0))); wr3(vec3(rs * vec4(0, 1, 0, 0))); wr3(vec3(rs * vec4(0, 0, 1, 0)));
        } else if (line[0] == 'M') {
            float x[16]; for (int i = 0; i < 16; ++i) x[i] = rd(p);
            mat4 rs; memcpy(&rs[0], x, 16 * sizeof(float));
            rs = inverse(rs);
            wr3(vec3(rs * vec4(1, 0, 0, 0))); wr3(vec3(rs * vec4(0, 1, 0, 0))); wr3(vec3(rs * vec4(0, 0, 1, 0)));
        } else continue;
        printf("\n");
    }
    return 0;
}
