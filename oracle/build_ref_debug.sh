#!/usr/bin/env bash
# TEST INFRASTRUCTURE (parity diagnostics).  Builds oracle/_ref/libref_cuda_dbg.so: the reference's CUDA integrator
# with printf probes of NAMED per-bounce variables (ray, hit, light sample, BSDF samples, beta, Li) for ONE pixel,
# inserted into a scratch copy of the staged sources (never into the repo).  The probes only read variables the
# kernels already hold, so the arithmetic is untouched; scripts/parity_diag.py verifies that by comparing the
# probe build's image bit for bit with libref_cuda.so before trusting its output.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
OUT="$HERE/_ref"
WORK="$(mktemp -d /tmp/b200pt_refdbg.XXXXXX)"
trap 'rm -rf "$WORK"' EXIT
[ -d "$REF/src" ] || { echo "no reference at $REF" >&2; exit 0; }
mkdir -p "$OUT"
"$HERE/stage_ref.sh" "$REF" "$WORK" >/dev/null
python3 - "$WORK/src/pathtracer.cu" <<'PY'
import sys
p = sys.argv[1]
s = open(p, encoding='latin-1').read()
i0 = s.index('__global__ void Path(int iter, int maxDepth){'); i1 = s.index('__global__ void Volpath(int iter, int maxDepth){')
i2 = s.index('//**************************VolPath End')
head, path, vol, tail = s[:i0], s[i0:i1], s[i1:i2], s[i2:]
F3 = '%a %a %a'
def v3(n): return f'(double){n}.x, (double){n}.y, (double){n}.z'
def probe(tag, fmt, args): return f'\nif ((int)pixel == dbg_pixel) printf("{tag} {fmt}\\n", {args});\n'
H = probe('H', f'%d o {F3} d {F3} t %a pos {F3} nor {F3} uv %a %a dpdu {F3} beta {F3} Li {F3} prim %d %d', 'bounces, ' + ', '.join([v3('r.o'), v3('r.d'), '(double)r.tmax', v3('pos'), v3('nor'), '(double)uv.x, (double)uv.y', v3('dpdu'), v3('beta'), v3('Li'), 'isect.matIdx, isect.lightIdx']))
Lp = probe('L', f'lpdf %a cpdf %a sd {F3} tmax %a rad {F3} idx %d', '(double)lightPdf, (double)choicePdf, ' + v3('shadowRay.d') + ', (double)shadowRay.tmax, ' + v3('radiance') + ', idx')
M = probe('M', f'out {F3} fr {F3} pdf %a', v3('out') + ', ' + v3('fr') + ', (double)pdf')
Cc = probe('C', f'out {F3} fr {F3} pdf %a', v3('out') + ', ' + v3('fr') + ', (double)pdf')
D = probe('D', f'Li {F3} Ld {F3}', v3('Li') + ', ' + v3('Ld'))
E = probe('E', f'Li {F3}', v3('Li'))
V = probe('V', f'beta {F3} dist %a med %d', v3('beta') + ', (double)sampledDist, (int)sampledMedium')
def patch(body, vol):
    a = 'float3 dpdu = isect.dpdu;'
    assert body.count(a) == 1; body = body.replace(a, a + H)
    a = 'shadowRay.medium = r.medium;'
    assert body.count(a) == (2 if vol else 1); body = body.replace(a, a + Lp)
    a = 'SampleBSDF(material, -r.d, nor, uv, dpdu, us, out, fr, pdf);'
    assert body.count(a) == 1; body = body.replace(a, a + M)
    a = 'SampleBSDF(material, -r.d, nor, uv, dpdu, u, out, fr, pdf);'
    assert body.count(a) == 1; body = body.replace(a, a + Cc)
    a = 'Li += beta*Ld;'
    assert body.count(a) == 1; body = body.replace(a, a + D)
    a = 'if (!IsInf(Li) && !IsNan(Li))'
    assert body.count(a) == 1; body = body.replace(a, E + a)
    if vol:
        a = 'if (IsBlack(beta)) break;'
        assert body.count(a) == 1; body = body.replace(a, V + a)
    return body
decl = '__device__ int dbg_pixel = -1;\nextern "C" void refdbg_set(int p){ cudaMemcpyToSymbol(dbg_pixel, &p, sizeof(int)); cudaDeviceSynchronize(); }\n'
open(p, 'w', encoding='latin-1').write(head + decl + patch(path, False) + patch(vol, True) + tail)
PY
INC="-I$REF/include -I$HERE/../include -I$HERE/refbuild"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"
$NVCC $ARCH -O3 -w -Xcompiler -fPIC -I"$WORK/src" $INC -c "$WORK/src/pathtracer.cu" -o "$WORK/pt.o"
$NVCC $ARCH -O3 -w -Xcompiler -fPIC -DREFDBG -I"$WORK/src" $INC -c "$HERE/refbuild/ref_cuda_harness.cu" -o "$WORK/harness.o"
g++ -O2 -w -fPIC -I"$WORK/src" $INC -I/usr/local/cuda/include -c "$WORK/src/bvh.cpp" -o "$WORK/bvh.o"
$NVCC -shared $ARCH "$WORK/pt.o" "$WORK/harness.o" "$WORK/bvh.o" -o "$OUT/libref_cuda_dbg.so"
ls -la "$OUT/libref_cuda_dbg.so"
