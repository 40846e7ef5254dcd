#!/usr/bin/env bash
# TEST INFRASTRUCTURE.  Compiles the reference's OWN sources (from /root/reference, where they lie) into
#   oracle/_ref/libref_cuda.so       the reference CUDA integrator for sm_100a + headless C-ABI driver (GPU oracle of record)
#   oracle/_ref/libref_host.so       the same kernel bodies compiled by g++ (no FMA contraction; pins oracle/pt_oracle.cpp)
#   oracle/_ref/libref_host_fast.so  same, -O3 -march=x86-64-v3, for the CPU-baseline timing only
#   oracle/_ref/libadapter.so        the product's reference-signature adapter compiled against the reference headers
#   oracle/_ref/libadapter_emu.so    the same adapter linked against the CPU emulation build (tests/emu), for the GPU-less suite
# Only binaries are written into the repo tree (oracle/_ref/ is git-ignored, but travels with gpurun).
# The reference's own build system (CMake + Windows libs) is NOT used; see DESIGN.md.
set -euo pipefail
HERE="$(cd "$(dirname "$0")" && pwd)"
REF="${1:-/root/reference}"
OUT="$HERE/_ref"
WORK="$(mktemp -d /tmp/b200pt_refbuild.XXXXXX)"
trap 'rm -rf "$WORK"' EXIT
[ -d "$REF/src" ] || { echo "no reference at $REF; keeping prebuilt oracle/_ref" >&2; exit 0; }
mkdir -p "$OUT"
"$HERE/stage_ref.sh" "$REF" "$WORK" >/dev/null
INC="-I$REF/include -I$HERE/../include -I$HERE/refbuild"
NVCC="${NVCC:-/usr/local/cuda/bin/nvcc}"
ARCH="-gencode arch=compute_100a,code=sm_100a"

# ---- CUDA build: unmodified arithmetic, default FMA contraction (what a user of the reference gets) ----
$NVCC $ARCH -O3 -w -Xcompiler -fPIC -I"$WORK/src" $INC -c "$WORK/src/pathtracer.cu" -o "$WORK/pt.o"
$NVCC $ARCH -O3 -w -Xcompiler -fPIC -I"$WORK/src" $INC -c "$HERE/refbuild/ref_cuda_harness.cu" -o "$WORK/harness.o"
g++ -O2 -w -fPIC -I"$WORK/src" $INC -I/usr/local/cuda/include -c "$WORK/src/bvh.cpp" -o "$WORK/bvh.o"
$NVCC -shared $ARCH "$WORK/pt.o" "$WORK/harness.o" "$WORK/bvh.o" -o "$OUT/libref_cuda.so"

# ---- host build of the same kernel bodies ----
mkdir -p "$WORK/hostsrc"; cp "$WORK"/src/* "$WORK/hostsrc/"
python3 - "$WORK/hostsrc" <<'PY'
import sys, re
d = sys.argv[1]
s = open(d + '/pathtracer.cu', encoding='latin-1').read()
s = s[:s.index('void BeginRender(')]                      # cut CUDA-runtime plumbing (BeginRender..Render)
s = re.sub(r'(?m)^(\s*)SPPMSetParam << <1, 1 >> >\(.*$', r'\1/* launch removed in host build */', s)
# nvcc device code evaluates multi-draw argument lists left to right; g++ goes right to left.
s = s.replace('make_float3(uniform(rng), uniform(rng), uniform(rng))', 'seq3_(uniform, rng)')
s = s.replace('make_float2(uniform(rng), uniform(rng))', 'seq2_(uniform, rng)')
s = s.replace('make_float4(uniform(rng), uniform(rng), uniform(rng), uniform(rng))', 'seq4_(uniform, rng)')
s = s.replace('UniformDisk(uniform(rng), uniform(rng), unuse)', 'UniformDisk(seq2_(uniform, rng), unuse)')
s = s.replace('CosineHemiSphere(uniform(rng), uniform(rng), nor, pdf)', 'CosineHemiSphere(seq2_(uniform, rng), nor, pdf)')
helpers = '''
template<class U, class R> static inline float2 seq2_(U& uniform, R& rng){ float a = uniform(rng); float b = uniform(rng); return make_float2(a, b); }
template<class U, class R> static inline float3 seq3_(U& uniform, R& rng){ float a = uniform(rng); float b = uniform(rng); float c = uniform(rng); return make_float3(a, b, c); }
template<class U, class R> static inline float4 seq4_(U& uniform, R& rng){ float a = uniform(rng); float b = uniform(rng); float c = uniform(rng); float d = uniform(rng); return make_float4(a, b, c, d); }
static inline float2 UniformDisk(float2 u, float& pdf){ return UniformDisk(u.x, u.y, pdf); }
static inline float3 CosineHemiSphere(float2 u, float3& n, float& pdf){ return CosineHemiSphere(u.x, u.y, n, pdf); }
'''
i = s.index('Camera* dev_camera;')
open(d + '/pathtracer_host.cu', 'w', encoding='latin-1').write(s[:i] + helpers + s[i:])
# the reference's host fallbacks for fminf/fmaxf clash with libm under g++; libm's have CUDA's NaN semantics
c = open(d + '/cutil_math.h', encoding='latin-1').read()
i0 = c.index('#ifndef __CUDACC__'); i1 = c.index('#endif', i0) + len('#endif')
open(d + '/cutil_math.h', 'w', encoding='latin-1').write(c[:i0] + '#include <math.h>' + c[i1:])
PY
HOSTFLAGS="-fopenmp -fPIC -w -x c++ -include $HERE/refbuild/host_shim.h -I$WORK/hostsrc $INC -I/usr/local/cuda/include"
g++ -O2 -ffp-contract=off $HOSTFLAGS -c "$HERE/refbuild/ref_host_harness.cpp" -o "$WORK/hh.o"
g++ -shared -fopenmp "$WORK/hh.o" "$WORK/bvh.o" -o "$OUT/libref_host.so" -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread
g++ -O3 -march=x86-64-v3 $HOSTFLAGS -c "$HERE/refbuild/ref_host_harness.cpp" -o "$WORK/hhf.o"
g++ -shared -fopenmp "$WORK/hhf.o" "$WORK/bvh.o" -o "$OUT/libref_host_fast.so" -L/usr/local/cuda/lib64 -lcudart_static -ldl -lrt -lpthread

# ---- drop-in adapter check: the product's BeginRender/Render/EndRender (gpu-pathtracer_b200/host/pathtracer_adapter.cpp)
# compiled against the reference's own headers + a headless driver; links the product library libb200pt.so ----
HOSTDIR="$HERE/../gpu-pathtracer_b200/host"; CSRC="$HERE/../gpu-pathtracer_b200/csrc"
if [ -f "$CSRC/libb200pt.so" ]; then
  g++ -O2 -w -fPIC -std=c++17 -I"$WORK/src" $INC -I/usr/local/cuda/include -c "$HOSTDIR/pathtracer_adapter.cpp" -o "$WORK/adapter.o"
  g++ -O2 -w -fPIC -std=c++17 -I"$WORK/src" $INC -I/usr/local/cuda/include -c "$HOSTDIR/adapter_harness.cpp" -o "$WORK/adapter_h.o"
  g++ -shared "$WORK/adapter.o" "$WORK/adapter_h.o" "$WORK/bvh.o" -o "$OUT/libadapter.so" -L"$CSRC" -lb200pt \
      -Wl,-rpath,'$ORIGIN/../../gpu-pathtracer_b200/csrc' -L/usr/local/cuda/lib64 -lcudart
  # the same adapter against the CPU emulation build of the product (drop-in boundary on GPU-less machines);
  # -Bsymbolic: the harness's cudaMalloc & co. bind to the host shim in this .so even when a real libcudart is loaded
  EMU="$HERE/../tests/emu"
  if [ -f "$EMU/libb200pt_emu.so" ]; then
    g++ -O2 -w -fPIC -std=c++17 -I/usr/local/cuda/include -c "$HERE/refbuild/cudart_host_shim.cpp" -o "$WORK/cudart_shim.o"
    g++ -shared "$WORK/adapter.o" "$WORK/adapter_h.o" "$WORK/bvh.o" "$WORK/cudart_shim.o" -o "$OUT/libadapter_emu.so" -L"$EMU" -lb200pt_emu \
        -Wl,-Bsymbolic -Wl,-rpath,'$ORIGIN/../../tests/emu'
  fi
fi
ls -la "$OUT"
