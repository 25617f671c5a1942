#!/usr/bin/env bash
# Builds the UNMODIFIED reference (yalcinerbora/mray) CPU-device tracer from the sources where
# they lie under /root/reference into oracle/_ref/ (git-ignored). Test/baseline infrastructure
# only: nothing in the product path links or loads these files.
#   usage: oracle/ref_build/build_ref.sh [release|parity]
#     parity  : -O2, IEEE fp (no unsafe math)          -> libTracerDLL_CPU.so        (checker)
#     release : reference Release flags (-O3 -funsafe-math-optimizations -fno-math-errno,
#               CMake/Include/CompilerOptions.cmake:L117-121) -> libTracerDLL_CPU_rel.so (timing)
set -euo pipefail
MODE=${1:-parity}
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REF=${MRAY_REFERENCE:-/root/reference}
R=$REF/Source
OUT=$ROOT/oracle/_ref
SITE=$(python -c 'import sysconfig; print(sysconfig.get_paths()["purelib"])')
if [ ! -d "$R" ]; then echo "reference not present at $REF; keeping prebuilt oracle/_ref" ; exit 0; fi
case $MODE in
  parity)  OPT="-O2"; SUF="" ;;
  release) OPT="-O3 -funsafe-math-optimizations -fno-math-errno"; SUF="_rel" ;;
  safe)    OPT="-O1 -fno-strict-aliasing -fwrapv -fno-delete-null-pointer-checks"; SUF="_safe" ;;
  *) echo "bad mode"; exit 1 ;;
esac
W=$OUT/work$SUF
mkdir -p "$W/obj" "$W/gen" "$W/inc" "$W/fmtinc"
cp "$HERE/gcc13_shim.h" "$HERE/mray_cmake.h" "$W/inc/"
ln -sfn "$SITE/torch/include/fmt" "$W/fmtinc/fmt"
ln -sfn "$SITE/flashinfer/data/spdlog/include/spdlog" "$W/fmtinc/spdlog"

DEFS="-DMRAY_GCC -DMRAY_LINUX -DNDEBUG -DFMT_HEADER_ONLY -DSPDLOG_FMT_EXTERNAL -DMRAY_GPU_BACKEND_CPU \
 -DMRAY_TRACER_DEVICE_SHARED_EXPORT -DMRAY_CORE_SHARED_EXPORT -DMRAY_TRANSIENT_POOL_SHARED_EXPORT"
INCS="-I$R -I$W/inc -I$W/fmtinc -I$W/gen -I$R/TracerDLL -I$R/Tracer"
CXXF="-std=c++23 $OPT -fPIC -x c++ -include $W/inc/gcc13_shim.h $DEFS $INCS -fpermissive -fno-access-control -w"
echo "$CXXF" > "$W/cxxflags.txt"

# 1. kernel/type generator (reference's own build-time tool)
if [ ! -x "$W/kgen" ]; then
  g++ -std=c++23 -O1 -DMRAY_GCC -DMRAY_LINUX -DNDEBUG -DFMT_HEADER_ONLY -I$R -I$W/fmtinc -I$W/inc \
      $R/TracerKernelGen/*.cpp $R/Core/Error.cpp $R/Core/Log.cpp -o "$W/kgen"
fi
grep -v -E '^R +(TexViewRenderer|HashGridRenderer|GuidedPTRendererSpectral)' \
     "$R/TracerDLL/TracerTypeGenInput.txt" > "$W/TypeGenInput.txt"
NGEN=8
if [ ! -f "$W/gen/_GEN_CommonKernels.cu" ]; then
  # hwAccel=1 keeps the generator from bailing out on the missing "CPU" HW tag; the Embree (HW_CPU)
  # instantiation lines it emits are build artefacts we drop (MRAY_ENABLE_HW_ACCELERATION is off).
  (cd "$W" && ./kgen "$W/TypeGenInput.txt" $NGEN 1 "$W/gen" MRAY_GPU_BACKEND_CPU CPU)
  sed -i '/Embree/d' "$W"/gen/_GEN_*.cu
fi
ls "$W/gen"

SRCS=()
for f in System ThreadPool SharedLibrary ColorFunctions MemAlloc Log Error Timer MRayDataType; do SRCS+=("$R/Core/$f.cpp"); done
SRCS+=("$R/TransientPool/TransientPool.cpp")
for f in DeviceMemoryCPU GPUSystemCPU TextureCPU; do SRCS+=("$R/Device/CPU/$f.cpp"); done
for f in TextureMemory.cpp StreamingTextureCache.cpp GenericTexture.cpp ColorConverter.cu TextureFilter.cu \
  PrimitiveDefaultTriangle.cu PrimitivesDefault.cu TransformsDefault.cu MaterialsDefault.cpp CamerasDefault.cpp \
  MediumsDefault.cpp LightsDefault.cu AcceleratorCommon.cu AcceleratorLinear.cu AcceleratorLBVH.cu RendererCommon.cu \
  PathTracerRendererBase.cu RenderImage.cpp TexViewRenderer.cu SurfaceRenderer.cu HashGridRenderer.cu Random.cu \
  Distributions.cu SobolMatrices.cpp SpectrumContext.cu RayPartitioner.cu HashGrid.cu MediaTracker.cu GenericGroup.cpp \
  TracerBase.cpp; do SRCS+=("$R/Tracer/$f"); done
for f in EntryPoint.cpp Tracer.cu PathTracerRenderer.cu; do SRCS+=("$R/TracerDLL/$f"); done
for f in "$W"/gen/_GEN_*.cu; do SRCS+=("$f"); done

compile_one() {
  src=$1; W=$2; shift 2
  base=$(echo "$src" | sed 's#[/.]#_#g')
  obj="$W/obj/$base.o"
  if [ -f "$obj" ] && [ "$obj" -nt "$src" ]; then exit 0; fi
  if g++ $(cat "$W/cxxflags.txt") -c "$src" -o "$obj" 2> "$obj.log"; then echo "OK   $src"; else echo "FAIL $src (see $obj.log)"; exit 1; fi
}
export -f compile_one
printf '%s\n' "${SRCS[@]}" | xargs -P "$(nproc)" -I{} bash -c 'compile_one "$@"' _ {} "$W"

g++ -shared -o "$OUT/libTracerDLL_CPU$SUF.so" "$W"/obj/*.o -lpthread -latomic -ldl
echo "LINK_OK $OUT/libTracerDLL_CPU$SUF.so"

# taps: a TU of ours that forwards C arrays to the reference's own kernels (see ref_taps.cpp)
g++ $(cat "$W/cxxflags.txt") -c "$HERE/ref_taps.cpp" -o "$W/ref_taps.o"
g++ -shared -o "$OUT/libref_taps$SUF.so" "$W/ref_taps.o" -L"$OUT" -lTracerDLL_CPU$SUF \
    -Wl,-rpath,'$ORIGIN' -lpthread -latomic -ldl
echo "LINK_OK $OUT/libref_taps$SUF.so"
