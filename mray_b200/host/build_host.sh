#!/usr/bin/env bash
# Builds the host-side callers of the plugin against the reference's own headers:
#   mray_b200/lib/libSceneLoaderB200.so  — SceneLoaderI for the JSON scene format (host/scene_loader.cpp)
#   mray_b200/lib/mray_b200_run          — the head-less run command (host/run_main.cpp)
#   mray_b200/lib/mray_b200_spectra_lut_gen — SpectraLUTGen with the optimisation on the device (host/spectra_lut_gen.cpp)
# Like the plugin this only runs where /root/reference exists (needs oracle/ref_build/build_ref.sh + build_plugin.sh to have
# run: compile flags and libmray_refcore.so); the built files travel with the snapshot.
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REF=${MRAY_REFERENCE:-/root/reference}
W=$ROOT/oracle/_ref/work
if [ ! -d "$REF/Source" ] || [ ! -f "$W/cxxflags.txt" ]; then echo "reference not present; keeping prebuilt host tools"; exit 0; fi
LIB=$ROOT/mray_b200/lib
g++ $(cat "$W/cxxflags.txt") -c "$HERE/scene_loader.cpp" -o "$W/scene_loader.o"
g++ -shared -o "$LIB/libSceneLoaderB200.so" "$W/scene_loader.o" -L"$LIB" -lmray_refcore -Wl,-rpath,'$ORIGIN' -Wl,--no-undefined
FLAGS=$(sed 's/-fPIC//' "$W/cxxflags.txt")
g++ $FLAGS "$HERE/run_main.cpp" -o "$LIB/mray_b200_run" -L"$LIB" -lmray_refcore -Wl,-rpath,'$ORIGIN' -lpthread -latomic -ldl -rdynamic
# the spectral LUT generator: the reference's SpectraLUTGen command line, optimisation on the device
g++ $FLAGS -I"$ROOT/include" "$HERE/spectra_lut_gen.cpp" -o "$LIB/mray_b200_spectra_lut_gen" -L"$LIB" -lmray_b200 -lmray_refcore -Wl,-rpath,'$ORIGIN' -lpthread -latomic -ldl
echo "HOST_OK"
