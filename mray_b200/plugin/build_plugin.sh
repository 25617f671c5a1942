#!/usr/bin/env bash
# Builds mray_b200/lib/libTracerDLL_B200.so — the TracerI plugin object — against the reference's own
# headers (they cannot be copied into this repository, so this only runs where /root/reference exists;
# the built .so travels with the snapshot). Needs oracle/ref_build/build_ref.sh to have run once
# (flags file + the reference's Core/TransientPool objects the plugin shares with its host).
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REF=${MRAY_REFERENCE:-/root/reference}
W=$ROOT/oracle/_ref/work
if [ ! -d "$REF/Source" ] || [ ! -f "$W/cxxflags.txt" ]; then echo "reference not present; keeping prebuilt plugin"; exit 0; fi
mkdir -p "$ROOT/mray_b200/lib"
# the reference's Core + TransientPool objects as one shared library (MRay ships them as libCore / libTransientPool):
# the HOST's runtime the plugin is loaded into, so it sits next to the plugin; the test-side driver keeps its own copy
if [ "${1:-}" != "plugin-only" ]; then
g++ -shared -o "$ROOT/mray_b200/lib/libmray_refcore.so" "$W"/obj/*Source_Core_*.o "$W"/obj/*TransientPool*.o -lpthread -latomic -ldl
cp "$ROOT/mray_b200/lib/libmray_refcore.so" "$ROOT/oracle/_ref/libmray_refcore.so"
fi
g++ $(cat "$W/cxxflags.txt") -I"$ROOT/include" -c "$HERE/tracer_b200.cpp" -o "$W/tracer_b200.o"
g++ -shared -o "$ROOT/mray_b200/lib/libTracerDLL_B200.so" "$W/tracer_b200.o" -L"$ROOT/mray_b200/lib" -lmray_b200 \
    -lmray_refcore -Wl,-rpath,'$ORIGIN' -Wl,--no-undefined
if [ "${1:-}" = "plugin-only" ]; then echo "PLUGIN_OK"; exit 0; fi   # leave the test-side driver alone (e.g. while it is rendering goldens)
# the TracerI driver (test side) shares the same Core library
g++ $(cat "$W/cxxflags.txt") -c "$ROOT/oracle/ref_build/tracer_driver.cpp" -o "$W/tracer_driver.o"
g++ -shared -rdynamic -o "$ROOT/oracle/_ref/libtracer_driver.so" "$W/tracer_driver.o" -L"$ROOT/oracle/_ref" -lmray_refcore \
    -Wl,-rpath,'$ORIGIN' -lpthread -latomic -ldl
# a process to host the driver in (the reference's spectral renderer looks for SpectraLUT/ next to the executable)
g++ -O1 -std=c++17 "$ROOT/oracle/ref_build/ref_render_host.cpp" -o "$ROOT/oracle/_ref/ref_render_host" -ldl
echo "PLUGIN_OK"
