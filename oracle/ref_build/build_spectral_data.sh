#!/usr/bin/env bash
# Builds the reference's own SpectraLUTGen tool and the spectrum tap from /root/reference, runs the tool
# to produce SpectraLUT/ACES_CG.mrspectra (an INPUT of the hot path), and installs
# it where the product loads it (mray_b200/data/, untracked because of its size). Needs build_ref.sh and
# the plugin script to have run (compile flags + libmray_refcore / libTracerDLL_CPU).
set -euo pipefail
HERE=$(cd "$(dirname "$0")" && pwd)
ROOT=$(cd "$HERE/../.." && pwd)
REF=${MRAY_REFERENCE:-/root/reference}
OUT=$ROOT/oracle/_ref
W=$OUT/work
if [ ! -d "$REF/Source" ] || [ ! -f "$W/cxxflags.txt" ]; then echo "reference not present; keeping prebuilt spectral data"; exit 0; fi
FLAGS=$(sed 's/-fPIC//' "$W/cxxflags.txt")
mkdir -p "$OUT/SpectraLUT" "$ROOT/mray_b200/data"
if [ ! -x "$OUT/spectra_lut_gen" ]; then
  g++ $FLAGS -O2 "$REF/Source/SpectraLUTGen/main.cpp" -o "$OUT/spectra_lut_gen" -L"$OUT" -lmray_refcore -Wl,-rpath,'$ORIGIN' -lpthread
fi
if [ ! -f "$OUT/SpectraLUT/ACES_CG.mrspectra" ]; then
  "$OUT/spectra_lut_gen" 64 ACES_CG "$OUT/SpectraLUT"
fi
cp -u "$OUT/SpectraLUT/ACES_CG.mrspectra" "$ROOT/mray_b200/data/ACES_CG.mrspectra"
g++ $FLAGS "$HERE/ref_spectrum_tap.cpp" -o "$OUT/ref_spectrum_tap" -L"$OUT" -lTracerDLL_CPU -Wl,-rpath,'$ORIGIN' -lpthread -latomic -ldl
# sampler tap (RNGGroupSobol / RNGGroupZSobol on the reference's CPU backend; golden vectors + generator matrices)
g++ $FLAGS "$HERE/ref_rng_tap.cpp" -o "$OUT/ref_rng_tap" -L"$OUT" -lTracerDLL_CPU -Wl,-rpath,'$ORIGIN' -lpthread -latomic -ldl
echo "SPECTRAL_DATA_OK"
