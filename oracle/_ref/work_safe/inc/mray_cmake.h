// Stand-in for the header CMake would generate from Source/mray_cmake.h.in
// (oracle/_ref build only).
#pragma once
#include <string_view>
using namespace std::string_view_literals;
#define MRAY_VERSION_MAJOR 0
#define MRAY_VERSION_MINOR 1
#define MRAY_VERSION_PATCH 0
#define MRAY_BUILD_VISOR 0
#define MRAY_HOST_ARCH_BASIC
#define MRAY_PROJECT_NAME "MRay"sv
#define MRAY_PROJECT_DESCRIPTION "oracle build"sv
#define MRAY_PLATFORM_NAME "Linux"sv
#define MRAY_COMPILER_NAME "g++"sv
#define MRAY_GPU_PLATFORM_NAME "CPU"sv
#define MRAY_GPU_COMPILER_NAME "g++"sv
