// =========================== //
// GENERATED FILE DO NOT EDIT! //
// =========================== //
#ifdef MRAY_WINDOWS
    // After nvcc passes through
    // some residual code caught by msvc
    // and "unreachable code" is generated
    // TODO: Investigate
    #pragma warning( disable : 4702)
#endif

// Definitions
#include "Tracer/RayGenKernels.h"
#include "Tracer/RenderWork.h"

// Implementations
#include "Tracer/RayGenKernels.kt.h"
#include "Tracer/RenderWork.kt.h"
#include "Tracer/TextureView.hpp"

// Types
#include "InstantiationMacros.h"

#include "RequestedTypes.h"
#include "Tracer/SurfaceRenderer.h"

// Kernel Work Instantiations
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupPassthrough, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupPassthrough, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupPassthrough, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupPassthrough, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupLambert, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupLambert, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupLambert, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupLambert, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupReflect, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupReflect, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupReflect, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupReflect, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupRefract, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupRefract, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupRefract, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupRefract, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupUnreal, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupUnreal, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupUnreal, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(SurfaceRenderer, PrimGroupSphere, MatGroupUnreal, TransformGroupSingle, 1);

// Kernel Light Work Instantiations
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(SurfaceRenderer, LightGroupSkysphere<CoOctaCoordConverter>, TransformGroupSingle, 0);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(SurfaceRenderer, LightGroupSkysphere<SphericalCoordConverter>, TransformGroupIdentity, 0);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(SurfaceRenderer, LightGroupSkysphere<SphericalCoordConverter>, TransformGroupSingle, 0);

// Kernel Camera Work Instantiations

// Kernel Media Work Instantiations

