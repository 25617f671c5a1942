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
#include "PathTracerRenderer.h"

// Kernel Work Instantiations
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupPassthrough, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupPassthrough, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupPassthrough, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupPassthrough, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupLambert, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupLambert, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupLambert, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupLambert, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupReflect, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupReflect, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupReflect, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupReflect, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupRefract, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupRefract, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupRefract, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupRefract, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupUnreal, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupUnreal, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupUnreal, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupSphere, MatGroupUnreal, TransformGroupSingle, 1);

// Kernel Light Work Instantiations
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupSkysphere<CoOctaCoordConverter>, TransformGroupIdentity, 1);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupSkysphere<CoOctaCoordConverter>, TransformGroupSingle, 0);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupSkysphere<CoOctaCoordConverter>, TransformGroupSingle, 1);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupSkysphere<SphericalCoordConverter>, TransformGroupIdentity, 0);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupSkysphere<SphericalCoordConverter>, TransformGroupIdentity, 1);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupSkysphere<SphericalCoordConverter>, TransformGroupSingle, 0);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupSkysphere<SphericalCoordConverter>, TransformGroupSingle, 1);

// Kernel Camera Work Instantiations

// Kernel Media Work Instantiations
MRAY_RENDERER_MEDIUM_KERNEL_INSTANTIATE(PathTracerRendererRGB, MediumGroupHomogeneous, TransformGroupIdentity, 0);
MRAY_RENDERER_MEDIUM_KERNEL_INSTANTIATE(PathTracerRendererRGB, MediumGroupHomogeneous, TransformGroupIdentity, 1);
MRAY_RENDERER_MEDIUM_KERNEL_INSTANTIATE(PathTracerRendererRGB, MediumGroupHomogeneous, TransformGroupIdentity, 2);

