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
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupPassthrough, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupPassthrough, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupPassthrough, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupPassthrough, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupLambert, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupLambert, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupLambert, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupLambert, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupReflect, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupReflect, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupReflect, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupReflect, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupRefract, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupRefract, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupRefract, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupRefract, TransformGroupSingle, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupUnreal, TransformGroupIdentity, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupUnreal, TransformGroupIdentity, 1);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupUnreal, TransformGroupSingle, 0);
MRAY_RENDERER_KERNEL_INSTANTIATE(PathTracerRendererRGB, PrimGroupTriangle, MatGroupUnreal, TransformGroupSingle, 1);

// Kernel Light Work Instantiations
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupNull, TransformGroupIdentity, 0);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupNull, TransformGroupIdentity, 1);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupPrim<PrimGroupTriangle>, TransformGroupIdentity, 0);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupPrim<PrimGroupTriangle>, TransformGroupIdentity, 1);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupPrim<PrimGroupTriangle>, TransformGroupSingle, 0);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupPrim<PrimGroupTriangle>, TransformGroupSingle, 1);
MRAY_RENDERER_LIGHT_KERNEL_INSTANTIATE(PathTracerRendererRGB, LightGroupSkysphere<CoOctaCoordConverter>, TransformGroupIdentity, 0);

// Kernel Camera Work Instantiations

// Kernel Media Work Instantiations
MRAY_RENDERER_MEDIUM_KERNEL_INSTANTIATE(PathTracerRendererRGB, MediumGroupVacuum, TransformGroupIdentity, 0);
MRAY_RENDERER_MEDIUM_KERNEL_INSTANTIATE(PathTracerRendererRGB, MediumGroupVacuum, TransformGroupIdentity, 1);
MRAY_RENDERER_MEDIUM_KERNEL_INSTANTIATE(PathTracerRendererRGB, MediumGroupVacuum, TransformGroupIdentity, 2);

