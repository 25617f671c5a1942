#pragma once
// =========================== //
// GENERATED FILE DO NOT EDIT! //
// =========================== //

// Includes
#include "PathTracerRenderer.h"
#include "Tracer/SurfaceRenderer.h"
// Mandatory Headers
#include "Tracer/RenderWork.h"
#include "_GEN_RequestedTypes.h"

// ================= //
//     Renderers     //
// ================= //
template <class Renderer>
using EmptyRendererWorkTypes = RenderWorkTypePack
<
    Renderer, TypePack<>, TypePack<>, TypePack<>, TypePack<>
>;

template <class Renderer,
          template<class, class, class, class> class RenderWorkT,
          template<class, class, class> class RenderLightWorkT,
          template<class, class, class> class RenderCameraWorkT,
          template<class, class, class> class RenderMediumWorkT>
using RendererWorkTypes = RenderWorkTypePack
<
    Renderer,
    // RenderWork
    TypePack
    <
        RenderWorkT<Renderer, PrimGroupTriangle, MatGroupPassthrough, TransformGroupIdentity>,
        RenderWorkT<Renderer, PrimGroupTriangle, MatGroupPassthrough, TransformGroupSingle>,
        RenderWorkT<Renderer, PrimGroupTriangle, MatGroupLambert, TransformGroupIdentity>,
        RenderWorkT<Renderer, PrimGroupTriangle, MatGroupLambert, TransformGroupSingle>,
        RenderWorkT<Renderer, PrimGroupTriangle, MatGroupReflect, TransformGroupIdentity>,
        RenderWorkT<Renderer, PrimGroupTriangle, MatGroupReflect, TransformGroupSingle>,
        RenderWorkT<Renderer, PrimGroupTriangle, MatGroupRefract, TransformGroupIdentity>,
        RenderWorkT<Renderer, PrimGroupTriangle, MatGroupRefract, TransformGroupSingle>,
        RenderWorkT<Renderer, PrimGroupTriangle, MatGroupUnreal, TransformGroupIdentity>,
        RenderWorkT<Renderer, PrimGroupTriangle, MatGroupUnreal, TransformGroupSingle>,
        RenderWorkT<Renderer, PrimGroupSphere, MatGroupPassthrough, TransformGroupIdentity>,
        RenderWorkT<Renderer, PrimGroupSphere, MatGroupPassthrough, TransformGroupSingle>,
        RenderWorkT<Renderer, PrimGroupSphere, MatGroupLambert, TransformGroupIdentity>,
        RenderWorkT<Renderer, PrimGroupSphere, MatGroupLambert, TransformGroupSingle>,
        RenderWorkT<Renderer, PrimGroupSphere, MatGroupReflect, TransformGroupIdentity>,
        RenderWorkT<Renderer, PrimGroupSphere, MatGroupReflect, TransformGroupSingle>,
        RenderWorkT<Renderer, PrimGroupSphere, MatGroupRefract, TransformGroupIdentity>,
        RenderWorkT<Renderer, PrimGroupSphere, MatGroupRefract, TransformGroupSingle>,
        RenderWorkT<Renderer, PrimGroupSphere, MatGroupUnreal, TransformGroupIdentity>,
        RenderWorkT<Renderer, PrimGroupSphere, MatGroupUnreal, TransformGroupSingle>
    >,
    // Lights
    TypePack
    <
        RenderLightWorkT<Renderer, LightGroupNull, TransformGroupIdentity>,
        RenderLightWorkT<Renderer, LightGroupPrim<PrimGroupTriangle>, TransformGroupIdentity>,
        RenderLightWorkT<Renderer, LightGroupPrim<PrimGroupTriangle>, TransformGroupSingle>,
        RenderLightWorkT<Renderer, LightGroupSkysphere<CoOctaCoordConverter>, TransformGroupIdentity>,
        RenderLightWorkT<Renderer, LightGroupSkysphere<CoOctaCoordConverter>, TransformGroupSingle>,
        RenderLightWorkT<Renderer, LightGroupSkysphere<SphericalCoordConverter>, TransformGroupIdentity>,
        RenderLightWorkT<Renderer, LightGroupSkysphere<SphericalCoordConverter>, TransformGroupSingle>
    >,
    // Camera
    TypePack
    <
        RenderCameraWorkT<Renderer, CameraGroupPinhole, TransformGroupIdentity>,
        RenderCameraWorkT<Renderer, CameraGroupPinhole, TransformGroupSingle>
    >,
    // And finally, Media
    TypePack
    <
        RenderMediumWorkT<Renderer, MediumGroupVacuum, TransformGroupIdentity>,
        RenderMediumWorkT<Renderer, MediumGroupHomogeneous, TransformGroupIdentity>
    >
>;

using RendererTypeList = TypePack
<
    SurfaceRenderer,
    PathTracerRendererRGB,
    PathTracerRendererSpectral
>;

// Currently empty
using RendererWorkTypesList = TypePack
<
    RendererWorkTypes<SurfaceRenderer, SurfaceRenderWork, SurfaceRenderLightWork, SurfaceRenderCamWork, RenderMediumWork>,
    RendererWorkTypes<PathTracerRendererRGB, PathTracerRenderWork, PathTracerRenderLightWork, PathTracerRenderCamWork, RenderMediumWork>,
    RendererWorkTypes<PathTracerRendererSpectral, PathTracerRenderWork, PathTracerRenderLightWork, PathTracerRenderCamWork, RenderMediumWork>
>;
