#pragma once
// =========================== //
// GENERATED FILE DO NOT EDIT! //
// =========================== //

// Includes
#include "Tracer/AcceleratorLBVH.h"
#include "Tracer/AcceleratorLinear.h"
#include "Tracer/CamerasDefault.h"
#include "Tracer/LightsDefault.h"
#include "Tracer/MaterialsDefault.h"
#include "Tracer/MediumsDefault.h"
#include "Tracer/PrimitiveDefaultTriangle.h"
#include "Tracer/PrimitivesDefault.h"
#include "Tracer/TransformsDefault.h"
// Mandatory Headers
#include "Tracer/MetaLight.h"
//#include "Tracer/RenderWork.h"
#include "Tracer/AcceleratorWork.h"
// Guarded Headers
#if defined(MRAY_GPU_BACKEND_CPU) && defined(MRAY_ENABLE_HW_ACCELERATION)
    #include "Tracer/Embree/AcceleratorEmbree.h"
#endif

// ================= //
//     Primitives    //
// ================= //
using PrimGTypes = TypePack
<
    PrimGroupEmpty,
    PrimGroupTriangle,
    PrimGroupSphere
>;

// ================= //
//     Materials     //
// ================= //
using MatGTypes = TypePack
<
    MatGroupPassthrough,
    MatGroupLambert,
    MatGroupReflect,
    MatGroupRefract,
    MatGroupUnreal
>;

// ================= //
//     Transforms    //
// ================= //
using TransformGTypes = TypePack
<
    TransformGroupIdentity,
    TransformGroupSingle,
    TransformGroupMulti
>;

// ================= //
//      Cameras      //
// ================= //
using CamGTypes = TypePack
<
    CameraGroupPinhole
>;

// ================= //
//      Mediums      //
// ================= //
using MedGTypes = TypePack
<
    MediumGroupVacuum,
    MediumGroupHomogeneous
>;

// ================= //
//      Lights       //
// ================= //
using LightGTypes = TypePack
<
    LightGroupNull,
    LightGroupPrim<PrimGroupTriangle>,
    LightGroupSkysphere<CoOctaCoordConverter>,
    LightGroupSkysphere<SphericalCoordConverter>
>;

using MetaLightList = MetaLightArrayT
<
    TypePack<LightGroupNull, TransformGroupIdentity>,
    TypePack<LightGroupPrim<PrimGroupTriangle>, TransformGroupIdentity>,
    TypePack<LightGroupPrim<PrimGroupTriangle>, TransformGroupSingle>,
    TypePack<LightGroupSkysphere<CoOctaCoordConverter>, TransformGroupIdentity>,
    TypePack<LightGroupSkysphere<CoOctaCoordConverter>, TransformGroupSingle>,
    TypePack<LightGroupSkysphere<SphericalCoordConverter>, TransformGroupIdentity>,
    TypePack<LightGroupSkysphere<SphericalCoordConverter>, TransformGroupSingle>
>;

// ================= //
//    Accelerators   //
// ================= //
template <class Base, template<class> class Group>
using DefaultAccelTypePack = AccelTypePack
<
    Base,
    TypePack
    <
        Group<PrimGroupTriangle>,
        Group<PrimGroupSphere>
    >,
    TypePack
    <
        AcceleratorWork<Group<PrimGroupTriangle>, TransformGroupIdentity>,
        AcceleratorWork<Group<PrimGroupTriangle>, TransformGroupSingle>,
        AcceleratorWork<Group<PrimGroupSphere>, TransformGroupIdentity>,
        AcceleratorWork<Group<PrimGroupSphere>, TransformGroupSingle>
    >
>;

using DefaultLinearAccelTypePack = DefaultAccelTypePack<BaseAcceleratorLinear, AcceleratorGroupLinear>;
using DefaultBVHAccelTypePack = DefaultAccelTypePack<BaseAcceleratorLBVH, AcceleratorGroupLBVH>;
#if defined(MRAY_GPU_BACKEND_CPU) && defined(MRAY_ENABLE_HW_ACCELERATION)
    using DefaultDeviceAccelTypePack = DefaultAccelTypePack<BaseAcceleratorEmbree, AcceleratorGroupEmbree>;
#endif
