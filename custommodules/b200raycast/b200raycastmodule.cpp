// Registers the CUDA-backed raycasters (single- and multi-volume).  Two registrations each (SURVEY.md 8b "Running unchanged workspaces"):
//   - under its own class name "B200SingleVolumeRaycaster" (always);
//   - under the reference's type name "SingleVolumeRaycaster" so that existing .vws files deserialise onto it
//     (VoreenModule::registerSerializableType(typeName, obj), V/src/core/voreenmodule.cpp:169-176).  Type lookup is
//     first-match over modules in registration order (V/src/core/voreenapplication.cpp:1588-1595), so the Base module
//     must not register the same name first: patches/basemodule_guard.patch puts its two lines under
//     #ifndef VRN_MODULE_B200RAYCAST.  Without the patch, load workspaces through
//     NetworkConverter::changeProcessorType(elem, "SingleVolumeRaycaster", "B200SingleVolumeRaycaster").
#include "b200raycastmodule.h"
#include "processors/b200singlevolumeraycaster.h"
#include "processors/b200multivolumeraycaster.h"
#include "processors/b200simpleraycaster.h"
#include "processors/b200rgbraycaster.h"

namespace voreen {

B200RaycastModule::B200RaycastModule(const std::string& modulePath)
    : VoreenModule(modulePath)
{
    setID("B200Raycast");
    setGuiName("B200 Raycasting (CUDA sm_100a)");
    registerProcessor(new B200SingleVolumeRaycaster());
    registerSerializableType("SingleVolumeRaycaster", new B200SingleVolumeRaycaster());
    registerProcessor(new B200MultiVolumeRaycaster());
    registerSerializableType("MultiVolumeRaycaster", new B200MultiVolumeRaycaster());
    registerProcessor(new B200SimpleRaycaster());
    registerSerializableType("SimpleRaycaster", new B200SimpleRaycaster());
    registerProcessor(new B200RGBRaycaster());
    registerSerializableType("RGBRaycaster", new B200RGBRaycaster());
}

std::string B200RaycastModule::getDescription() const {
    return "Direct volume ray casting on NVIDIA B200 through libb200rc (no GLSL, no CPU fallback).";
}

} // namespace
