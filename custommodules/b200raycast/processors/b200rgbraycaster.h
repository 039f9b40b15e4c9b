// B200RGBRaycaster -- the drop-in for voreen::RGBRaycaster (V/modules/base/processors/render/rgbraycaster.cpp:38-50):
// VolumeRaycaster() base (same ports), properties "raycast.prg" (inert), "transferFunction", "applyColorModulation", "camera".  process() replaces the
// rc_rgb.frag draw call by b200rc_render with raycaster = B200RC_RAYCASTER_RGB.
#ifndef VRN_B200RGBRAYCASTER_H
#define VRN_B200RGBRAYCASTER_H

#include "voreen/core/processors/volumeraycaster.h"
#include "voreen/core/properties/cameraproperty.h"
#include "voreen/core/properties/boolproperty.h"
#include "voreen/core/properties/shaderproperty.h"
#include "voreen/core/properties/transfunc/1d/1dkeys/transfunc1dkeysproperty.h"

#include <vector>

#include "b200rc.h"

namespace voreen {

class B200RGBRaycaster : public VolumeRaycaster {
public:
    B200RGBRaycaster();
    virtual Processor* create() const            { return new B200RGBRaycaster(); }
    virtual std::string getClassName() const     { return "B200RGBRaycaster"; }
    virtual std::string getCategory() const      { return "Raycasting"; }
    virtual CodeState getCodeState() const       { return CODE_STATE_EXPERIMENTAL; }

protected:
    virtual void setDescriptions() {
        VolumeRaycaster::setDescriptions();
        setDescription("RGBRaycaster semantics (RGB volumes: colour from the data, opacity from a 1D transfer function over the hue) computed by hand-written CUDA for NVIDIA B200.");
    }
    virtual void initialize();
    virtual void deinitialize();
    virtual void beforeProcess();
    virtual void process();

private:
    void failOn(int status, const char* what);

    ShaderProperty shader_;                         ///< inert; keeps "raycast.prg" deserialisable
    TransFunc1DKeysProperty transferFunc_;
    BoolProperty applyColorModulation_;
    CameraProperty camera_;

    b200rc_ctx* ctx_;
    const VolumeBase* residentVolume_;
    std::vector<uint16_t> hostColor_;
    std::vector<float> hostDepth_;

    static const std::string loggerCat_;
};

} // namespace

#endif
