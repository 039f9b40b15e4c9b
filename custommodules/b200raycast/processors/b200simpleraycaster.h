// B200SimpleRaycaster -- the drop-in for voreen::SimpleRaycaster (V/modules/base/processors/render/simpleraycaster.cpp:36-50):
// VolumeRaycaster() base (same ports), properties "shader" (inert), "transferFunction", "camera".  process() replaces the
// rc_simple.frag draw call by b200rc_render with raycaster = B200RC_RAYCASTER_SIMPLE.
#ifndef VRN_B200SIMPLERAYCASTER_H
#define VRN_B200SIMPLERAYCASTER_H

#include "voreen/core/processors/volumeraycaster.h"
#include "voreen/core/properties/cameraproperty.h"
#include "voreen/core/properties/shaderproperty.h"
#include "voreen/core/properties/transfunc/1d/1dkeys/transfunc1dkeysproperty.h"

#include <vector>

#include "b200rc.h"

namespace voreen {

class B200SimpleRaycaster : public VolumeRaycaster {
public:
    B200SimpleRaycaster();
    virtual Processor* create() const            { return new B200SimpleRaycaster(); }
    virtual std::string getClassName() const     { return "B200SimpleRaycaster"; }
    virtual std::string getCategory() const      { return "Raycasting"; }
    virtual CodeState getCodeState() const       { return CODE_STATE_EXPERIMENTAL; }

protected:
    virtual void setDescriptions() {
        VolumeRaycaster::setDescriptions();
        setDescription("SimpleRaycaster semantics (unshaded DVR with a 1D transfer function) computed by hand-written CUDA for NVIDIA B200.");
    }
    virtual void initialize();
    virtual void deinitialize();
    virtual void beforeProcess();
    virtual void process();

private:
    void failOn(int status, const char* what);

    ShaderProperty shader_;                         ///< inert; keeps "shader" deserialisable
    TransFunc1DKeysProperty transferFunc_;
    CameraProperty camera_;

    b200rc_ctx* ctx_;
    const VolumeBase* residentVolume_;
    std::vector<uint16_t> hostColor_;
    std::vector<float> hostDepth_;

    static const std::string loggerCat_;
};

} // namespace

#endif
