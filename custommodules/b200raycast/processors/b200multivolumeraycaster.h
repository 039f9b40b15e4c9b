// B200MultiVolumeRaycaster -- the drop-in for voreen::MultiVolumeRaycaster.
//
// Same base class and the same additional ports ("volume2..4", "entryPort2..4", "exitPort2..4", "entryPortGlobal",
// "exitPortGlobal") and property IDs ("classification2..4", "shading1..4", "transferFunction1..4", "camera", "raycast.prg")
// as V/modules/base/processors/render/multivolumeraycaster.cpp:38-110, so a workspace written for the GLSL processor
// deserialises onto this one unchanged.  process() replaces the rc_multivolume.frag draw call by b200rc_render with
// multi_volume = 1: world-space rays from the global entry/exit images, one (entry, exit) window per connected volume.
#ifndef VRN_B200MULTIVOLUMERAYCASTER_H
#define VRN_B200MULTIVOLUMERAYCASTER_H

#include "voreen/core/processors/volumeraycaster.h"
#include "voreen/core/properties/cameraproperty.h"
#include "voreen/core/properties/shaderproperty.h"
#include "voreen/core/properties/transfunc/1d/1dkeys/transfunc1dkeysproperty.h"

#include <vector>

#include "b200rc.h"

namespace voreen {

class B200MultiVolumeRaycaster : public VolumeRaycaster {
public:
    B200MultiVolumeRaycaster();
    virtual Processor* create() const            { return new B200MultiVolumeRaycaster(); }
    virtual std::string getClassName() const     { return "B200MultiVolumeRaycaster"; }
    virtual std::string getCategory() const      { return "Raycasting"; }
    virtual CodeState getCodeState() const       { return CODE_STATE_EXPERIMENTAL; }
    virtual bool isReady() const;

protected:
    virtual void setDescriptions() {
        VolumeRaycaster::setDescriptions();
        setDescription("MultiVolumeRaycaster semantics (up to four volumes with their own transformation, transfer function, "
                       "classification and shading mode, marched along common world-space rays) computed by hand-written CUDA "
                       "for NVIDIA B200 instead of a GLSL fragment shader.");
    }
    virtual void initialize();
    virtual void deinitialize();
    virtual void beforeProcess();
    virtual void process();

private:
    static const int kVolumes = 4;
    void failOn(int status, const char* what);
    void uploadVolume(int slot, const VolumeBase* vol);

    VolumePort volumeInport2_; RenderPort entryPort2_; RenderPort exitPort2_;
    VolumePort volumeInport3_; RenderPort entryPort3_; RenderPort exitPort3_;
    VolumePort volumeInport4_; RenderPort entryPort4_; RenderPort exitPort4_;
    RenderPort entryPortGlobal_; RenderPort exitPortGlobal_;

    ShaderProperty shaderProp_;                     ///< inert; keeps "raycast.prg" deserialisable
    StringOptionProperty classificationMode2_, classificationMode3_, classificationMode4_;
    StringOptionProperty shadeMode1_, shadeMode2_, shadeMode3_, shadeMode4_;
    TransFunc1DKeysProperty transferFunc1_, transferFunc2_, transferFunc3_, transferFunc4_;
    CameraProperty camera_;

    b200rc_ctx* ctx_;
    const VolumeBase* resident_[kVolumes];          ///< what each device slot currently holds
    float mvOpacityCorrectionFactor_;               ///< the uniform's life cycle, see process()
    std::vector<float> entry_, exit_;
    std::vector<uint16_t> hostColor_[B200RC_MAX_OUTPUTS];
    std::vector<float> hostDepth_;

    static const std::string loggerCat_;
};

} // namespace

#endif
