// B200SingleVolumeRaycaster -- the drop-in for voreen::SingleVolumeRaycaster.
//
// Same base class (VolumeRaycaster => identical ports "volumehandle.volumehandle", "image.entrypoints",
// "image.exitpoints", "image.output[,1,2]" and identical quality / compositing / shading / gradient properties,
// V/src/core/processors/volumeraycaster.cpp:44-170) and the same additional property IDs as
// V/modules/base/processors/render/singlevolumeraycaster.cpp:46-67, so a workspace written for the GLSL processor
// deserialises onto this one unchanged.  process() replaces the GLSL draw call by calls into the C ABI of
// include/b200rc.h; "raycast.prg" is kept as an inert ShaderProperty so that its serialised value still loads.
#ifndef VRN_B200SINGLEVOLUMERAYCASTER_H
#define VRN_B200SINGLEVOLUMERAYCASTER_H

#include "voreen/core/processors/volumeraycaster.h"
#include "voreen/core/properties/cameraproperty.h"
#include "voreen/core/properties/shaderproperty.h"
#include "voreen/core/properties/vectorproperty.h"
#include "voreen/core/properties/transfunc/transfunctypeproperty.h"
#include "voreen/core/properties/transfunc/1d/1dkeys/transfunc1dkeysproperty.h"
#include "voreen/core/properties/transfunc/1d/1dgaussian/transfunc1dgaussianproperty.h"
#include "voreen/core/properties/transfunc/2d/2dprimitives/transfunc2dprimitivesproperty.h"

#include <memory>
#include <vector>

#include "b200rc.h"

namespace voreen {

class B200SingleVolumeRaycaster : public VolumeRaycaster {
public:
    B200SingleVolumeRaycaster();
    virtual ~B200SingleVolumeRaycaster();
    virtual Processor* create() const            { return new B200SingleVolumeRaycaster(); }
    virtual std::string getClassName() const     { return "B200SingleVolumeRaycaster"; }
    virtual std::string getCategory() const      { return "Raycasting"; }
    virtual CodeState getCodeState() const       { return CODE_STATE_EXPERIMENTAL; }
    virtual bool isReady() const;

protected:
    virtual void setDescriptions() {
        VolumeRaycaster::setDescriptions();
        setDescription("SingleVolumeRaycaster semantics (DVR/MIP/MOP/MIDA/ISO/FHP/FHN, 1D/2D/pre-integrated "
                       "classification, on-the-fly gradients, Phong and friends) computed by hand-written CUDA for "
                       "NVIDIA B200 instead of a GLSL fragment shader.");
    }
    virtual void initialize();
    virtual void deinitialize();
    virtual void beforeProcess();
    virtual void process();

private:
    static const int kChannels = B200RC_MAX_CHANNELS;

    void failOn(int status, const char* what);      ///< LERROR + VoreenException unless status == B200RC_OK
    void uploadVolume();                            ///< INVALID_PROGRAM on the volume port (volumeraycaster.cpp:44)
    void uploadTransferFunctions(float stepSize);   ///< hands the TFs over; the library re-uploads / rebuilds only what changed
    void fillParams(b200rc_params& p, b200rc_camera& cam, const tgt::ivec2& size, float stepSize);
    void writeOutport(RenderPort& port, const uint16_t* rgba16, const float* depth, const tgt::ivec2& size);
    void volumeInportOnChange();                    ///< singlevolumeraycaster.cpp:531-557
    void adjustPropertyVisibilities();              ///< singlevolumeraycaster.cpp:474-529 (incl. the pre-integration rejection rules)

    /// page-locked staging buffer (b200rc_host_alloc), grown on demand
    struct Pinned {
        void* ptr; size_t bytes;
        Pinned() : ptr(0), bytes(0) {}
        ~Pinned() { if (ptr) b200rc_host_free(ptr); }
        void* ensure(size_t n) {
            if (n > bytes) { if (ptr) b200rc_host_free(ptr); ptr = 0; bytes = 0; if (b200rc_host_alloc(n, &ptr) != B200RC_OK) return 0; bytes = n; }
            return ptr;
        }
    private:
        Pinned(const Pinned&); Pinned& operator=(const Pinned&);
    };

    TransFuncTypeProperty transFuncTypeProp_;
    std::unique_ptr<TransFunc1DKeysProperty> tf1d_[kChannels];
    std::unique_ptr<TransFunc2DPrimitivesProperty> tf2d_[kChannels];
    std::unique_ptr<TransFunc1DGaussianProperty> tfGauss_[kChannels];
    std::unique_ptr<FloatVec3Property> channelShift_[kChannels];
    BoolProperty enableChannelShift_;
    ShaderProperty shaderProp_;                     ///< inert; keeps "raycast.prg" deserialisable
    CameraProperty camera_;

    b200rc_ctx* ctx_;
    const VolumeBase* residentVolume_;              ///< volume currently held in HBM (re-upload whenever the port changed)
    const void* residentPreint_[kChannels];         ///< PreIntegrationTable handed over last (Voreen caches them per TF + step)
    Pinned pinEntry_, pinExit_, pinEntryDepth_, pinExitDepth_, pinColor_[B200RC_MAX_OUTPUTS], pinDepth_, pinScaled_, pinScaledDepth_;

    static const std::string loggerCat_;
};

} // namespace

#endif
