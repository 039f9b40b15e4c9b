// See the header for the contract.
#include "b200rgbraycaster.h"

#include "voreen/core/datastructures/volume/volumeram.h"
#include "voreen/core/datastructures/transfunc/1d/transfunc1d.h"
#include "voreen/core/utils/exception.h"
#include "tgt/textureunit.h"

#include <cstring>
#include <memory>

namespace voreen {

const std::string B200RGBRaycaster::loggerCat_("voreen.b200raycast.B200RGBRaycaster");

B200RGBRaycaster::B200RGBRaycaster()
    : VolumeRaycaster()
    , shader_("raycast.prg", "Raycasting Shader (unused)", "rc_rgb.frag", "passthrough.vert", "", Processor::VALID, Property::LOD_DEBUG)
    , transferFunc_("transferFunction", "Transfer function", Processor::INVALID_RESULT)
    , applyColorModulation_("applyColorModulation", "Apply color modulation", false, Processor::INVALID_RESULT)
    , camera_("camera", "Camera", tgt::Camera(tgt::vec3(0.f, 0.f, 3.5f), tgt::vec3(0.f, 0.f, 0.f), tgt::vec3(0.f, 1.f, 0.f)), true)
    , ctx_(0)
    , residentVolume_(0)
{
    volumeInport_.showTextureAccessProperties(true);
    addProperty(shader_);
    addProperty(transferFunc_);
    addProperty(applyColorModulation_);
    addProperty(camera_);       // required for the depth value calculation (rgbraycaster.cpp:47-49)
}

void B200RGBRaycaster::failOn(int status, const char* what) {
    if (status == B200RC_OK) return;
    const std::string msg = std::string(what) + ": " + (ctx_ ? b200rc_last_error(ctx_) : "no context");
    LERROR(msg);
    throw VoreenException(msg);
}

void B200RGBRaycaster::initialize() {
    VolumeRaycaster::initialize();
    if (b200rc_abi_version() != B200RC_ABI_VERSION) throw VoreenException("libb200rc ABI version mismatch");
    if (b200rc_create(0, &ctx_) != B200RC_OK) { ctx_ = 0; throw VoreenException("b200raycast: no usable B200 device (there is no CPU fallback)"); }
}

void B200RGBRaycaster::deinitialize() {
    if (ctx_) b200rc_destroy(ctx_);
    ctx_ = 0; residentVolume_ = 0;
    VolumeRaycaster::deinitialize();
}

void B200RGBRaycaster::beforeProcess() {
    VolumeRaycaster::beforeProcess();
    transferFunc_.setVolume(volumeInport_.getData());
    if (volumeInport_.hasChanged()) residentVolume_ = 0;       // new voxels (or a reused address), whatever the invalidation level
    if (volumeInport_.hasChanged() && volumeInport_.getData())
        camera_.adaptInteractionToScene(volumeInport_.getData()->getBoundingBox().getBoundingBox());
}

void B200RGBRaycaster::process() {
    const bool coarse = QualityMode.isInteractionMode() && interactionCoarseness_.get() > 1;        // rgbraycaster.cpp:63-75
    const tgt::ivec2 outSize = outport1_.getSize();
    const tgt::ivec2 size = coarse ? outSize / interactionCoarseness_.get() : outSize;
    const VolumeBase* vol = volumeInport_.getData();

    if (vol != residentVolume_) {
        const VolumeRAM* ram = vol->getRepresentation<VolumeRAM>();
        if (!ram) throw VoreenException("b200raycast: no VolumeRAM representation");
        b200rc_volume_desc d;
        std::memset(&d, 0, sizeof(d));
        const tgt::svec3 dim = ram->getDimensions();
        for (int a = 0; a < 3; ++a) d.dims[a] = (int32_t)dim[a];
        d.channels = (int32_t)ram->getNumChannels();
        const size_t bytesPerSample = ram->getBytesPerVoxel() / ram->getNumChannels();
        const std::string fmt = ram->getFormat();
        if (fmt.find("float") != std::string::npos && bytesPerSample == 4) d.dtype = B200RC_F32;
        else if (bytesPerSample == 2 && fmt.find("uint16") != std::string::npos) d.dtype = B200RC_U16;
        else if (bytesPerSample == 1 && fmt.find("uint8") != std::string::npos) d.dtype = B200RC_U8;
        else throw VoreenException("b200raycast: volume format '" + fmt + "' is not one the GL path uploads either (volumegl.cpp:122-139)");
        const tgt::vec3 sp = vol->getSpacing(), off = vol->getOffset();
        for (int a = 0; a < 3; ++a) { d.spacing[a] = sp[a]; d.offset[a] = off[a]; }
        std::memcpy(d.phys_to_world, vol->getPhysicalToWorldMatrix().elem, 16 * sizeof(float));
        const RealWorldMapping rwm = vol->getRealWorldMapping();
        d.rwm_scale = rwm.getScale(); d.rwm_offset = rwm.getOffset();
        failOn(b200rc_upload_volume(ctx_, 0, ram->getData(), &d), "b200rc_upload_volume");
        residentVolume_ = vol;
    }
    TransFunc1D* tf = transferFunc_.get();
    if (!tf) throw VoreenException("b200raycast: no transfer function");
    tgt::Texture* tex = tf->getTexture();
    if (!tex->getCpuTextureData()) tex->downloadTexture();
    const tgt::vec2 dom = tf->getDomain();
    failOn(b200rc_set_tf1d(ctx_, 0, tex->getCpuTextureData(), tex->getWidth(), dom.x, dom.y), "b200rc_set_tf1d");

    b200rc_params p;
    std::memset(&p, 0, sizeof(p));
    p.raycaster = B200RC_RAYCASTER_RGB;
    p.rgb_color_modulation = applyColorModulation_.get() ? 1 : 0;      // rgbraycaster.cpp:98
    p.num_volumes = 1;
    p.vol[0].classification = B200RC_CLASS_TF;
    p.vol[0].filter = volumeInport_.getTextureFilterModeProperty().getValue() == GL_NEAREST ? B200RC_FILTER_NEAREST : B200RC_FILTER_LINEAR;
    p.vol[0].clamp = volumeInport_.getTextureClampModeProperty().getValue() == GL_CLAMP_TO_BORDER ? B200RC_CLAMP_TO_BORDER : B200RC_CLAMP_TO_EDGE;
    p.vol[0].border_intensity = volumeInport_.getTextureBorderIntensityProperty().get();
    p.sampling_step = getSamplingStepSize(vol);             // bindVolumes sets samplingStepSize_ (volumeraycaster.cpp:312-322)
    p.output_connected[0] = 1;
    p.sampler = B200RC_SAMPLER_EXACT;
    p.empty_space_skipping = 1;
    p.out_format = B200RC_OUT_RGBA16;
    p.quantize_depth24 = 1;

    b200rc_camera cam;
    const tgt::Camera c = camera_.get();
    std::memcpy(cam.view, c.getViewMatrix().elem, 16 * sizeof(float));
    std::memcpy(cam.proj, c.getProjectionMatrix(size).elem, 16 * sizeof(float));
    const tgt::vec3 pos = c.getPositionWithOffsets();
    cam.position[0] = pos.x; cam.position[1] = pos.y; cam.position[2] = pos.z;
    cam.near_dist = c.getNearDist(); cam.far_dist = c.getFarDist();

    const size_t px = (size_t)size.x * size.y;
    std::unique_ptr<GLubyte[]> entry(entryPort_.getColorTexture()->downloadTextureToBuffer(GL_RGBA, GL_FLOAT));
    std::unique_ptr<GLubyte[]> exitp(exitPort_.getColorTexture()->downloadTextureToBuffer(GL_RGBA, GL_FLOAT));
    std::unique_ptr<GLubyte[]> entryD(entryPort_.getDepthTexture()->downloadTextureToBuffer(GL_DEPTH_COMPONENT, GL_FLOAT));
    std::unique_ptr<GLubyte[]> exitD(exitPort_.getDepthTexture()->downloadTextureToBuffer(GL_DEPTH_COMPONENT, GL_FLOAT));
    hostColor_.resize(px * 4); hostDepth_.resize(px);
    void* outColor[B200RC_MAX_OUTPUTS] = {&hostColor_[0], 0, 0};
    failOn(b200rc_render(ctx_, &p, &cam, size.x, size.y, reinterpret_cast<const float*>(entry.get()), reinterpret_cast<const float*>(entryD.get()),
                         reinterpret_cast<const float*>(exitp.get()), reinterpret_cast<const float*>(exitD.get()), outColor, &hostDepth_[0]), "b200rc_render");

    const uint16_t* colData = &hostColor_[0];
    const float* depthData = &hostDepth_[0];
    std::vector<uint16_t> up; std::vector<float> upDepth;
    if (outSize != size) {      // rescaleRendering (rgbraycaster.cpp:118-121)
        up.resize((size_t)outSize.x * outSize.y * 4); upDepth.resize((size_t)outSize.x * outSize.y);
        failOn(b200rc_rescale(ctx_, colData, depthData, size.x, size.y, B200RC_OUT_RGBA16, &up[0], &upDepth[0], outSize.x, outSize.y, 0), "b200rc_rescale");
        colData = &up[0]; depthData = &upDepth[0];
    }
    outport1_.getColorTexture()->bind();
    glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, outSize.x, outSize.y, GL_RGBA, GL_UNSIGNED_SHORT, colData);
    outport1_.getDepthTexture()->bind();
    glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, outSize.x, outSize.y, GL_DEPTH_COMPONENT, GL_FLOAT, depthData);
    outport1_.validateResult();
    tgt::TextureUnit::setZeroUnit();
    LGL_ERROR;
}

} // namespace
