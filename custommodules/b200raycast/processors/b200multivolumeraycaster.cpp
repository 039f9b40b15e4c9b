// See the header for the contract.  Every block cites the GLSL-path code it stands in for.
#include "b200multivolumeraycaster.h"

#include "voreen/core/datastructures/volume/volumeram.h"
#include "voreen/core/datastructures/transfunc/1d/transfunc1d.h"
#include "voreen/core/datastructures/transfunc/1d/preintegrationtable.h"
#include "voreen/core/ports/conditions/portconditionvolumetype.h"
#include "voreen/core/utils/classificationmodes.h"
#include "voreen/core/utils/glsl.h"
#include "voreen/core/utils/exception.h"
#include "tgt/textureunit.h"

#include <algorithm>
#include <cstring>
#include <memory>

namespace voreen {

const std::string B200MultiVolumeRaycaster::loggerCat_("voreen.b200raycast.B200MultiVolumeRaycaster");

namespace {
struct Opt { const char* key; int value; };
const Opt kGradient[] = {{"none", B200RC_GRAD_NONE}, {"forward-differences", B200RC_GRAD_FORWARD}, {"central-differences", B200RC_GRAD_CENTRAL},
                         {"sobel", B200RC_GRAD_SOBEL}, {"filtered", B200RC_GRAD_FILTERED}};
const Opt kShading[] = {{"none", B200RC_SHADE_NONE}, {"phong-diffuse", B200RC_SHADE_PHONG_D}, {"phong-specular", B200RC_SHADE_PHONG_S},
                        {"phong-diffuse-ambient", B200RC_SHADE_PHONG_DA}, {"phong-diffuse-specular", B200RC_SHADE_PHONG_DS},
                        {"phong", B200RC_SHADE_PHONG}, {"toon", B200RC_SHADE_TOON}, {"cook-torrance", B200RC_SHADE_COOK_TORRANCE},
                        {"oren-nayar", B200RC_SHADE_OREN_NAYAR}, {"ward", B200RC_SHADE_WARD}};
const Opt kCompositing[] = {{"dvr", B200RC_COMP_DVR}, {"mip", B200RC_COMP_MIP}, {"mop", B200RC_COMP_MOP}, {"mida", B200RC_COMP_MIDA},
                            {"iso", B200RC_COMP_ISO}, {"fhp", B200RC_COMP_FHP}, {"fhn", B200RC_COMP_FHN}};
template <size_t N>
int lookup(const Opt (&table)[N], const std::string& key, const char* what) {
    for (size_t i = 0; i < N; ++i)
        if (key == table[i].key) return table[i].value;
    throw VoreenException(std::string("b200raycast: unsupported ") + what + " option '" + key + "'");
}
void copyMat(const tgt::mat4& m, float out[16]) { std::memcpy(out, m.elem, 16 * sizeof(float)); }   // tgt::mat4 is row-major

// getVoxelSamplingStepSize (multivolumeraycaster.cpp:221-223)
float voxelSamplingStepSize(const VolumeBase* vol, float worldSamplingStepSize) {
    return tgt::min(worldSamplingStepSize * vol->getPhysicalToTextureMatrix().getScalingPart());
}
} // namespace

B200MultiVolumeRaycaster::B200MultiVolumeRaycaster()
    : VolumeRaycaster(true, true)
    , volumeInport2_(Port::INPORT, "volume2", "Volume2 Input", false, Processor::INVALID_PROGRAM)
    , entryPort2_(Port::INPORT, "entryPort2", "Entry-points of Volume2", false, Processor::INVALID_RESULT)
    , exitPort2_(Port::INPORT, "exitPort2", "Exit-points of Volume2", false, Processor::INVALID_RESULT)
    , volumeInport3_(Port::INPORT, "volume3", "Volume3 Input", false, Processor::INVALID_PROGRAM)
    , entryPort3_(Port::INPORT, "entryPort3", "Entry-points of Volume3", false, Processor::INVALID_RESULT)
    , exitPort3_(Port::INPORT, "exitPort3", "Exit-points of Volume3", false, Processor::INVALID_RESULT)
    , volumeInport4_(Port::INPORT, "volume4", "Volume4 Input", false, Processor::INVALID_PROGRAM)
    , entryPort4_(Port::INPORT, "entryPort4", "Entry-points of Volume4", false, Processor::INVALID_RESULT)
    , exitPort4_(Port::INPORT, "exitPort4", "Exit-points of Volume4", false, Processor::INVALID_RESULT)
    , entryPortGlobal_(Port::INPORT, "entryPortGlobal", "Global Entry-points Input", false, Processor::INVALID_RESULT)
    , exitPortGlobal_(Port::INPORT, "exitPortGlobal", "Global Exit-points Input", false, Processor::INVALID_RESULT)
    , shaderProp_("raycast.prg", "Raycasting Shader (unused)", "rc_multivolume.frag", "passthrough.vert", "", Processor::VALID, Property::LOD_DEBUG)
    , classificationMode2_("classification2", "Classification 2", Processor::INVALID_PROGRAM)
    , classificationMode3_("classification3", "Classification 3", Processor::INVALID_PROGRAM)
    , classificationMode4_("classification4", "Classification 4", Processor::INVALID_PROGRAM)
    , shadeMode1_("shading1", "Shading 1", Processor::INVALID_PROGRAM)
    , shadeMode2_("shading2", "Shading 2", Processor::INVALID_PROGRAM)
    , shadeMode3_("shading3", "Shading 3", Processor::INVALID_PROGRAM)
    , shadeMode4_("shading4", "Shading 4", Processor::INVALID_PROGRAM)
    , transferFunc1_("transferFunction1", "Transfer Function 1")
    , transferFunc2_("transferFunction2", "Transfer Function 2")
    , transferFunc3_("transferFunction3", "Transfer Function 3")
    , transferFunc4_("transferFunction4", "Transfer Function 4")
    , camera_("camera", "Camera", tgt::Camera(tgt::vec3(0.f, 0.f, 3.5f), tgt::vec3(0.f, 0.f, 0.f), tgt::vec3(0.f, 1.f, 0.f)), true)
    , ctx_(0)
    , mvOpacityCorrectionFactor_(0.f)
{
    for (int i = 0; i < kVolumes; ++i) resident_[i] = 0;
    // ports and properties in the order of multivolumeraycaster.cpp:66-135
    volumeInport2_.addCondition(new PortConditionVolumeTypeGL());
    volumeInport3_.addCondition(new PortConditionVolumeTypeGL());
    volumeInport4_.addCondition(new PortConditionVolumeTypeGL());
    volumeInport_.showTextureAccessProperties(true);
    volumeInport2_.showTextureAccessProperties(true);
    volumeInport3_.showTextureAccessProperties(true);
    volumeInport4_.showTextureAccessProperties(true);
    addPort(volumeInport2_); addPort(entryPort2_); addPort(exitPort2_);
    addPort(volumeInport3_); addPort(entryPort3_); addPort(exitPort3_);
    addPort(volumeInport4_); addPort(entryPort4_); addPort(exitPort4_);
    addPort(entryPortGlobal_);
    addPort(exitPortGlobal_);

    addProperty(shaderProp_);
    addProperty(classificationMode_);
    ClassificationModes::fillProperty(&classificationMode2_); addProperty(classificationMode2_);
    ClassificationModes::fillProperty(&classificationMode3_); addProperty(classificationMode3_);
    ClassificationModes::fillProperty(&classificationMode4_); addProperty(classificationMode4_);
    addProperty(transferFunc1_); addProperty(transferFunc2_); addProperty(transferFunc3_); addProperty(transferFunc4_);
    addProperty(camera_);
    addProperty(gradientMode_);
    GLSL::fillShadingModesProperty(shadeMode1_); addProperty(shadeMode1_);
    GLSL::fillShadingModesProperty(shadeMode2_); addProperty(shadeMode2_);
    GLSL::fillShadingModesProperty(shadeMode3_); addProperty(shadeMode3_);
    GLSL::fillShadingModesProperty(shadeMode4_); addProperty(shadeMode4_);
    addProperty(compositingMode1_); addProperty(compositingMode2_); addProperty(compositingMode3_);
    addProperty(isoValue_);
    Property* lighting[] = {&lightPosition_, &lightAmbient_, &lightDiffuse_, &lightSpecular_, &materialShininess_,
                            &applyLightAttenuation_, &lightAttenuation_};
    for (size_t i = 0; i < sizeof(lighting) / sizeof(lighting[0]); ++i) { addProperty(lighting[i]); lighting[i]->setGroupID("lighting"); }
    setPropertyGroupGuiName("lighting", "Lighting Parameters");
}

void B200MultiVolumeRaycaster::failOn(int status, const char* what) {
    if (status == B200RC_OK) return;
    const std::string msg = std::string(what) + ": " + (ctx_ ? b200rc_last_error(ctx_) : "no context");
    LERROR(msg);
    throw VoreenException(msg);     // caught per processor by the NetworkEvaluator (networkevaluator.cpp:383-394)
}

void B200MultiVolumeRaycaster::initialize() {
    VolumeRaycaster::initialize();
    if (b200rc_abi_version() != B200RC_ABI_VERSION)
        throw VoreenException("libb200rc ABI version mismatch");
    if (b200rc_create(0, &ctx_) != B200RC_OK) {
        ctx_ = 0;
        throw VoreenException("b200raycast: no usable B200 device (there is no CPU fallback)");
    }
}

void B200MultiVolumeRaycaster::deinitialize() {
    if (ctx_) b200rc_destroy(ctx_);
    ctx_ = 0;
    for (int i = 0; i < kVolumes; ++i) resident_[i] = 0;
    VolumeRaycaster::deinitialize();
}

// multivolumeraycaster.cpp:166-188: global entry/exit; every volume either fully connected or not at all; one outport
bool B200MultiVolumeRaycaster::isReady() const {
    if (!ctx_ || !entryPortGlobal_.isReady() || !exitPortGlobal_.isReady()) return false;
    const VolumePort* vp[kVolumes] = {&volumeInport_, &volumeInport2_, &volumeInport3_, &volumeInport4_};
    const RenderPort* ep[kVolumes] = {&entryPort_, &entryPort2_, &entryPort3_, &entryPort4_};
    const RenderPort* xp[kVolumes] = {&exitPort_, &exitPort2_, &exitPort3_, &exitPort4_};
    for (int i = 0; i < kVolumes; ++i) {
        const bool all = vp[i]->isReady() && ep[i]->isReady() && xp[i]->isReady();
        const bool none = !vp[i]->isReady() && !ep[i]->isReady() && !xp[i]->isReady();
        if (!all && !none) return false;
    }
    return outport1_.isReady() || outport2_.isReady() || outport3_.isReady();
}

void B200MultiVolumeRaycaster::beforeProcess() {
    VolumeRaycaster::beforeProcess();
    // INVALID_PROGRAM rebuilds the GLSL program, which resets its uniforms to 0 -- mvOpacityCorrectionFactor_ included
    if (getInvalidationLevel() >= Processor::INVALID_PROGRAM) mvOpacityCorrectionFactor_ = 0.f;
    transferFunc1_.setVolume(volumeInport_.getData());
    transferFunc2_.setVolume(volumeInport2_.getData());
    transferFunc3_.setVolume(volumeInport3_.getData());
    transferFunc4_.setVolume(volumeInport4_.getData());
    if (volumeInport_.hasChanged() || volumeInport2_.hasChanged() || volumeInport3_.hasChanged() || volumeInport4_.hasChanged()) {
        tgt::Bounds b;
        const VolumePort* vp[kVolumes] = {&volumeInport_, &volumeInport2_, &volumeInport3_, &volumeInport4_};
        for (int i = 0; i < kVolumes; ++i)
            if (vp[i]->hasChanged()) resident_[i] = 0;     // new voxels (or a reused address), whatever the invalidation level
        for (int i = 0; i < kVolumes; ++i)
            if (vp[i]->hasData()) b.addVolume(vp[i]->getData()->getBoundingBox().getBoundingBox());
        if (length(b.diagonal()) != 0.0) camera_.adaptInteractionToScene(b);
    }
}

void B200MultiVolumeRaycaster::uploadVolume(int slot, const VolumeBase* vol) {
    if (resident_[slot] == vol) return;
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
    copyMat(vol->getPhysicalToWorldMatrix(), d.phys_to_world);
    const RealWorldMapping rwm = vol->getRealWorldMapping();
    d.rwm_scale = rwm.getScale(); d.rwm_offset = rwm.getOffset();
    failOn(b200rc_upload_volume(ctx_, slot, ram->getData(), &d), "b200rc_upload_volume");
    resident_[slot] = vol;
}

void B200MultiVolumeRaycaster::process() {
    const bool coarse = QualityMode.isInteractionMode() && interactionCoarseness_.get() > 1;        // multivolumeraycaster.cpp:227-229
    const tgt::ivec2 size = coarse ? outport1_.getSize() / interactionCoarseness_.get() : outport1_.getSize();
    const size_t px = (size_t)size.x * size.y;

    VolumePort* vp[kVolumes] = {&volumeInport_, &volumeInport2_, &volumeInport3_, &volumeInport4_};
    RenderPort* ep[kVolumes] = {&entryPort_, &entryPort2_, &entryPort3_, &entryPort4_};
    RenderPort* xp[kVolumes] = {&exitPort_, &exitPort2_, &exitPort3_, &exitPort4_};
    StringOptionProperty* cls[kVolumes] = {&classificationMode_, &classificationMode2_, &classificationMode3_, &classificationMode4_};
    StringOptionProperty* shade[kVolumes] = {&shadeMode1_, &shadeMode2_, &shadeMode3_, &shadeMode4_};
    TransFunc1DKeysProperty* tfp[kVolumes] = {&transferFunc1_, &transferFunc2_, &transferFunc3_, &transferFunc4_};

    // the connected volumes, in port order, become device slots 0..n-1 (VOLUME_n_ACTIVE, multivolumeraycaster.cpp:419-445;
    // the shader visits the active volumes in that order at every sample)
    int active[kVolumes], n = 0;
    for (int i = 0; i < kVolumes; ++i)
        if (vp[i]->isReady()) active[n++] = i;
    if (n == 0) return;

    // world-space step of the finest volume, its texture-space step for the opacity correction (:326-344)
    float voxelSizeWorld = 999.f, voxelSizeTexture = 999.f;
    for (int s = 0; s < n; ++s) {
        const VolumeBase* v = vp[active[s]]->getData();
        const tgt::vec3 dim(v->getDimensions());
        const tgt::vec3 cubeSizeWorld = v->getCubeSize() * v->getPhysicalToWorldMatrix().getScalingPart();
        const float t = tgt::max(cubeSizeWorld / dim);
        if (t < voxelSizeWorld) { voxelSizeWorld = t; voxelSizeTexture = tgt::max(1.f / dim); }
    }
    float stepWorld = voxelSizeWorld / samplingRate_.get(), stepTexture = voxelSizeTexture / samplingRate_.get();
    if (QualityMode.isInteractionMode()) { stepWorld /= interactionQuality_.get(); stepTexture /= interactionQuality_.get(); }
    // the uniform is only written while some connected output composites with DVR (:352-357); otherwise it keeps its value
    if (compositingMode1_.isSelected("dvr") || (compositingMode2_.isSelected("dvr") && outport1_.isConnected()) ||
        (compositingMode3_.isSelected("dvr") && outport2_.isConnected()))
        mvOpacityCorrectionFactor_ = stepTexture / stepWorld;

    b200rc_params p;
    std::memset(&p, 0, sizeof(p));
    p.multi_volume = 1;
    p.num_volumes = n;
    p.sampling_step = stepWorld;
    p.mv_opacity_correction_factor = mvOpacityCorrectionFactor_;
    p.gradient = lookup(kGradient, gradientMode_.get(), "gradient");
    for (int s = 0; s < n; ++s) {
        const int i = active[s];
        const VolumeBase* v = vp[i]->getData();
        if (resident_[s] != v) uploadVolume(s, v);
        b200rc_volume_params& q = p.vol[s];
        const std::string c = cls[i]->get();
        q.classification = (c == "transfer-function") ? B200RC_CLASS_TF : B200RC_CLASS_PREINT;
        q.shading = lookup(kShading, shade[i]->get(), "shading");
        q.filter = vp[i]->getTextureFilterModeProperty().getValue() == GL_NEAREST ? B200RC_FILTER_NEAREST : B200RC_FILTER_LINEAR;
        q.clamp = vp[i]->getTextureClampModeProperty().getValue() == GL_CLAMP_TO_BORDER ? B200RC_CLAMP_TO_BORDER : B200RC_CLAMP_TO_EDGE;
        q.border_intensity = vp[i]->getTextureBorderIntensityProperty().get();
        // ClassificationModes::bindTexture with the per-volume voxel step (:366-378)
        TransFunc1D* tf = tfp[i]->get();
        tgt::Texture* tex = tf->getTexture();
        if (!tex->getCpuTextureData()) tex->downloadTexture();
        const tgt::vec2 dom = tf->getDomain();
        failOn(b200rc_set_tf1d(ctx_, s, tex->getCpuTextureData(), tex->getWidth(), dom.x, dom.y), "b200rc_set_tf1d");
        const float voxelStep = voxelSamplingStepSize(v, stepWorld);
        if (c == "pre-integrated-gpu")
            failOn(b200rc_build_preint(ctx_, s, std::min(tex->getWidth(), 1024), voxelStep, B200RC_PI_GPU), "b200rc_build_preint");
        else if (c == "pre-integrated-cpu" || c == "pre-integrated-approximation") {
            const PreIntegrationTable* t = tf->getPreIntegrationTable(voxelStep, 0, c == "pre-integrated-approximation", false);
            failOn(b200rc_set_preint(ctx_, s, reinterpret_cast<const float*>(t->getTable()), (int)t->getDimension()), "b200rc_set_preint");
        }
    }
    for (int s = n; s < kVolumes; ++s)
        if (resident_[s]) { b200rc_release_volume(ctx_, s); resident_[s] = 0; }

    const StringOptionProperty* comp[3] = {&compositingMode1_, &compositingMode2_, &compositingMode3_};
    const RenderPort* out[3] = {&outport1_, &outport2_, &outport3_};
    for (int i = 0; i < 3; ++i) {
        p.compositing[i] = lookup(kCompositing, comp[i]->get(), "compositing");
        p.output_connected[i] = out[i]->isConnected() ? 1 : 0;
    }
    p.iso_value = isoValue_.get();          // rc_multivolume.frag:180-182 passes literal 0 for the MIDA gamma and the iso value; the library does too
    const tgt::vec4 lp = lightPosition_.get();
    for (int i = 0; i < 4; ++i) p.light.position[i] = lp[i];
    for (int i = 0; i < 3; ++i) {
        p.light.ambient[i] = lightAmbient_.get()[i]; p.light.diffuse[i] = lightDiffuse_.get()[i];
        p.light.specular[i] = lightSpecular_.get()[i]; p.light.attenuation[i] = lightAttenuation_.get()[i];
    }
    p.light.shininess = materialShininess_.get();
    p.light.apply_attenuation = applyLightAttenuation_.get() ? 1 : 0;
    p.sampler = B200RC_SAMPLER_EXACT;
    p.empty_space_skipping = 1;         // per-volume skipping of TF-transparent cells (pixels identical)
    p.out_format = B200RC_OUT_RGBA16;
    p.quantize_depth24 = 1;

    b200rc_camera cam;
    const tgt::Camera c = camera_.get();
    copyMat(c.getViewMatrix(), cam.view);
    copyMat(c.getProjectionMatrix(size), cam.proj);
    const tgt::vec3 pos = c.getPositionWithOffsets();
    cam.position[0] = pos.x; cam.position[1] = pos.y; cam.position[2] = pos.z;
    cam.near_dist = c.getNearDist(); cam.far_dist = c.getFarDist();

    // entry/exit: [global, volume slots 0..n-1], RGBA32F each; depth of the global images only (rc_multivolume.frag:97-131)
    entry_.resize(px * 4 * (size_t)(n + 1)); exit_.resize(px * 4 * (size_t)(n + 1));
    auto grab = [&](RenderPort& port, float* dst) {
        std::unique_ptr<GLubyte[]> b(port.getColorTexture()->downloadTextureToBuffer(GL_RGBA, GL_FLOAT));
        std::memcpy(dst, b.get(), px * 4 * sizeof(float));
    };
    grab(entryPortGlobal_, &entry_[0]); grab(exitPortGlobal_, &exit_[0]);
    for (int s = 0; s < n; ++s) { grab(*ep[active[s]], &entry_[px * 4 * (size_t)(s + 1)]); grab(*xp[active[s]], &exit_[px * 4 * (size_t)(s + 1)]); }
    std::unique_ptr<GLubyte[]> entryD(entryPortGlobal_.getDepthTexture()->downloadTextureToBuffer(GL_DEPTH_COMPONENT, GL_FLOAT));
    std::unique_ptr<GLubyte[]> exitD(exitPortGlobal_.getDepthTexture()->downloadTextureToBuffer(GL_DEPTH_COMPONENT, GL_FLOAT));

    void* outColor[B200RC_MAX_OUTPUTS] = {0, 0, 0};
    for (int i = 0; i < B200RC_MAX_OUTPUTS; ++i)
        if (p.output_connected[i]) { hostColor_[i].resize(px * 4); outColor[i] = &hostColor_[i][0]; }
    hostDepth_.resize(px);
    failOn(b200rc_render(ctx_, &p, &cam, size.x, size.y, &entry_[0], reinterpret_cast<const float*>(entryD.get()),
                         &exit_[0], reinterpret_cast<const float*>(exitD.get()), outColor, &hostDepth_[0]), "b200rc_render");

    RenderPort* outp[3] = {&outport1_, &outport2_, &outport3_};
    for (int i = 0; i < 3; ++i) {
        if (!p.output_connected[i]) continue;
        if (i == 2 && !outport2_.isConnected()) continue;       // quirk: internalRenderPort3_ is copied only if outport2_ is connected (:412-413)
        // rescaleRendering (volumeraycaster.cpp:348-377): internal size -> outport size with GL_LINEAR; the identity at equal size
        const tgt::ivec2 outSize = outp[i]->getSize();
        const uint16_t* colData = &hostColor_[i][0];
        const float* depthData = &hostDepth_[0];
        std::vector<uint16_t> up; std::vector<float> upDepth;
        if (outSize != size) {
            up.resize((size_t)outSize.x * outSize.y * 4);
            if (i == 0) upDepth.resize((size_t)outSize.x * outSize.y);
            failOn(b200rc_rescale(ctx_, colData, i == 0 ? depthData : 0, size.x, size.y, B200RC_OUT_RGBA16, &up[0], i == 0 ? &upDepth[0] : 0,
                                  outSize.x, outSize.y, 0), "b200rc_rescale");
            colData = &up[0]; depthData = i == 0 ? &upDepth[0] : 0;
        }
        tgt::Texture* col = outp[i]->getColorTexture();
        col->bind();
        glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, outSize.x, outSize.y, GL_RGBA, GL_UNSIGNED_SHORT, colData);
        if (i == 0) {
            tgt::Texture* dt = outp[i]->getDepthTexture();
            dt->bind();
            glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, outSize.x, outSize.y, GL_DEPTH_COMPONENT, GL_FLOAT, depthData);
        }
        outp[i]->validateResult();
    }
    tgt::TextureUnit::setZeroUnit();
    LGL_ERROR;
}

} // namespace
