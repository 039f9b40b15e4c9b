// See the header for the contract.  Every block cites the GLSL-path code it stands in for.
#include "b200singlevolumeraycaster.h"

#include "voreen/core/datastructures/volume/volumeram.h"
#include "voreen/core/datastructures/transfunc/1d/transfunc1d.h"
#include "voreen/core/datastructures/transfunc/1d/preintegrationtable.h"
#include "voreen/core/datastructures/callback/memberfunctioncallback.h"
#include "voreen/core/utils/classificationmodes.h"
#include "voreen/core/utils/exception.h"
#include "tgt/textureunit.h"

#include <cstring>

namespace voreen {

const std::string B200SingleVolumeRaycaster::loggerCat_("voreen.b200raycast.B200SingleVolumeRaycaster");

namespace {

// option string -> ABI enum tables; the strings are the property option keys of the reference
// (gradient: volumeraycaster.cpp:126-131, shading: glsl.cpp:397-409, compositing: volumeraycaster.cpp:379-388)
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

const char* kTf1dIds[] = {"transferFunction", "transferFunction1", "transferFunction2", "transferFunction3"};
const char* kTf2dIds[] = {"transFunc2D", "transFunc2D_2", "transFunc2D_3", "transFunc2D_4"};
const char* kTfGaussIds[] = {"transFuncGaussian1", "transFuncGaussian2", "transFuncGaussian3", "transFuncGaussian4"};
const char* kShiftIds[] = {"channelShift1", "channelShift2", "channelShift3", "channelShift4"};

} // namespace

B200SingleVolumeRaycaster::B200SingleVolumeRaycaster()
    : VolumeRaycaster(true, true)       // default properties + port group, as singlevolumeraycaster.cpp:47
    , transFuncTypeProp_("transFuncType", "TF Mode")
    , enableChannelShift_("enableChannelShift", "Enable Channel Shift", false, Processor::INVALID_PROGRAM, Property::LOD_ADVANCED)
    , shaderProp_("raycast.prg", "Raycasting Shader (unused)", "rc_singlevolume.frag", "passthrough.vert", "", Processor::VALID, Property::LOD_DEBUG)
    , camera_("camera", "Camera", tgt::Camera(tgt::vec3(0.f, 0.f, 3.5f), tgt::vec3(0.f), tgt::vec3(0.f, 1.f, 0.f)), true)
    , ctx_(0)
    , residentVolume_(0)
{
    for (int c = 0; c < kChannels; ++c) residentPreint_[c] = 0;
    volumeInport_.onChange(MemberFunctionCallback<B200SingleVolumeRaycaster>(this, &B200SingleVolumeRaycaster::volumeInportOnChange));
    addProperty(shaderProp_);
    addProperty(transFuncTypeProp_);
    for (int c = 0; c < kChannels; ++c) {
        const std::string n = " (Channel " + std::string(1, char('1' + c)) + ")";
        tf1d_[c].reset(new TransFunc1DKeysProperty(kTf1dIds[c], "Transfer Function 1D" + n));
        tf2d_[c].reset(new TransFunc2DPrimitivesProperty(kTf2dIds[c], "Transfer Function 2D" + n));
        tfGauss_[c].reset(new TransFunc1DGaussianProperty(kTfGaussIds[c], "Transfer Function 1D Gaussian" + n));
        channelShift_[c].reset(new FloatVec3Property(kShiftIds[c], "Channel Shift" + n, tgt::vec3::zero, tgt::vec3(-100.f), tgt::vec3(100.f)));
    }
    for (int c = 0; c < kChannels; ++c) { addProperty(*tf1d_[c]); tf1d_[c]->setGroupID("tf"); }
    for (int c = 0; c < kChannels; ++c) { addProperty(*tf2d_[c]); tf2d_[c]->setGroupID("tf"); }
    for (int c = 0; c < kChannels; ++c) { addProperty(*tfGauss_[c]); tfGauss_[c]->setGroupID("tf"); }
    transFuncTypeProp_.addTransFuncProperty(tf1d_[0].get());
    transFuncTypeProp_.addTransFuncProperty(tf2d_[0].get());
    transFuncTypeProp_.addTransFuncProperty(tfGauss_[0].get());
    transFuncTypeProp_.setGroupID("tf");
    setPropertyGroupGuiName("tf", "Transfer Function");

    addProperty(camera_);
    addProperty(gradientMode_);
    addProperty(classificationMode_);
    addProperty(shadeMode_);
    addProperty(gammaValue1_); addProperty(gammaValue2_); addProperty(gammaValue3_);
    addProperty(compositingMode1_); addProperty(compositingMode2_); addProperty(compositingMode3_);
    addProperty(isoValue_);

    Property* lighting[] = {&lightPosition_, &lightAmbient_, &lightDiffuse_, &lightSpecular_, &materialShininess_,
                            &applyLightAttenuation_, &lightAttenuation_};
    for (size_t i = 0; i < sizeof(lighting) / sizeof(lighting[0]); ++i) { addProperty(lighting[i]); lighting[i]->setGroupID("lighting"); }
    setPropertyGroupGuiName("lighting", "Lighting Parameters");

    addProperty(enableChannelShift_); enableChannelShift_.setGroupID("channelShift");
    for (int c = 0; c < kChannels; ++c) { addProperty(*channelShift_[c]); channelShift_[c]->setGroupID("channelShift"); }
    setPropertyGroupGuiName("channelShift", "Channel Shift");

    // same listeners as singlevolumeraycaster.cpp:160-168: GUI state and the classification rules follow these properties
    const MemberFunctionCallback<B200SingleVolumeRaycaster> adjust(this, &B200SingleVolumeRaycaster::adjustPropertyVisibilities);
    classificationMode_.onChange(adjust); shadeMode_.onChange(adjust);
    compositingMode1_.onChange(adjust); compositingMode2_.onChange(adjust); compositingMode3_.onChange(adjust);
    applyLightAttenuation_.onChange(adjust); transFuncTypeProp_.onChange(adjust); enableChannelShift_.onChange(adjust);
    volumeInport_.onNewData(adjust);
}

// singlevolumeraycaster.cpp:531-557: camera interaction + light range follow the scene, every TF property learns the volume
// (TransFunc1DKeysProperty::setVolume resizes the TF texture to 2^bits texels and fits the domain on first use,
// transfunc1dkeysproperty.cpp:94-140) -- without this an unchanged workspace would classify through a texture of the wrong width
void B200SingleVolumeRaycaster::volumeInportOnChange() {
    if (volumeInport_.hasData()) {
        const tgt::Bounds bb = volumeInport_.getData()->getBoundingBox().getBoundingBox();
        camera_.adaptInteractionToScene(bb, tgt::min(volumeInport_.getData()->getSpacing()));
        lightPosition_.setMaxDist(10.f * tgt::length(bb.diagonal()));       // LIGHT_POSITION_MAX_DIST_MULTIPLIER (:40)
    }
    const int channels = volumeInport_.hasData() ? (int)volumeInport_.getData()->getNumChannels() : 1;
    for (int c = 0; c < kChannels; ++c) {
        if (c > 0 && channels < c + 1) break;
        tf1d_[c]->setVolume(volumeInport_.getData(), c);
        tf2d_[c]->setVolume(volumeInport_.getData(), c);
        tfGauss_[c]->setVolume(volumeInport_.getData(), c);
    }
    residentVolume_ = 0;        // whatever the invalidation level: a changed port means new voxels (or a reused address)
}

// singlevolumeraycaster.cpp:474-529
void B200SingleVolumeRaycaster::adjustPropertyVisibilities() {
    const int channels = volumeInport_.hasData() ? (int)volumeInport_.getData()->getNumChannels() : 1;
    const bool tf1d = transFuncTypeProp_.getValue() == TFT_1DKEYS, tf2d = transFuncTypeProp_.getValue() == TFT_2DPRIMITIVES,
               tfg = transFuncTypeProp_.getValue() == TFT_1DGAUSSIAN;
    for (int c = 0; c < kChannels; ++c) {
        tf1d_[c]->setVisibleFlag(channels >= c + 1 && tf1d);
        tf2d_[c]->setVisibleFlag(channels >= c + 1 && tf2d);
        tfGauss_[c]->setVisibleFlag(channels >= c + 1 && tfg);
        channelShift_[c]->setVisibleFlag(enableChannelShift_.get() && channels >= c + 1);
    }
    lightAttenuation_.setVisibleFlag(applyLightAttenuation_.get());
    gammaValue1_.setVisibleFlag(compositingMode1_.isSelected("mida"));
    gammaValue2_.setVisibleFlag(compositingMode2_.isSelected("mida"));
    gammaValue3_.setVisibleFlag(compositingMode3_.isSelected("mida"));
    // the two rejection rules of the reference (:516-528): no pre-integration with 2D transfer functions or with MIP
    if (tf2d && !(classificationMode_.get() == "transfer-function")) {
        LWARNING("Pre-integration is not allowed with 2d transfer functions. Switching to normal transfer function mode.");
        classificationMode_.select("transfer-function");
    }
    if ((compositingMode1_.get() == "mip" || compositingMode2_.get() == "mip" || compositingMode3_.get() == "mip")
        && !(classificationMode_.get() == "transfer-function")) {
        LWARNING("Pre-integration is not allowed with Mip rendering. Switching to normal transfer function mode.");
        classificationMode_.select("transfer-function");
    }
}

B200SingleVolumeRaycaster::~B200SingleVolumeRaycaster() {}

void B200SingleVolumeRaycaster::failOn(int status, const char* what) {
    if (status == B200RC_OK) return;
    const std::string msg = std::string(what) + ": " + (ctx_ ? b200rc_last_error(ctx_) : "no context");
    LERROR(msg);
    // NetworkEvaluator catches per processor (V/src/core/network/networkevaluator.cpp:383-394); nothing crosses the C ABI
    throw VoreenException(msg);
}

// Processor owns its GPU context from initialize() to deinitialize() (pattern: V/modules/opencl/processors/openclprocessor.cpp:44-58).
void B200SingleVolumeRaycaster::initialize() {
    VolumeRaycaster::initialize();
    if (b200rc_abi_version() != B200RC_ABI_VERSION)
        throw VoreenException("libb200rc ABI version mismatch");
    const int status = b200rc_create(0, &ctx_);
    if (status != B200RC_OK) {
        ctx_ = 0;
        throw VoreenException("b200raycast: no usable B200 device (there is no CPU fallback)");
    }
    adjustPropertyVisibilities();       // singlevolumeraycaster.cpp:173-178
}

void B200SingleVolumeRaycaster::deinitialize() {
    if (ctx_) b200rc_destroy(ctx_);
    ctx_ = 0;
    residentVolume_ = 0;
    for (int c = 0; c < kChannels; ++c) residentPreint_[c] = 0;
    VolumeRaycaster::deinitialize();
}

// same readiness rule as singlevolumeraycaster.cpp:187-202: all inports, at least one outport
bool B200SingleVolumeRaycaster::isReady() const {
    if (!ctx_ || !entryPort_.isReady() || !exitPort_.isReady() || !volumeInport_.isReady()) return false;
    return outport1_.isReady() || outport2_.isReady() || outport3_.isReady();
}

void B200SingleVolumeRaycaster::uploadVolume() {
    const VolumeBase* vol = volumeInport_.getData();
    if (vol == residentVolume_) return;
    const VolumeRAM* ram = vol->getRepresentation<VolumeRAM>();            // volumeram.h:76-95,249-250
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
    const RealWorldMapping rwm = vol->getRealWorldMapping();               // volumebase.cpp:158-231
    d.rwm_scale = rwm.getScale(); d.rwm_offset = rwm.getOffset();
    failOn(b200rc_upload_volume(ctx_, 0, ram->getData(), &d), "b200rc_upload_volume");
    residentVolume_ = vol;
}

// ClassificationModes::bindTexture + TransFunc1D::setUniform (classificationmodes.cpp:56-73, transfunc1d.cpp:184-194)
void B200SingleVolumeRaycaster::uploadTransferFunctions(float stepSize) {
    const int nch = (int)volumeInport_.getData()->getNumChannels();
    const std::string cls = classificationMode_.get();
    for (int c = 0; c < nch && c < kChannels; ++c) {
        TransFunc1D* tf = 0;
        switch (transFuncTypeProp_.getValue()) {
            case TFT_1DKEYS:     tf = tf1d_[c]->get(); break;
            case TFT_1DGAUSSIAN: tf = tfGauss_[c]->get(); break;
            default: break;
            case TFT_2DPRIMITIVES: {
                if (c > 0) continue;                                        // the GLSL path binds channel 1 only (singlevolumeraycaster.cpp:376-386)
                TransFunc2DPrimitives* t2 = tf2d_[0]->get();
                tgt::Texture* tex = t2->getTexture();
                tex->downloadTexture();
                const tgt::vec2 dx = t2->getDomain(0), dy = t2->getDomain(1);
                failOn(b200rc_set_tf2d(ctx_, 0, tex->getCpuTextureData(), tex->getWidth(), tex->getHeight(), dx.x, dx.y, dy.x, dy.y), "b200rc_set_tf2d");
                continue;
            }
        }
        if (!tf) throw VoreenException("b200raycast: no transfer function");
        tgt::Texture* tex = tf->getTexture();                               // RGBA8, width 2^bits <= 4096 (transfuncpropertybase.cpp:155-165)
        if (!tex->getCpuTextureData()) tex->downloadTexture();
        const tgt::vec2 dom = tf->getDomain();
        failOn(b200rc_set_tf1d(ctx_, c, tex->getCpuTextureData(), tex->getWidth(), dom.x, dom.y), "b200rc_set_tf1d");
        if (cls == "pre-integrated-gpu") {
            // preintegration.frag:35-87 re-done on the device from the TF texels (RGBA16 quantisation included)
            failOn(b200rc_build_preint(ctx_, c, std::min(tex->getWidth(), 1024), stepSize, B200RC_PI_GPU), "b200rc_build_preint");
        } else if (cls == "pre-integrated-cpu" || cls == "pre-integrated-approximation") {
            // Voreen's own table (preintegrationtable.cpp:67-204), handed over as is: bit-identical classification.  The TF
            // caches its tables per (step, dimension, mode) and drops them when its keys change (preintegrationtablemap.cpp:59-97),
            // so an unchanged pointer is an unchanged table; the library compares content hashes on top of that
            const PreIntegrationTable* t = tf->getPreIntegrationTable(stepSize, 0, cls == "pre-integrated-approximation", false);
            if (residentPreint_[c] != t->getTable()) {
                failOn(b200rc_set_preint(ctx_, c, reinterpret_cast<const float*>(t->getTable()), (int)t->getDimension()), "b200rc_set_preint");
                residentPreint_[c] = t->getTable();
            }
        }
    }
}

void B200SingleVolumeRaycaster::fillParams(b200rc_params& p, b200rc_camera& cam, const tgt::ivec2& size, float stepSize) {
    std::memset(&p, 0, sizeof(p));
    p.num_volumes = 1;
    b200rc_volume_params& v = p.vol[0];
    const std::string cls = classificationMode_.get();
    v.classification = (transFuncTypeProp_.getValue() == TFT_2DPRIMITIVES) ? B200RC_CLASS_TF2D
                     : (cls == "transfer-function" ? B200RC_CLASS_TF : B200RC_CLASS_PREINT);
    v.shading = lookup(kShading, shadeMode_.get(), "shading");
    // VolumePort texture access properties (volumeport.cpp:262-287); GL enums -> ABI enums
    v.filter = volumeInport_.getTextureFilterModeProperty().getValue() == GL_NEAREST ? B200RC_FILTER_NEAREST : B200RC_FILTER_LINEAR;
    v.clamp = volumeInport_.getTextureClampModeProperty().getValue() == GL_CLAMP_TO_BORDER ? B200RC_CLAMP_TO_BORDER : B200RC_CLAMP_TO_EDGE;
    v.border_intensity = volumeInport_.getTextureBorderIntensityProperty().get();
    p.sampling_step = stepSize;
    p.gradient = lookup(kGradient, gradientMode_.get(), "gradient");
    const StringOptionProperty* comp[3] = {&compositingMode1_, &compositingMode2_, &compositingMode3_};
    const RenderPort* out[3] = {&outport1_, &outport2_, &outport3_};
    for (int i = 0; i < 3; ++i) {
        p.compositing[i] = lookup(kCompositing, comp[i]->get(), "compositing");
        p.output_connected[i] = out[i]->isConnected() ? 1 : 0;
    }
    p.iso_value = isoValue_.get();
    // the reference's uniform wiring, quirk included (singlevolumeraycaster.cpp:393-400): gammaValue2_ follows compositing1
    p.mida_gamma[0] = compositingMode1_.isSelected("mida") ? gammaValue1_.get() : 0.f;
    p.mida_gamma[1] = compositingMode1_.isSelected("mida") ? gammaValue2_.get() : 0.f;
    p.mida_gamma[2] = compositingMode2_.isSelected("mida") ? gammaValue3_.get() : 0.f;
    p.enable_channel_shift = enableChannelShift_.get() ? 1 : 0;
    const tgt::vec3 dim(volumeInport_.getData()->getDimensions());
    for (int c = 0; c < kChannels; ++c) {
        const tgt::vec3 s = channelShift_[c]->get() / dim;                  // singlevolumeraycaster.cpp:270-279
        p.channel_shift[c][0] = s.x; p.channel_shift[c][1] = s.y; p.channel_shift[c][2] = s.z;
    }
    // VolumeRenderer lighting (volumerenderer.cpp:41-106)
    const tgt::vec4 lp = lightPosition_.get();
    for (int i = 0; i < 4; ++i) p.light.position[i] = lp[i];
    for (int i = 0; i < 3; ++i) {
        p.light.ambient[i] = lightAmbient_.get()[i]; p.light.diffuse[i] = lightDiffuse_.get()[i];
        p.light.specular[i] = lightSpecular_.get()[i]; p.light.attenuation[i] = lightAttenuation_.get()[i];
    }
    p.light.shininess = materialShininess_.get();
    p.light.apply_attenuation = applyLightAttenuation_.get() ? 1 : 0;
    p.sampler = B200RC_SAMPLER_EXACT;
    p.empty_space_skipping = 2;                                             // TF-transparent cells, and for MIP cells that cannot beat the running maximum
    p.out_format = B200RC_OUT_RGBA16;                                        // RenderPort default (renderport.h:116)
    p.quantize_depth24 = 1;

    const tgt::Camera c = camera_.get();
    copyMat(c.getViewMatrix(), cam.view);
    copyMat(c.getProjectionMatrix(size), cam.proj);
    const tgt::vec3 pos = c.getPositionWithOffsets();
    cam.position[0] = pos.x; cam.position[1] = pos.y; cam.position[2] = pos.z;
    cam.near_dist = c.getNearDist(); cam.far_dist = c.getFarDist();
}

void B200SingleVolumeRaycaster::beforeProcess() {
    VolumeRaycaster::beforeProcess();
    // INVALID_PROGRAM in the GLSL path = recompile; here the kernels are pre-built.  A changed volume port -- whatever the
    // invalidation level it arrived with -- means the voxels in HBM are stale
    if (volumeInport_.hasChanged()) residentVolume_ = 0;
    if (getInvalidationLevel() >= Processor::INVALID_PROGRAM)
        for (int c = 0; c < kChannels; ++c) residentPreint_[c] = 0;
}

void B200SingleVolumeRaycaster::writeOutport(RenderPort& port, const uint16_t* rgba16, const float* depth, const tgt::ivec2& size) {
    // non-sharing upload pattern of V/modules/opencl/processors/singleoctreeraycastercl.cpp:1765-1768; the port's render
    // target is RGBA16 + DEPTH24, the same formats the GLSL path writes (rescaleRendering is the identity at equal size).
    // `size` is the port's own size here (process() rescales coarse renderings first), so the render target keeps its size
    tgt::Texture* col = port.getColorTexture();
    col->bind();
    glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, size.x, size.y, GL_RGBA, GL_UNSIGNED_SHORT, rgba16);
    if (depth) {
        tgt::Texture* dt = port.getDepthTexture();
        dt->bind();
        glTexSubImage2D(GL_TEXTURE_2D, 0, 0, 0, size.x, size.y, GL_DEPTH_COMPONENT, GL_FLOAT, depth);
    }
    port.validateResult();
    LGL_ERROR;
}

namespace {
// glGetTexImage straight into a page-locked buffer (what Texture::downloadTextureToBuffer does into heap memory,
// V/ext/tgt/texture.cpp; pattern cpuraycaster.cpp:166-169)
void downloadInto(tgt::Texture* tex, GLenum format, void* dst) {
    tex->bind();
    glGetTexImage(GL_TEXTURE_2D, 0, format, GL_FLOAT, dst);
}
}

void B200SingleVolumeRaycaster::process() {
    const bool coarse = QualityMode.isInteractionMode() && interactionCoarseness_.get() > 1;     // singlevolumeraycaster.cpp:253-255
    const tgt::ivec2 size = coarse ? outport1_.getSize() / interactionCoarseness_.get() : outport1_.getSize();
    const float stepSize = getSamplingStepSize(volumeInport_.getData());                            // volumeraycaster.cpp:312-322

    uploadVolume();                         // no-op unless the volume port changed
    uploadTransferFunctions(stepSize);      // hands the TFs over; the library skips whatever is unchanged (no per-frame allocation)
    b200rc_params params; b200rc_camera cam;
    fillParams(params, cam, size, stepSize);

    // entry/exit RenderPorts -> page-locked host buffers; the entry/exit images have the size of the internal render targets
    // because MeshEntryExitPoints follows the same coarseness rule
    const size_t px = (size_t)size.x * size.y;
    float* entry = static_cast<float*>(pinEntry_.ensure(px * 16));
    float* exitp = static_cast<float*>(pinExit_.ensure(px * 16));
    float* entryD = static_cast<float*>(pinEntryDepth_.ensure(px * 4));
    float* exitD = static_cast<float*>(pinExitDepth_.ensure(px * 4));
    float* depth = static_cast<float*>(pinDepth_.ensure(px * 4));
    if (!entry || !exitp || !entryD || !exitD || !depth) throw VoreenException("b200raycast: out of page-locked host memory");
    downloadInto(entryPort_.getColorTexture(), GL_RGBA, entry);
    downloadInto(exitPort_.getColorTexture(), GL_RGBA, exitp);
    downloadInto(entryPort_.getDepthTexture(), GL_DEPTH_COMPONENT, entryD);
    downloadInto(exitPort_.getDepthTexture(), GL_DEPTH_COMPONENT, exitD);

    void* outColor[B200RC_MAX_OUTPUTS] = {0, 0, 0};
    for (int i = 0; i < B200RC_MAX_OUTPUTS; ++i)
        if (params.output_connected[i]) {
            outColor[i] = pinColor_[i].ensure(px * 8);
            if (!outColor[i]) throw VoreenException("b200raycast: out of page-locked host memory");
        }

    failOn(b200rc_render(ctx_, &params, &cam, size.x, size.y, entry, entryD, exitp, exitD, outColor, depth), "b200rc_render");

    // rescaleRendering (volumeraycaster.cpp:348-377): internal size -> outport size with GL_LINEAR; the identity at equal size
    RenderPort* out[3] = {&outport1_, &outport2_, &outport3_};
    for (int i = 0; i < 3; ++i) {
        if (!params.output_connected[i]) continue;
        const tgt::ivec2 outSize = out[i]->getSize();
        if (outSize == size) { writeOutport(*out[i], static_cast<const uint16_t*>(outColor[i]), i == 0 ? depth : 0, size); continue; }
        const size_t opx = (size_t)outSize.x * outSize.y;
        uint16_t* up = static_cast<uint16_t*>(pinScaled_.ensure(opx * 8));
        float* upDepth = i == 0 ? static_cast<float*>(pinScaledDepth_.ensure(opx * 4)) : 0;
        if (!up || (i == 0 && !upDepth)) throw VoreenException("b200raycast: out of page-locked host memory");
        failOn(b200rc_rescale(ctx_, outColor[i], i == 0 ? depth : 0, size.x, size.y, B200RC_OUT_RGBA16,
                              up, upDepth, outSize.x, outSize.y, 0), "b200rc_rescale");
        writeOutport(*out[i], up, upDepth, outSize);
    }
    tgt::TextureUnit::setZeroUnit();
    LGL_ERROR;
}

} // namespace
