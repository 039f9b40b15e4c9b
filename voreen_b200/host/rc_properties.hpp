// rc_properties.hpp -- property IDs / option strings of the reference's raycaster processors and their translation
// into the POD parameter blocks of include/b200rc.h.  Header-only, plain C++11, no Voreen and no CUDA dependency:
// the SAME code is used by the standalone host mirror (b200host.cpp) and by the Voreen drop-in processor
// (custommodules/b200raycast/processors/*.cpp), so both interpret a workspace identically.
//
// Mirrors: VolumeRaycaster ctor/properties (V/src/core/processors/volumeraycaster.cpp:42-157,312-322,379-388),
// VolumeRenderer light/material (V/src/core/processors/volumerenderer.cpp:41-71), SingleVolumeRaycaster
// (V/modules/base/processors/render/singlevolumeraycaster.cpp:46-170,267-303,388-400,477-529), ClassificationModes
// (V/src/core/utils/classificationmodes.cpp:36-90), ShadeMode strings (V/src/core/utils/glsl.cpp:397-409), VolumePort
// texture access properties (V/src/core/ports/volumeport.cpp:262-287), MeshEntryExitPoints
// (V/modules/base/processors/entryexitpoints/meshentryexitpoints.cpp:47-94).   V/ = /root/reference/voreen/
#pragma once
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include "../../include/b200rc.h"

namespace b200rc {

inline int lookup(const char* const* names, int n, const std::string& s) {
    for (int i = 0; i < n; ++i)
        if (s == names[i]) return i;
    return -1;
}
inline int gradient_from_string(const std::string& s) {
    static const char* const k[] = {"none", "forward-differences", "central-differences", "sobel", "filtered"};
    return lookup(k, 5, s);
}
inline int shading_from_string(const std::string& s) {
    static const char* const k[] = {"none", "phong-diffuse", "phong-specular", "phong-diffuse-ambient", "phong-diffuse-specular",
                                    "phong", "toon", "cook-torrance", "oren-nayar", "ward"};
    return lookup(k, 10, s);
}
inline int compositing_from_string(const std::string& s) {
    static const char* const k[] = {"dvr", "mip", "mop", "mida", "iso", "fhp", "fhn"};
    return lookup(k, 7, s);
}
// classification option -> (marcher mode, table builder); -1 if unknown
inline int classification_from_string(const std::string& s, int* builder) {
    if (s == "transfer-function") { *builder = -1; return B200RC_CLASS_TF; }
    if (s == "pre-integrated-gpu") { *builder = B200RC_PI_GPU; return B200RC_CLASS_PREINT; }
    if (s == "pre-integrated-cpu") { *builder = B200RC_PI_CPU; return B200RC_CLASS_PREINT; }
    if (s == "pre-integrated-approximation") { *builder = B200RC_PI_APPROX; return B200RC_CLASS_PREINT; }
    return -1;
}

struct RaycasterProperties {
    // VolumeRaycaster
    float samplingRate = 2.f;               // volumeraycaster.cpp:56
    float interactionQuality = 1.f;         // :66
    bool interactionMode = false;           // QualityMode.isInteractionMode()
    float isoValue = 0.5f;                  // :57
    float gammaValue[3] = {0.f, 0.f, 0.f};  // gammaValue, gammaValue1, gammaValue2 (:67-69)
    std::string gradient = "central-differences";
    std::string classification = "transfer-function";
    std::string shading = "phong";
    std::string compositing[3] = {"dvr", "dvr", "dvr"};     // compositing, compositing2, compositing3 (:62-64)
    bool outputConnected[3] = {true, false, false};
    // VolumeRenderer
    float lightPosition[4] = {2.3f, 1.5f, 1.5f, 1.f};
    float lightAmbient[3] = {0.4f, 0.4f, 0.4f}, lightDiffuse[3] = {0.8f, 0.8f, 0.8f}, lightSpecular[3] = {0.6f, 0.6f, 0.6f};
    float lightAttenuation[3] = {1.f, 0.f, 0.f};
    bool applyLightAttenuation = false;
    float materialShininess = 60.f;
    // SingleVolumeRaycaster
    bool enableChannelShift = false;
    float channelShift[4][3] = {{0, 0, 0}, {0, 0, 0}, {0, 0, 0}, {0, 0, 0}};    // voxels
    // VolumePort
    std::string textureFilterMode = "linear", textureClampMode = "clamp-to-edge";
    float textureBorderIntensity = 0.f;
    // MeshEntryExitPoints
    std::string cameraInsideVolumeTechnique = "gpu";
    bool jitterEntryPoints = false;
    float jitterStepLength = 0.005f;
    bool useFloatRenderTargets = false;
    // implementation knobs (no reference counterpart)
    bool emptySpaceSkipping = true;
    std::string sampler = "exact";
    std::string outputFormat = "rgba32f";
    bool collectStats = false;

    static bool parse_bool(const std::string& v, bool* out) {
        if (v == "true" || v == "1") { *out = true; return true; }
        if (v == "false" || v == "0") { *out = false; return true; }
        return false;
    }
    static bool parse_floats(const std::string& v, float* out, int n) {
        const char* s = v.c_str();
        for (int i = 0; i < n; ++i) {
            char* end = nullptr;
            const float f = std::strtof(s, &end);
            if (end == s) return false;
            out[i] = f; s = end;
        }
        return true;
    }
    // returns false for an unknown id or an invalid option (the reference: "property not found" / option rejected)
    bool set(const std::string& id, const std::string& v) {
        if (id == "samplingRate") return parse_floats(v, &samplingRate, 1) && samplingRate >= 0.01f && samplingRate <= 20.f;
        if (id == "interactionQuality") return parse_floats(v, &interactionQuality, 1) && interactionQuality >= 0.01f && interactionQuality <= 1.f;
        if (id == "interactionMode") return parse_bool(v, &interactionMode);
        if (id == "isoValue") return parse_floats(v, &isoValue, 1);
        if (id == "gammaValue") return parse_floats(v, &gammaValue[0], 1);
        if (id == "gammaValue1") return parse_floats(v, &gammaValue[1], 1);
        if (id == "gammaValue2") return parse_floats(v, &gammaValue[2], 1);
        if (id == "gradient") { if (gradient_from_string(v) < 0) return false; gradient = v; return true; }
        if (id == "classification") { int b; if (classification_from_string(v, &b) < 0) return false; classification = v; return true; }
        if (id == "shading") { if (shading_from_string(v) < 0) return false; shading = v; return true; }
        if (id == "compositing" || id == "compositing2" || id == "compositing3") {       // volumeraycaster.cpp:62-64
            if (compositing_from_string(v) < 0) return false;
            compositing[id == "compositing" ? 0 : (id == "compositing2" ? 1 : 2)] = v; return true;
        }
        if (id == "lightPosition") return parse_floats(v, lightPosition, 4);
        if (id == "lightAmbient") return parse_floats(v, lightAmbient, 3);
        if (id == "lightDiffuse") return parse_floats(v, lightDiffuse, 3);
        if (id == "lightSpecular") return parse_floats(v, lightSpecular, 3);
        if (id == "lightAttenuation") return parse_floats(v, lightAttenuation, 3);
        if (id == "applyLightAttenuation") return parse_bool(v, &applyLightAttenuation);
        if (id == "materialShininess") return parse_floats(v, &materialShininess, 1);
        if (id == "enableChannelShift") return parse_bool(v, &enableChannelShift);
        if (id == "channelShift1" || id == "channelShift2" || id == "channelShift3" || id == "channelShift4")
            return parse_floats(v, channelShift[id[12] - '1'], 3);
        if (id == "textureFilterMode") { if (v != "nearest" && v != "linear") return false; textureFilterMode = v; return true; }
        if (id == "textureClampMode") { if (v != "clamp-to-edge" && v != "clamp-to-border") return false; textureClampMode = v; return true; }
        if (id == "textureBorderIntensity") return parse_floats(v, &textureBorderIntensity, 1);
        if (id == "cameraInsideVolumeTechnique") { if (v != "none" && v != "gpu" && v != "cpu") return false; cameraInsideVolumeTechnique = v; return true; }
        if (id == "jitterEntryPoints") return parse_bool(v, &jitterEntryPoints);
        if (id == "jitterStepLength") return parse_floats(v, &jitterStepLength, 1);
        if (id == "useFloatRenderTargets") return parse_bool(v, &useFloatRenderTargets);
        if (id == "emptySpaceSkipping") return parse_bool(v, &emptySpaceSkipping);
        if (id == "sampler") { if (v != "exact" && v != "texture") return false; sampler = v; return true; }
        if (id == "outputFormat") { if (v != "rgba16" && v != "rgba32f") return false; outputFormat = v; return true; }
        if (id == "collectStats") return parse_bool(v, &collectStats);
        if (id == "raycast.prg") return true;       // ShaderProperty kept as an inert property so workspaces deserialise
        return false;
    }
};

// VolumeRaycaster::getSamplingStepSize (volumeraycaster.cpp:312-322)
inline float sampling_step_size(const int dims[3], const RaycasterProperties& pr) {
    const int m = dims[0] > dims[1] ? (dims[0] > dims[2] ? dims[0] : dims[2]) : (dims[1] > dims[2] ? dims[1] : dims[2]);
    float step = 1.f / ((float)m * pr.samplingRate);
    if (pr.interactionMode) step /= pr.interactionQuality;
    return step;
}

// SingleVolumeRaycaster::adjustPropertyVisibilities (singlevolumeraycaster.cpp:516-528): pre-integration is switched
// back to "transfer-function" with MIP on any output (and with 2D TFs); returns the effective classification string.
inline std::string effective_classification(const RaycasterProperties& pr, bool tf2d) {
    if (pr.classification != "transfer-function") {
        if (tf2d) return "transfer-function";
        if (pr.compositing[0] == "mip" || pr.compositing[1] == "mip" || pr.compositing[2] == "mip") return "transfer-function";
    }
    return pr.classification;
}

// generateHeader + process(): properties -> b200rc_params (single-volume semantics).  dims = volume dimensions.
inline bool to_params(const RaycasterProperties& pr, const int dims[3], bool tf2d, b200rc_params* P, int* preint_builder, std::string* err) {
    std::memset(P, 0, sizeof *P);
    P->num_volumes = 1; P->multi_volume = 0;
    int builder = -1;
    const int cls = tf2d ? B200RC_CLASS_TF2D : classification_from_string(effective_classification(pr, tf2d), &builder);
    const int shade = shading_from_string(pr.shading), grad = gradient_from_string(pr.gradient);
    if (cls < 0 || shade < 0 || grad < 0) { if (err) *err = "unknown classification/shading/gradient option"; return false; }
    if (preint_builder) *preint_builder = builder;
    for (int v = 0; v < B200RC_MAX_VOLUMES; ++v) {
        P->vol[v].classification = cls; P->vol[v].shading = shade;
        P->vol[v].filter = pr.textureFilterMode == "nearest" ? B200RC_FILTER_NEAREST : B200RC_FILTER_LINEAR;
        P->vol[v].clamp = pr.textureClampMode == "clamp-to-border" ? B200RC_CLAMP_TO_BORDER : B200RC_CLAMP_TO_EDGE;
        P->vol[v].border_intensity = pr.textureBorderIntensity;
    }
    P->sampling_step = sampling_step_size(dims, pr);
    P->gradient = grad;
    for (int o = 0; o < 3; ++o) {
        P->compositing[o] = compositing_from_string(pr.compositing[o]);
        if (P->compositing[o] < 0) { if (err) *err = "unknown compositing option"; return false; }
        P->output_connected[o] = pr.outputConnected[o] ? 1 : 0;
    }
    P->iso_value = pr.isoValue;
    // quirk (singlevolumeraycaster.cpp:393-400): the uniform gammaValue2_ is set when *compositing* (output 1) is mida and
    // gammaValue3_ when compositing2 (output 2) is; an unset uniform is 0
    P->mida_gamma[0] = pr.compositing[0] == "mida" ? pr.gammaValue[0] : 0.f;
    P->mida_gamma[1] = pr.compositing[0] == "mida" ? pr.gammaValue[1] : 0.f;
    P->mida_gamma[2] = pr.compositing[1] == "mida" ? pr.gammaValue[2] : 0.f;
    P->enable_channel_shift = pr.enableChannelShift ? 1 : 0;
    for (int c = 0; c < 4; ++c)
        for (int a = 0; a < 3; ++a) P->channel_shift[c][a] = pr.channelShift[c][a] / (float)dims[a];   // :270-279
    for (int i = 0; i < 4; ++i) P->light.position[i] = pr.lightPosition[i];
    for (int i = 0; i < 3; ++i) {
        P->light.ambient[i] = pr.lightAmbient[i]; P->light.diffuse[i] = pr.lightDiffuse[i];
        P->light.specular[i] = pr.lightSpecular[i]; P->light.attenuation[i] = pr.lightAttenuation[i];
    }
    P->light.shininess = pr.materialShininess;
    P->light.apply_attenuation = pr.applyLightAttenuation ? 1 : 0;
    P->sampler = pr.sampler == "texture" ? B200RC_SAMPLER_TEX : B200RC_SAMPLER_EXACT;
    P->empty_space_skipping = pr.emptySpaceSkipping ? 1 : 0;
    P->out_format = pr.outputFormat == "rgba16" ? B200RC_OUT_RGBA16 : B200RC_OUT_RGBA32F;
    P->quantize_depth24 = 1;
    P->collect_stats = pr.collectStats ? 1 : 0;
    return true;
}

inline void to_eep_params(const RaycasterProperties& pr, b200rc_eep_params* E) {
    std::memset(E, 0, sizeof *E);
    E->proxy_mode = B200RC_PROXY_BBOX; E->proxy_cell_size = 16; E->proxy_threshold = 1.f;
    E->inside_technique = pr.cameraInsideVolumeTechnique == "none" ? B200RC_INSIDE_NONE : B200RC_INSIDE_GPU;
    E->jitter = pr.jitterEntryPoints ? 1 : 0; E->jitter_step_length = pr.jitterStepLength;
    E->world_coords = 0;
}

}  // namespace b200rc
