// b200host.cpp -- C++ host-side mirror of the reference's processor interface for the ray-casting path
// (include/b200rc_host.h).  Plays Voreen's role outside Voreen: camera, transfer-function texture generation, property
// handling, invalidation (what has to be re-uploaded / rebuilt before the next process()), then calls the C ABI of
// include/b200rc.h.  No rendering arithmetic lives here -- every pixel comes out of libb200rc.so (sm_100a CUDA).
//
// V/ = /root/reference/voreen/
#include "../../include/b200rc_host.h"
#include "rc_properties.hpp"

#include <cctype>
#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <utility>
#include <cstdio>
#include <cstring>
#include <new>
#include <limits>
#include <string>
#include <vector>

namespace {

// ---- tgt::mat4 / vec3 arithmetic the camera needs (row-major, column vectors; V/ext/tgt/matrix.h, vector.h) ----
struct V3 { float x, y, z; };
inline V3 sub(V3 a, V3 b) { return {a.x - b.x, a.y - b.y, a.z - b.z}; }
inline V3 cross(V3 a, V3 b) { return {a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x}; }
// tgt::normalize multiplies by the reciprocal length; lengthSq accumulates from 0 (vector.h:802-820)
inline V3 normalize(V3 v) {
    float l2 = 0.f; l2 += v.x * v.x; l2 += v.y * v.y; l2 += v.z * v.z;
    const float r = 1.f / std::sqrt(l2);
    return {v.x * r, v.y * r, v.z * r};
}
void mat_mul(const float a[16], const float b[16], float out[16]) {
    float r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            r[i * 4 + j] = s;
        }
    std::memcpy(out, r, sizeof r);
}
// mat4::createLookAt (matrix.h:1196-1210): rotation rows (strafe, up, -look) times translate(-eye)
void look_at(V3 eye, V3 focus, V3 up, float out[16]) {
    const V3 look = normalize(sub(focus, eye));
    const V3 strafe = normalize(cross(look, normalize(up)));
    const V3 up2 = cross(strafe, look);
    const float rot[16] = {strafe.x, strafe.y, strafe.z, 0.f, up2.x, up2.y, up2.z, 0.f, -look.x, -look.y, -look.z, 0.f, 0.f, 0.f, 0.f, 1.f};
    const float tr[16] = {1.f, 0.f, 0.f, -eye.x, 0.f, 1.f, 0.f, -eye.y, 0.f, 0.f, 1.f, -eye.z, 0.f, 0.f, 0.f, 1.f};
    mat_mul(rot, tr, out);
}
// mat4::createPerspective (matrix.h:1230-1248)
void perspective(float fov_rad, float aspect, float n, float f, float out[16]) {
    // tgt's `tan` is the C double overload: T(1) / tan(fov / T(2)) is evaluated in double and rounded once (pinned against the
    // reference header by tests/test_oracle_pins.py)
    const float c = (float)(1.0 / ::tan((double)(fov_rad / 2.f)));
    const float m[16] = {c / aspect, 0.f, 0.f, 0.f, 0.f, c, 0.f, 0.f, 0.f, 0.f, (n + f) / (n - f), (2.f * f * n) / (n - f), 0.f, 0.f, -1.f, 0.f};
    std::memcpy(out, m, sizeof m);
}
// Camera::getProjectionMatrix, PERSPECTIVE (camera.cpp:188-217)
void camera_projection(float fovy_deg, float frustum_ratio, float n, float f, float window_ratio, float out[16]) {
    const float PIf = 3.14159265358979323846f;
    float fovy = fovy_deg < 6.f ? 6.f : (fovy_deg > 175.f ? 175.f : fovy_deg);
    if (window_ratio >= 1.0f) perspective((PIf / 180.f) * fovy, frustum_ratio * window_ratio, n, f, out);
    else perspective((float)(::atan(::tan((double)((PIf / 180.f) * (fovy / 2.f))) / (double)(window_ratio * frustum_ratio)) * 2), frustum_ratio * window_ratio, n, f, out);
}

inline int iround(float f) { return (int)std::floor(f + 0.5f); }     // tgt::iround (tgt_math.h:213-225)

// TransFunc1DKeys::getMappingForValueUByte (transfunc1dkeys.cpp:119-154): the interpolation happens in GLubyte += float,
// i.e. a float sum converted back with truncation
void keys_mapping(const b200host_tf_key* keys, int n, float value, uint8_t out[4]) {
    if (n == 0) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    value = value < 0.f ? 0.f : (value > 1.f ? 1.f : value);
    int it = 0;
    while (it != n && value > keys[it].intensity) ++it;
    if (it == 0) { std::memcpy(out, keys[0].color_l, 4); return; }
    if (it == n) { std::memcpy(out, keys[n - 1].color_r, 4); return; }
    const b200host_tf_key& l = keys[it - 1]; const b200host_tf_key& r = keys[it];
    const float fraction = (value - l.intensity) / (r.intensity - l.intensity);
    for (int c = 0; c < 4; ++c) {
        const float sum = (float)l.color_r[c] + (float)((int)r.color_l[c] - (int)l.color_r[c]) * fraction;
        out[c] = (uint8_t)(int)sum;
    }
}

struct TransFunc {
    std::vector<b200host_tf_key> keys;
    float domain[2] = {0.f, 1.f}, threshold[2] = {0.f, 1.f}, gamma = 1.f;
    bool set = false, dirty = true;
};

}  // namespace

// ---- flat C API -------------------------------------------------------------------------------------------------------
extern "C" int b200host_camera(const float position[3], const float focus[3], const float up[3], float fovy_deg, float frustum_ratio,
                               float near_dist, float far_dist, int width, int height, b200rc_camera* out) {
    if (!position || !focus || !up || !out || width < 1 || height < 1) return B200RC_ERR_INVALID;
    look_at({position[0], position[1], position[2]}, {focus[0], focus[1], focus[2]}, {up[0], up[1], up[2]}, out->view);
    camera_projection(fovy_deg, frustum_ratio, near_dist, far_dist, (float)width / (float)height, out->proj);
    for (int i = 0; i < 3; ++i) out->position[i] = position[i];
    out->near_dist = near_dist; out->far_dist = far_dist;
    return B200RC_OK;
}

extern "C" int b200host_tf1d_keys_texture(const b200host_tf_key* keys, int n_keys, const float threshold[2], float gamma, int alpha_mode,
                                          int width, uint8_t* out) {
    if ((!keys && n_keys > 0) || n_keys < 0 || width < 2 || !out) return B200RC_ERR_INVALID;
    const float thr0 = threshold ? threshold[0] : 0.f, thr1 = threshold ? threshold[1] : 1.f;
    // TransFunc1D::createTextureData (transfunc1d.cpp:85-116); quirk kept: threshold compared against round(thr * width)
    const int front_end = iround(thr0 * (float)width), back_start = iround(thr1 * (float)width);
    for (int x = 0; x < width; ++x) {
        uint8_t* px = out + 4 * x;
        if (x < front_end || x > back_start) px[0] = px[1] = px[2] = px[3] = 0;
        else {
            float v = (float)x / (float)(width - 1);
            if (gamma != 1.f) v = std::pow(v, gamma);            // applyGammaToIntensity (:144-149)
            keys_mapping(keys, n_keys, v, px);
        }
        if (alpha_mode == 2) { if (px[3]) px[3] = 255; }        // TF_ONE_ALPHA (transfuncbase.cpp:177-210)
        else if (alpha_mode == 0) px[3] = 0;                    // TF_ZERO_ALPHA
    }
    return B200RC_OK;
}

// TransFuncMappingCurve::getOpacityAt / getColorAt (transfuncmappingcurve.cpp:257-288) and
// TransFunc1DGaussian::getMappingForValueUByte (transfunc1dgaussian.cpp:122-178)
static float curve_opacity(const b200host_tf_curve& c, float x) {
    // ::exp(double) is what the reference's unqualified exp() resolves to: product and sum in double, returned as float
    const double e = ::exp(static_cast<double>(-(x - c.intensity) * (x - c.intensity) / (2 * c.width)));
    const float a = (x < c.intensity ? c.color_l[3] : c.color_r[3]) / 255.f;
    return static_cast<float>(e * (a - c.base_value) + c.base_value);
}
static void gaussian_mapping(const b200host_tf_curve* curves, int n, float x, uint8_t out[4]) {
    out[0] = out[1] = out[2] = out[3] = 0;
    if (n <= 0) return;
    float col[4] = {0.f, 0.f, 0.f, 0.f}, alphaSum = 0.f;
    for (int i = 0; i < n; ++i) {
        const b200host_tf_curve& c = curves[i];
        if (!c.active) continue;
        const float yRatio = static_cast<float>(::exp(static_cast<double>(-(x - c.intensity) * (x - c.intensity) / (2 * c.width))));
        const uint8_t* peak = x < c.intensity ? c.color_l : c.color_r;
        const uint8_t* base = x < c.intensity ? c.base_color_l : c.base_color_r;
        uint8_t cc[4];
        for (int k = 0; k < 3; ++k) cc[k] = static_cast<uint8_t>((peak[k] * yRatio + base[k] * (1.f - yRatio)));
        cc[3] = static_cast<uint8_t>(curve_opacity(c, x) * 255);
        float alpha = curve_opacity(c, x);
        if (alpha < std::numeric_limits<float>::min()) alpha = std::numeric_limits<float>::min();
        alphaSum += alpha;
        col[0] += (cc[0] * alpha); col[1] += (cc[1] * alpha); col[2] += (cc[2] * alpha);
        col[3] = (cc[3] > col[3] ? cc[3] : col[3]);
    }
    if (alphaSum > 1.f) alphaSum = 1.f;
    for (int k = 0; k < 3; ++k) col[k] = col[k] / alphaSum;
    for (int k = 0; k < 4; ++k) { float v = col[k]; v = v < 0.f ? 0.f : (v > 255.f ? 255.f : v); out[k] = static_cast<uint8_t>(v); }     // tgt::clamp; NaN (0/0) -> 0 as min/max do
}

extern "C" int b200host_tf1d_gaussian_texture(const b200host_tf_curve* curves, int n_curves, const float threshold[2], float gamma, int alpha_mode,
                                              int width, uint8_t* out) {
    if ((!curves && n_curves > 0) || n_curves < 0 || width < 2 || !out) return B200RC_ERR_INVALID;
    const float thr0 = threshold ? threshold[0] : 0.f, thr1 = threshold ? threshold[1] : 1.f;
    const int front_end = iround(thr0 * (float)width), back_start = iround(thr1 * (float)width);      // TransFunc1D::createTextureData (transfunc1d.cpp:85-116)
    for (int x = 0; x < width; ++x) {
        uint8_t* px = out + 4 * x;
        if (x < front_end || x > back_start) px[0] = px[1] = px[2] = px[3] = 0;
        else {
            float v = (float)x / (float)(width - 1);
            if (gamma != 1.f) v = std::pow(v, gamma);
            gaussian_mapping(curves, n_curves, v, px);
        }
        if (alpha_mode == 2) { if (px[3]) px[3] = 255; }
        else if (alpha_mode == 0) px[3] = 0;
    }
    return B200RC_OK;
}

extern "C" int b200host_tf1d_width_for(int dtype, int* out_width) {
    if (!out_width) return B200RC_ERR_INVALID;
    // bits per channel, float handled as 16 bit, 1D TFs limited to 12 bit (transfuncpropertybase.cpp:155-165,
    // transfunc1dkeysproperty.cpp:94-97)
    *out_width = dtype == B200RC_U8 ? 256 : 4096;
    return B200RC_OK;
}

extern "C" int b200host_volume_desc(const int32_t dims[3], int channels, const char* format, const float* spacing, const float* offset,
                                    const float* phys_to_world, b200rc_volume_desc* out) {
    if (!dims || !format || !out || channels < 1 || channels > 4) return B200RC_ERR_INVALID;
    std::memset(out, 0, sizeof *out);
    const std::string f(format);
    if (f == "uint8") { out->dtype = B200RC_U8; out->rwm_scale = 255.f; }
    else if (f == "uint16") { out->dtype = B200RC_U16; out->rwm_scale = 65535.f; }
    else if (f == "float") { out->dtype = B200RC_F32; out->rwm_scale = 1.f; }
    else return B200RC_ERR_INVALID;
    out->rwm_offset = 0.f;
    out->channels = channels;
    int m = 1;
    for (int a = 0; a < 3; ++a) { out->dims[a] = dims[a]; if (dims[a] > m) m = dims[a]; }
    for (int a = 0; a < 3; ++a) {
        out->spacing[a] = spacing ? spacing[a] : 2.0f / (float)m;
        out->offset[a] = offset ? offset[a] : (float)(-(dims[a] - 1) * 0.5f) * out->spacing[a];
    }
    for (int i = 0; i < 16; ++i) out->phys_to_world[i] = phys_to_world ? phys_to_world[i] : ((i % 5 == 0) ? 1.f : 0.f);
    return B200RC_OK;
}

// ---- OptimizedProxyGeometry: minboundingbox / visiblebricks ---------------------------------------------------------------------
namespace {
// alpha channel of PreIntegrationTable(tf, resolution 256, samplingStepSize 1.0).computeTable() (preintegrationtable.cpp:67-133);
// table[sb * res + sf] composites the TF entries from sf towards sb
void preint_alpha_table(const b200host_tf1d& tf, int res, float step, std::vector<float>& table) {
    std::vector<float> a((size_t)res, 0.f);
    const int front_end = iround(tf.threshold[0] * (float)res), back_start = iround(tf.threshold[1] * (float)res);
    for (int i = std::max(front_end, 0); i < back_start && i < res; ++i) {
        float v = (float)i / (float)(res - 1);
        if (tf.gamma != 1.f) v = std::pow(v, tf.gamma);         // applyGammaToIntensity
        uint8_t px[4];
        keys_mapping(tf.keys, tf.n_keys, v, px);
        float al = (float)px[3] / 255.f;
        if (tf.alpha_mode == 2 && al != 0.f) al = 1.f;          // TF_ONE_ALPHA
        else if (tf.alpha_mode == 0) al = 0.f;                  // TF_ZERO_ALPHA
        a[i] = al;
    }
    table.assign((size_t)res * res, 0.f);
    for (int sb = 0; sb < res; ++sb)
        for (int sf = 0; sf < res; ++sf) {
            float r;
            if (sb != sf) {
                const float scale = 1.0f / (std::fabs((float)(sb - sf)) + 1);
                const int incr = sb < sf ? -1 : 1;
                r = 0.f;
                for (int s = sf; (incr == 1 ? s <= sb : s >= sb) && ((double)r < 0.95); s += incr) {
                    float c = a[s];
                    if (c > 0.0f) {
                        c = 1.f - std::pow(1.f - c, step * 200.0f * scale);
                        r = r + (1.0f - r) * c;
                    }
                }
            } else {
                r = std::min(std::max(a[sf], 0.f), 1.f);
                r = 1.f - std::pow(1.f - r, step * 200.0f);
            }
            table[(size_t)sb * res + sf] = std::min(std::max(r, 0.f), 1.f);
        }
}
// TransFuncBase::realWorldToNormalized (transfuncbase.cpp:244-259)
float rw_to_normalized(float rw, const float dom[2]) {
    if (rw < dom[0]) return 0.f;
    if (rw > dom[1]) return 1.f;
    if (dom[0] == dom[1]) return 0.f;
    return (rw - dom[0]) / (dom[1] - dom[0]);
}

// ---- TransFunc2DPrimitives -----------------------------------------------------------------------------------------------------
struct PVert { float x, y, r, g, b, a; };
struct PCol { float r, g, b, a; };
inline PCol pcol(const uint8_t c[4]) { return {(float)c[0] / 255.f, (float)c[1] / 255.f, (float)c[2] / 255.f, (float)c[3] / 255.f}; }     // getColorNormalized
inline PVert pv(float x, float y, const PCol& c) { return {x, y, c.r, c.g, c.b, c.a}; }
inline PVert pv0(float x, float y, const PCol& c) { return {x, y, c.r, c.g, c.b, 0.f}; }                                                   // vec4(colour.xyz, 0)
// GL_TRIANGLE_STRIP / GL_TRIANGLE_FAN -> triangles
void strip(std::vector<PVert>& tris, const std::vector<PVert>& v) {
    for (size_t i = 0; i + 2 < v.size(); ++i) { tris.push_back(v[i]); tris.push_back(v[i + 1]); tris.push_back(v[i + 2]); }
}
void fan(std::vector<PVert>& tris, const std::vector<PVert>& v) {
    for (size_t i = 1; i + 1 < v.size(); ++i) { tris.push_back(v[0]); tris.push_back(v[i]); tris.push_back(v[i + 1]); }
}
// TransFuncTriangle::paint (:98-189) and TransFuncQuad::paint (:240-352): n corner points, faded border + full-colour core
void polygon_primitive(std::vector<PVert>& tris, const b200host_tf2d_primitive& p, int n) {
    float cx = 0.f, cy = 0.f;
    PCol cc = {0.f, 0.f, 0.f, 0.f}, col[4];
    for (int i = 0; i < n; ++i) {
        col[i] = pcol(p.color[i]);
        cx += p.position[i][0]; cy += p.position[i][1];
        cc.r += col[i].r; cc.g += col[i].g; cc.b += col[i].b; cc.a += col[i].a;
    }
    cx /= (float)n; cy /= (float)n; cc.r /= (float)n; cc.g /= (float)n; cc.b /= (float)n; cc.a /= (float)n;
    float mx[4], my[4], ix[4], iy[4];
    PCol mc[4];
    const float f = p.fuzziness;
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1) % n;
        mx[i] = 0.5f * p.position[i][0] + 0.5f * p.position[j][0]; my[i] = 0.5f * p.position[i][1] + 0.5f * p.position[j][1];
        mc[i] = {0.5f * col[i].r + 0.5f * col[j].r, 0.5f * col[i].g + 0.5f * col[j].g, 0.5f * col[i].b + 0.5f * col[j].b, 0.f};
        ix[i] = f * p.position[i][0] + (1.f - f) * cx; iy[i] = f * p.position[i][1] + (1.f - f) * cy;
    }
    auto outer = [&](int i) { return pv0(p.position[i][0], p.position[i][1], col[i]); };
    auto inner = [&](int i) { return pv(ix[i], iy[i], col[i]); };
    auto mid = [&](int i) { return pv(mx[i], my[i], mc[i]); };
    if (n == 3) {
        strip(tris, {outer(0), inner(0), mid(2), inner(2), outer(2), mid(1)});
        strip(tris, {inner(2), inner(1), mid(1), outer(1)});
        strip(tris, {outer(1), inner(1), mid(0), inner(0), outer(0)});
        tris.push_back(inner(0)); tris.push_back(inner(1)); tris.push_back(inner(2));
    } else {
        strip(tris, {outer(0), inner(0), mid(3), inner(3), outer(3), mid(2)});
        strip(tris, {inner(3), inner(2), mid(2), outer(2)});
        strip(tris, {outer(2), inner(2), mid(1), inner(1), outer(1), mid(0)});
        strip(tris, {inner(1), inner(0), mid(0), outer(0)});
        fan(tris, {pv(cx, cy, cc), inner(0), inner(1), inner(2), inner(3), inner(0)});
    }
}
// TransFuncBanana::paintInner (:405-490): three bands between four quadratic Bezier curves, one triangle strip
void banana_primitive(std::vector<PVert>& tris, const b200host_tf2d_primitive& p) {
    const int steps = 20;
    PCol c[4];
    for (int i = 0; i < 4; ++i) c[i] = pcol(p.color[i]);
    const float* P0 = p.position[0]; const float* P1 = p.position[1]; const float* P2 = p.position[2]; const float* P3 = p.position[3];
    auto mid2 = [](float b, float a, float d) { return (2.f * b) - (0.5f * a) - (0.5f * d); };
    const float t1[2] = {mid2(P1[0], P0[0], P3[0]), mid2(P1[1], P0[1], P3[1])}, t2[2] = {mid2(P2[0], P0[0], P3[0]), mid2(P2[1], P0[1], P3[1])};
    const PCol t1c = {mid2(c[1].r, c[0].r, c[3].r), mid2(c[1].g, c[0].g, c[3].g), mid2(c[1].b, c[0].b, c[3].b), mid2(c[1].a, c[0].a, c[3].a)};
    const PCol t2c = {mid2(c[2].r, c[0].r, c[3].r), mid2(c[2].g, c[0].g, c[3].g), mid2(c[2].b, c[0].b, c[3].b), mid2(c[2].a, c[0].a, c[3].a)};
    const float tc[2] = {(t1[0] + t2[0]) / 2.f, (t1[1] + t2[1]) / 2.f};
    const PCol tcc = {(t1c.r + t2c.r) / 2.f, (t1c.g + t2c.g) / 2.f, (t1c.b + t2c.b) / 2.f, (t1c.a + t2c.a) / 2.f};
    const float f = p.fuzziness;
    auto mix = [&](float a, float b) { return f * a + (1.f - f) * b; };
    const float t3[2] = {mix(t1[0], tc[0]), mix(t1[1], tc[1])}, t4[2] = {mix(t2[0], tc[0]), mix(t2[1], tc[1])};
    const PCol t3c = {mix(t1c.r, tcc.r), mix(t1c.g, tcc.g), mix(t1c.b, tcc.b), mix(t1c.a, tcc.a)}, t4c = {mix(t2c.r, tcc.r), mix(t2c.g, tcc.g), mix(t2c.b, tcc.b), mix(t2c.a, tcc.a)};
    auto bez = [](float t, float a, float m, float d) { return (((1 - t) * (1 - t)) * a) + ((2 * (1 - t) * t) * m) + ((t * t) * d); };
    auto bez_v = [&](float t, const float m[2], const PCol& mc, bool zero_alpha) {
        PVert v = {bez(t, P0[0], m[0], P3[0]), bez(t, P0[1], m[1], P3[1]), bez(t, c[0].r, mc.r, c[3].r), bez(t, c[0].g, mc.g, c[3].g), bez(t, c[0].b, mc.b, c[3].b),
                   bez(t, c[0].a, mc.a, c[3].a)};
        if (zero_alpha) v.a = 0.f;
        return v;
    };
    std::vector<PVert> s;
    const float* ma[3] = {t1, t3, t4}; const float* mb[3] = {t3, t4, t2};
    const PCol* ca[3] = {&t1c, &t3c, &t4c}; const PCol* cb[3] = {&t3c, &t4c, &t2c};
    for (int band = 0; band < 3; ++band) {
        s.push_back(pv(P0[0], P0[1], c[0]));
        for (int i = 0; i < steps; ++i) {
            const float t = (float)i / (float)(steps - 1);
            s.push_back(bez_v(t, ma[band], *ca[band], band == 0));      // the outermost curves carry alpha 0
            s.push_back(bez_v(t, mb[band], *cb[band], band == 2));
        }
        s.push_back(pv(P3[0], P3[1], c[3]));
    }
    strip(tris, s);
}

// "volumeoctree" mode: VolumeOctreeBackgroundThread::traverseOctreeAndCreateGeometry (optimizedproxygeometry.cpp:1657-1737) over the
// dense per-level node arrays of the library's octree (level k: (2^k)^3 nodes, x fastest; flags bit 0 exists, bit 3 has children).
// One channel (the library's octree holds one).  llf / urb = the clipped volume cube of processCube (clip bounds -+ 0.5).
int octree_proxy_geometry(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc, const b200rc_eep_params* eep, const b200host_tf1d* tfs, int n_tfs,
                          const float llf[3], const float urb[3], float* out, int max_boxes, int* n_boxes) {
    b200rc_octree_info info;
    int rc = b200rc_octree_get_info(ctx, slot, &info);
    if (rc) return rc;
    (void)n_tfs;
    const int* dim = desc->dims;
    const int L = info.num_levels;
    std::vector<std::vector<uint16_t>> mn((size_t)L), mx((size_t)L);
    std::vector<std::vector<uint8_t>> fl((size_t)L);
    for (int k = 0; k < L; ++k) {
        const size_t n = (size_t)1 << (3 * k);
        mn[k].resize(n); mx[k].resize(n); fl[k].resize(n);
        rc = b200rc_octree_download(ctx, slot, k, nullptr, mn[k].data(), mx[k].data(), fl[k].data(), nullptr);
        if (rc) return rc;
    }
    std::vector<float> pi;
    preint_alpha_table(tfs[0], 256, 1.f, pi);                                                   // getPreIntegrationTable(1.f, 256) (:1637)
    const double limit = 0.0001 * (double)eep->proxy_threshold;                                 // isRegionEmptyPi (:878-880)
    int n = 0;
    struct Item { int k, x, y, z; };
    std::vector<Item> stack;
    stack.push_back({0, 0, 0, 0});
    while (!stack.empty()) {
        const Item it = stack.back();
        stack.pop_back();
        const int G = 1 << it.k;
        const size_t idx = ((size_t)it.z * G + it.y) * G + it.x;
        if (!(fl[it.k][idx] & 1)) continue;
        const int span = info.octree_dim >> it.k;                                                // voxels per node edge at this level
        if ((fl[it.k][idx] & 8) && it.k + 1 < L) {
            // children in the reference's order (childID = x + 2 y + 4 z, :1722-1733); the stack reverses it, so push backwards
            for (int c = 7; c >= 0; --c) stack.push_back({it.k + 1, it.x * 2 + (c & 1), it.y * 2 + ((c >> 1) & 1), it.z * 2 + (c >> 2)});
            continue;
        }
        // leaf: normalised min / max -> real world -> TF domain (:1668-1676)
        const float mn_rw = ((float)mn[it.k][idx] / 65535.f) * desc->rwm_scale + desc->rwm_offset;
        const float mx_rw = ((float)mx[it.k][idx] / 65535.f) * desc->rwm_scale + desc->rwm_offset;
        const float fs = rw_to_normalized(mn_rw, tfs[0].domain), fe = rw_to_normalized(mx_rw, tfs[0].domain);
        const int i1 = std::min(std::max(iround(fs * 255.f), 0), 255), i2 = std::min(std::max(iround(fe * 255.f), 0), 255);
        if ((double)pi[(size_t)i2 * 256 + i1] <= limit) continue;
        const int pos[3] = {it.x, it.y, it.z};
        float lo[3], hi[3];
        bool outside_volume = false;
        for (int a = 0; a < 3; ++a) {
            const float node_llf = (float)(pos[a] * span), node_urb = (float)(pos[a] * span + span - 1);
            if (node_llf >= (float)dim[a]) outside_volume = true;                               // :1684-1685
            const float u = std::min(node_urb, (float)(dim[a] - 1));
            lo[a] = std::max(node_llf - 1.f, -0.5f);                                             // one voxel further for the trilinear footprint (:1693-1694)
            hi[a] = std::min(u + 1.f, (float)dim[a] - 0.5f);
        }
        if (outside_volume) continue;
        if (eep->enable_clipping) {             // containsVolume -> whole cube; intersects -> addCubeMeshClip; else nothing (:1700-1706)
            bool separated = false;
            for (int a = 0; a < 3; ++a) { lo[a] = std::max(lo[a], llf[a]); hi[a] = std::min(hi[a], urb[a]); separated = separated || lo[a] > hi[a]; }
            if (separated) continue;
        }
        if (n < max_boxes)
            for (int a = 0; a < 3; ++a) { out[6 * n + a] = (lo[a] + 0.5f) / (float)dim[a]; out[6 * n + 3 + a] = (hi[a] + 0.5f) / (float)dim[a]; }
        ++n;
    }
    *n_boxes = n;
    return B200RC_OK;
}
}  // namespace

extern "C" int b200host_tf2d_primitives_texture(const b200host_tf2d_primitive* prims, int n_prims, int width, int height,
                                                const float threshold_x[2], const float threshold_y[2], const float gamma[2], int alpha_mode,
                                                float* out_rgba32f, uint8_t* out_rgba8) {
    if (n_prims < 0 || (n_prims > 0 && !prims) || width < 2 || height < 2 || width > 4096 || height > 4096 || !threshold_x || !threshold_y || !gamma) return B200RC_ERR_INVALID;
    std::vector<PVert> tris;
    for (int i = 0; i < n_prims; ++i) {
        if (prims[i].type == 0) polygon_primitive(tris, prims[i], 3);
        else if (prims[i].type == 1) polygon_primitive(tris, prims[i], 4);
        else if (prims[i].type == 2) banana_primitive(tris, prims[i]);
        else return B200RC_ERR_INVALID;
    }
    // the help texture: updateHelpTexture + paint (:165-250,313-333)
    std::vector<float> help((size_t)width * height * 4, 0.f);
    for (size_t k = 0; k + 2 < tris.size(); k += 3) {
        const PVert &A = tris[k], &B = tris[k + 1], &Cc = tris[k + 2];
        const double area = ((double)B.x - A.x) * ((double)Cc.y - A.y) - ((double)B.y - A.y) * ((double)Cc.x - A.x);
        if (area == 0.0) continue;
        const float xmin = std::min(A.x, std::min(B.x, Cc.x)), xmax = std::max(A.x, std::max(B.x, Cc.x));
        const float ymin = std::min(A.y, std::min(B.y, Cc.y)), ymax = std::max(A.y, std::max(B.y, Cc.y));
        const int x0 = std::max(0, (int)std::floor(xmin * (float)width - 0.5f)), x1 = std::min(width - 1, (int)std::ceil(xmax * (float)width - 0.5f));
        const int y0 = std::max(0, (int)std::floor(ymin * (float)height - 0.5f)), y1 = std::min(height - 1, (int)std::ceil(ymax * (float)height - 0.5f));
        for (int y = y0; y <= y1; ++y)
            for (int x = x0; x <= x1; ++x) {
                const double px = ((double)x + 0.5) / (double)width, py = ((double)y + 0.5) / (double)height;
                const double w0 = (((double)B.x - px) * ((double)Cc.y - py) - ((double)B.y - py) * ((double)Cc.x - px)) / area;
                const double w1 = (((double)Cc.x - px) * ((double)A.y - py) - ((double)Cc.y - py) * ((double)A.x - px)) / area;
                const double w2 = (((double)A.x - px) * ((double)B.y - py) - ((double)A.y - py) * ((double)B.x - px)) / area;
                if (w0 < 0.0 || w1 < 0.0 || w2 < 0.0) continue;
                float* d = &help[((size_t)y * width + x) * 4];
                const float frag[4] = {(float)(w0 * A.r + w1 * B.r + w2 * Cc.r), (float)(w0 * A.g + w1 * B.g + w2 * Cc.g), (float)(w0 * A.b + w1 * B.b + w2 * Cc.b),
                                       (float)(w0 * A.a + w1 * B.a + w2 * Cc.a)};
                for (int ch = 0; ch < 4; ++ch) d[ch] = std::max(d[ch], frag[ch]);       // glBlendEquation(GL_MAX)
            }
    }
    // TransFunc2D::createTextureData (TF_FLOAT branch, :103-121) + getMappingForValueFloat (:150-160) + applyAlpha
    const int xf = iround(threshold_x[0] * (float)width), xb = iround(threshold_x[1] * (float)width);
    const int yf = iround(threshold_y[0] * (float)height), yb = iround(threshold_y[1] * (float)height);
    std::vector<float> tex((size_t)width * height * 4, 0.f);
    for (int y = 0; y < height; ++y) {
        if (y < yf || y > yb) continue;
        for (int x = 0; x < width; ++x) {
            if (x < xf || x > xb) continue;
            float vx = (float)x / (float)(width - 1), vy = (float)y / (float)(height - 1);
            if (gamma[0] != 1.f) vx = powf(vx, gamma[0]);
            if (gamma[1] != 1.f) vy = powf(vy, gamma[1]);
            const int tx = std::min(std::max((int)(vx * (float)width), 0), width - 1), ty = std::min(std::max((int)(vy * (float)height), 0), height - 1);
            for (int ch = 0; ch < 4; ++ch) tex[((size_t)y * width + x) * 4 + ch] = help[((size_t)ty * width + tx) * 4 + ch];
        }
    }
    if (alpha_mode != 1)
        for (size_t i = 0; i < (size_t)width * height; ++i) {
            if (alpha_mode == 2) { if (tex[i * 4 + 3] != 0.f) tex[i * 4 + 3] = 1.f; }
            else tex[i * 4 + 3] = 0.f;
        }
    if (out_rgba32f) std::memcpy(out_rgba32f, tex.data(), tex.size() * sizeof(float));
    if (out_rgba8)
        for (size_t i = 0; i < tex.size(); ++i) out_rgba8[i] = (uint8_t)std::min(std::max(iround(tex[i] * 255.f), 0), 255);
    return B200RC_OK;
}

extern "C" int b200host_optimized_proxy_geometry(b200rc_ctx* ctx, int slot, const b200rc_volume_desc* desc, const b200rc_eep_params* eep,
                                                 const b200host_tf1d* tfs, int n_tfs, float* out, int max_boxes, int* n_boxes) {
    if (!ctx || !desc || !eep || !out || !n_boxes || max_boxes < 1 || (n_tfs > 0 && !tfs)) return B200RC_ERR_INVALID;
    const int* dim = desc->dims;
    float llf[3], urb[3];
    for (int a = 0; a < 3; ++a) {           // the volume's cube incl. the half-voxel border, clipped (processCube :420-448)
        llf[a] = -0.5f; urb[a] = (float)(dim[a] - 1) + 0.5f;
        if (eep->enable_clipping) { llf[a] = std::max(llf[a], eep->clip_llf[a] - 0.5f); urb[a] = std::min(urb[a], eep->clip_urb[a] + 0.5f); }
    }
    auto emit = [&](int k, const float lo[3], const float hi[3]) {          // addCubeMesh (:480-491): tex = (coord + 0.5) / dim
        for (int a = 0; a < 3; ++a) { out[6 * k + a] = (lo[a] + 0.5f) / (float)dim[a]; out[6 * k + 3 + a] = (hi[a] + 0.5f) / (float)dim[a]; }
    };
    if (eep->proxy_mode == B200RC_PROXY_BBOX || n_tfs <= 0) { emit(0, llf, urb); *n_boxes = 1; return B200RC_OK; }

    if (eep->proxy_mode == B200RC_PROXY_VOLUME_OCTREE) return octree_proxy_geometry(ctx, slot, desc, eep, tfs, n_tfs, llf, urb, out, max_boxes, n_boxes);

    const int step = eep->proxy_cell_size < 1 ? 1 : eep->proxy_cell_size;
    int32_t cells[3] = {0, 0, 0};
    int rc = b200rc_region_minmax(ctx, slot, step, cells, nullptr, 0);
    if (rc) return rc;
    const int nc = desc->channels;
    const size_t ncell = (size_t)cells[0] * cells[1] * cells[2];
    std::vector<float> mm(ncell * nc * 2);
    rc = b200rc_region_minmax(ctx, slot, step, cells, mm.data(), mm.size());
    if (rc) return rc;
    std::vector<std::vector<float>> pi((size_t)std::min(nc, n_tfs));
    for (size_t c = 0; c < pi.size(); ++c) preint_alpha_table(tfs[c], 256, 1.f, pi[c]);         // getPreIntegrationTable(1.f, 256) (:1051)
    const double limit = 0.0001 * (double)eep->proxy_threshold;                                 // isRegionEmptyPi (:878-880)

    float bmin[3] = {3.0e38f, 3.0e38f, 3.0e38f}, bmax[3] = {-3.0e38f, -3.0e38f, -3.0e38f};
    bool defined = false;
    int n = 0;
    for (int cz = 0; cz < cells[2]; ++cz)
        for (int cy = 0; cy < cells[1]; ++cy)
            for (int cx = 0; cx < cells[0]; ++cx) {
                const size_t cell = ((size_t)cz * cells[1] + cy) * cells[0] + cx;
                bool opaque = false;
                for (size_t c = 0; c < pi.size() && !opaque; ++c) {
                    const float fs = rw_to_normalized(mm[(cell * nc + c) * 2], tfs[c].domain), fe = rw_to_normalized(mm[(cell * nc + c) * 2 + 1], tfs[c].domain);
                    const int i1 = std::min(std::max(iround(fs * 255.f), 0), 255), i2 = std::min(std::max(iround(fe * 255.f), 0), 255);      // classify (:256-280)
                    opaque = !((double)pi[c][(size_t)i2 * 256 + i1] <= limit);
                }
                if (!opaque) continue;
                const int pos[3] = {cx, cy, cz};
                float lo[3], hi[3];
                for (int a = 0; a < 3; ++a) {       // region bounds incl. the half-voxel border (:944-968)
                    const int l = pos[a] * step, u = std::min((pos[a] + 1) * step - 1, dim[a] - 1);
                    lo[a] = (float)l - 0.5f; hi[a] = (float)u + 0.5f;
                }
                if (eep->proxy_mode == B200RC_PROXY_MINBBOX) {
                    for (int a = 0; a < 3; ++a) { bmin[a] = std::min(bmin[a], lo[a]); bmax[a] = std::max(bmax[a], hi[a]); }
                    defined = true;
                } else {
                    if (eep->enable_clipping) {
                        bool outside = false;
                        for (int a = 0; a < 3; ++a) { lo[a] = std::max(lo[a], llf[a]); hi[a] = std::min(hi[a], urb[a]); outside = outside || lo[a] > hi[a]; }        // tgt::Bounds::intersects: touching still counts
                        if (outside) continue;
                    }
                    if (n < max_boxes) emit(n, lo, hi);
                    ++n;
                }
            }
    if (eep->proxy_mode == B200RC_PROXY_MINBBOX && defined) {
        bool outside = false;
        if (eep->enable_clipping)           // addCubeMeshClip against the clip bounds (:1089-1102)
            for (int a = 0; a < 3; ++a) { bmin[a] = std::max(bmin[a], llf[a]); bmax[a] = std::min(bmax[a], urb[a]); outside = outside || bmin[a] > bmax[a]; }
        if (!outside) { emit(0, bmin, bmax); n = 1; }
    }
    *n_boxes = n;
    return B200RC_OK;
}

// ---- .vvd (V/modules/core/io/vvdvolumereader.cpp:58-128, vvdformat.cpp:59-120,290-330) ---------------------------------------
// A deliberately small reader for the subset of XML the XmlSerializer writes with setUseAttributes(true): elements, attributes
// in double quotes, no CDATA, no entities in the values we read except the five predefined ones.
namespace {
struct XmlTag { std::string name; std::vector<std::pair<std::string, std::string>> attr; bool closing = false, self = false; size_t end = 0; };

std::string xml_unescape(const std::string& s) {
    std::string o; o.reserve(s.size());
    for (size_t i = 0; i < s.size(); ++i) {
        if (s[i] != '&') { o += s[i]; continue; }
        if (!s.compare(i, 5, "&amp;")) { o += '&'; i += 4; }
        else if (!s.compare(i, 4, "&lt;")) { o += '<'; i += 3; }
        else if (!s.compare(i, 4, "&gt;")) { o += '>'; i += 3; }
        else if (!s.compare(i, 6, "&quot;")) { o += '"'; i += 5; }
        else if (!s.compare(i, 6, "&apos;")) { o += '\''; i += 5; }
        else o += s[i];
    }
    return o;
}

// next tag at or after `pos` (comments, processing instructions and doctype skipped); false at the end of the text
bool xml_next(const std::string& t, size_t pos, XmlTag& tag) {
    for (;;) {
        const size_t lt = t.find('<', pos);
        if (lt == std::string::npos) return false;
        if (!t.compare(lt, 4, "<!--")) { const size_t e = t.find("-->", lt + 4); if (e == std::string::npos) return false; pos = e + 3; continue; }
        if (lt + 1 < t.size() && (t[lt + 1] == '?' || t[lt + 1] == '!')) { const size_t e = t.find('>', lt); if (e == std::string::npos) return false; pos = e + 1; continue; }
        size_t i = lt + 1;
        tag = XmlTag();
        if (i < t.size() && t[i] == '/') { tag.closing = true; ++i; }
        const size_t n0 = i;
        while (i < t.size() && !isspace((unsigned char)t[i]) && t[i] != '>' && t[i] != '/') ++i;
        tag.name = t.substr(n0, i - n0);
        for (;;) {
            while (i < t.size() && isspace((unsigned char)t[i])) ++i;
            if (i >= t.size()) return false;
            if (t[i] == '>') { tag.end = i + 1; return true; }
            if (t[i] == '/') { tag.self = true; ++i; continue; }
            const size_t k0 = i;
            while (i < t.size() && t[i] != '=' && !isspace((unsigned char)t[i]) && t[i] != '>') ++i;
            const std::string key = t.substr(k0, i - k0);
            while (i < t.size() && (isspace((unsigned char)t[i]) || t[i] == '=')) ++i;
            if (i >= t.size() || (t[i] != '"' && t[i] != '\'')) return false;
            const char q = t[i++];
            const size_t v0 = i;
            while (i < t.size() && t[i] != q) ++i;
            if (i >= t.size()) return false;
            tag.attr.emplace_back(key, xml_unescape(t.substr(v0, i - v0)));
            ++i;
        }
    }
}
const std::string* xml_attr(const XmlTag& t, const char* k) {
    for (const auto& a : t.attr) if (a.first == k) return &a.second;
    return nullptr;
}
bool xml_float(const XmlTag& t, const char* k, float& out) {
    const std::string* s = xml_attr(t, k);
    if (!s) return false;
    char* e = nullptr;
    const double v = strtod(s->c_str(), &e);
    if (e == s->c_str()) return false;
    out = (float)v;
    return true;
}
bool read_text_file(const char* path, std::string& out) {
    FILE* f = fopen(path, "rb");
    if (!f) return false;
    char buf[65536]; size_t n;
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
    fclose(f);
    return true;
}
void set_err(char* error, size_t n, const std::string& s) { if (error && n) { snprintf(error, n, "%s", s.c_str()); } }
}  // namespace

extern "C" int b200host_vvd_count(const char* vvd_path) {
    std::string t;
    if (!vvd_path || !read_text_file(vvd_path, t)) return -B200RC_ERR_INVALID;
    int n = 0; size_t pos = 0; XmlTag tag;
    while (xml_next(t, pos, tag)) { pos = tag.end; if (!tag.closing && tag.name == "Volume") ++n; }
    return n;
}

extern "C" int b200host_vvd_parse(const char* vvd_path, int index, b200rc_volume_desc* out, char* raw_path, size_t raw_path_len,
                                  char* error, size_t error_len) {
    if (!vvd_path || !out || index < 0) { set_err(error, error_len, "vvd: bad arguments"); return B200RC_ERR_INVALID; }
    std::string t;
    if (!read_text_file(vvd_path, t)) { set_err(error, error_len, std::string("Failed to open file '") + vvd_path + "' for reading."); return B200RC_ERR_INVALID; }
    // walk to <Volume> number `index`
    size_t pos = 0; XmlTag tag; int seen = -1; bool in_volume = false;
    std::string format, file, meta_name;
    int dims[3] = {0, 0, 0};
    bool have_raw = false, in_raw = false, in_paths = false, in_meta = false, other_repr = false;
    bool have_sp = false, have_off = false, have_tf = false, have_rwm = false;
    float sp[3] = {1.f, 1.f, 1.f}, off[3] = {0.f, 0.f, 0.f}, tf[16], rwm[2] = {1.f, 0.f};
    for (int i = 0; i < 16; ++i) tf[i] = (i % 5 == 0) ? 1.f : 0.f;
    while (xml_next(t, pos, tag)) {
        pos = tag.end;
        if (tag.name == "Volume") {
            if (!tag.closing) { ++seen; in_volume = (seen == index); }
            else if (in_volume) break;
            continue;
        }
        if (!in_volume) continue;
        if (tag.name == "RawData" && !tag.closing) {
            have_raw = true; in_raw = !tag.self;
            if (const std::string* s = xml_attr(tag, "format")) format = *s;
            if (const std::string* s = xml_attr(tag, "filename")) file = *s;       // old files (vvdformat.cpp:76-80)
            const char* names[3] = {"x", "y", "z"};
            for (int a = 0; a < 3; ++a) if (const std::string* s = xml_attr(tag, names[a])) dims[a] = atoi(s->c_str());
        } else if (tag.name == "RawData") in_raw = false;
        else if (tag.name == "RawDataNoExternalFile" || tag.name == "OctreeData") other_repr = true;
        else if (in_raw && tag.name == "paths") in_paths = !tag.closing && !tag.self;
        else if (in_raw && in_paths && tag.name == "item" && file.empty()) { if (const std::string* s = xml_attr(tag, "value")) file = *s; }
        else if (tag.name == "MetaItem") {
            if (tag.closing) { in_meta = false; continue; }
            in_meta = !tag.self;
            const std::string* n = xml_attr(tag, "name");
            meta_name = n ? *n : std::string();
        } else if (in_meta && tag.name == "value" && !tag.closing) {
            if (meta_name == "Spacing") have_sp = xml_float(tag, "x", sp[0]) && xml_float(tag, "y", sp[1]) && xml_float(tag, "z", sp[2]);
            else if (meta_name == "Offset") have_off = xml_float(tag, "x", off[0]) && xml_float(tag, "y", off[1]) && xml_float(tag, "z", off[2]);
            else if (meta_name == "RealWorldMapping") have_rwm = xml_float(tag, "scale", rwm[0]) && xml_float(tag, "offset", rwm[1]);
        } else if (in_meta && !tag.closing && !tag.name.compare(0, 9, "value.row") && tag.name.size() == 10 &&
                   (meta_name == "Transformation" || meta_name == "transformation")) {      // renamed from old files (vvdformat.cpp:317-318)
            const int r = tag.name[9] - '0';
            if (r >= 0 && r < 4) have_tf = (xml_float(tag, "x", tf[r * 4]) && xml_float(tag, "y", tf[r * 4 + 1]) && xml_float(tag, "z", tf[r * 4 + 2]) && xml_float(tag, "w", tf[r * 4 + 3])) || have_tf;
        }
    }
    if (seen < index) { set_err(error, error_len, "vvd: no such volume in file"); return B200RC_ERR_INVALID; }
    if (!have_raw) { set_err(error, error_len, other_repr ? "vvd: only RawData volumes are supported (RawDataNoExternalFile / OctreeData found)" : "vvd: no RawData element"); return B200RC_ERR_INVALID; }
    if (dims[0] < 1 || dims[1] < 1 || dims[2] < 1) { set_err(error, error_len, "vvd: bad dimensions"); return B200RC_ERR_INVALID; }
    // format: base type or VectorN(base)
    int channels = 1; std::string base = format;
    if (!format.compare(0, 6, "Vector") && format.size() > 9 && format[7] == '(' && format.back() == ')') {
        channels = format[6] - '0';
        base = format.substr(8, format.size() - 9);
    }
    if (channels < 1 || channels > 4 || b200host_volume_desc(dims, channels, base.c_str(), have_sp ? sp : nullptr, have_off ? off : nullptr, tf, out) != B200RC_OK) {
        set_err(error, error_len, "vvd: unsupported format '" + format + "' (uint8, uint16, float and their Vector2-4 forms are)");
        return B200RC_ERR_INVALID;
    }
    if (!have_sp) for (int a = 0; a < 3; ++a) out->spacing[a] = 1.f;           // Volume default meta data (volumebase.cpp getSpacing/getOffset)
    if (!have_off) for (int a = 0; a < 3; ++a) out->offset[a] = 0.f;
    if (have_rwm) { out->rwm_scale = rwm[0]; out->rwm_offset = rwm[1]; }       // else the denormalising mapping set by b200host_volume_desc
    (void)have_tf;
    if (raw_path && raw_path_len) {
        std::string p = file;
        if (p.empty()) { set_err(error, error_len, "vvd: RawData without a file name"); return B200RC_ERR_INVALID; }
        if (p[0] != '/') {
            const std::string v(vvd_path);
            const size_t slash = v.find_last_of('/');
            if (slash != std::string::npos) p = v.substr(0, slash + 1) + p;
        }
        if (p.size() + 1 > raw_path_len) { set_err(error, error_len, "vvd: raw path buffer too small"); return B200RC_ERR_INVALID; }
        snprintf(raw_path, raw_path_len, "%s", p.c_str());
    }
    return B200RC_OK;
}

// ---- SingleVolumeRaycaster mirror ------------------------------------------------------------------------------------------
struct b200host_svrc {
    b200rc_ctx* ctx = nullptr;
    b200rc::RaycasterProperties props;
    TransFunc tf[B200RC_MAX_CHANNELS];
    b200rc_volume_desc desc{};
    bool have_volume = false;
    bool have_camera = false;
    float cam_pos[3] = {0.f, 0.f, 3.5f}, cam_focus[3] = {0.f, 0.f, 0.f}, cam_up[3] = {0.f, 1.f, 0.f};   // CameraProperty default
    float fovy = 45.f, near_d = 0.1f, far_d = 50.f;
    // entry/exit handed over by the caller (image.entrypoints / image.exitpoints)
    int eep_w = 0, eep_h = 0;
    const float* entry = nullptr; const float* entry_depth = nullptr; const float* exitp = nullptr; const float* exit_depth = nullptr;
    // what the last beforeProcess() built (INVALID_PROGRAM analogue)
    int built_width = 0; std::string built_cls; float built_step = -1.f;
    std::string error;
    std::vector<uint8_t> tex;
};

static int fail(b200host_svrc* p, int rc, const std::string& what) {
    p->error = what;
    if (p->ctx && rc == B200RC_ERR_CUDA) p->error += std::string(": ") + b200rc_last_error(p->ctx);
    return rc;
}
static int forward(b200host_svrc* p, int rc, const char* what) {
    if (rc != B200RC_OK) p->error = std::string(what) + ": " + b200rc_last_error(p->ctx);
    return rc;
}

extern "C" int b200host_svrc_create(int device, b200host_svrc** out) {
    if (b200rc_abi_version() != B200RC_ABI_VERSION) return B200RC_ERR_INVALID;      // built against another b200rc.h: struct layouts differ
    if (!out) return B200RC_ERR_INVALID;
    *out = nullptr;
    b200host_svrc* p = new (std::nothrow) b200host_svrc();
    if (!p) return B200RC_ERR_OOM;
    const int rc = b200rc_create(device, &p->ctx);      // Processor::initialize(): no device -> no processor, no fallback
    if (rc != B200RC_OK) { delete p; return rc; }
    *out = p;
    return B200RC_OK;
}
extern "C" void b200host_svrc_destroy(b200host_svrc* p) {
    if (!p) return;
    b200rc_destroy(p->ctx);                             // Processor::deinitialize()
    delete p;
}
extern "C" const char* b200host_svrc_last_error(const b200host_svrc* p) { return p ? p->error.c_str() : "null processor"; }
extern "C" b200rc_ctx* b200host_svrc_context(b200host_svrc* p) { return p ? p->ctx : nullptr; }

extern "C" int b200host_svrc_set_property(b200host_svrc* p, const char* id, const char* value) {
    if (!p || !id || !value) return B200RC_ERR_INVALID;
    if (!p->props.set(id, value)) return fail(p, B200RC_ERR_INVALID, std::string("property not found or invalid value: ") + id + "=" + value);
    return B200RC_OK;
}
extern "C" int b200host_svrc_get_sampling_step(b200host_svrc* p, float* out) {
    if (!p || !out) return B200RC_ERR_INVALID;
    if (!p->have_volume) return fail(p, B200RC_ERR_NOT_READY, "no volume");
    *out = b200rc::sampling_step_size(p->desc.dims, p->props);
    return B200RC_OK;
}
extern "C" int b200host_svrc_set_volume(b200host_svrc* p, const void* data, const b200rc_volume_desc* desc) {
    if (!p || !data || !desc) return B200RC_ERR_INVALID;
    const int rc = forward(p, b200rc_upload_volume(p->ctx, 0, data, desc), "upload_volume");
    if (rc) return rc;
    p->desc = *desc; p->have_volume = true; p->built_width = 0;
    return B200RC_OK;
}
extern "C" int b200host_svrc_set_synthetic_volume(b200host_svrc* p, const b200rc_volume_desc* desc, uint64_t seed) {
    if (!p || !desc) return B200RC_ERR_INVALID;
    const int rc = forward(p, b200rc_synth_volume(p->ctx, 0, desc, seed), "synth_volume");
    if (rc) return rc;
    p->desc = *desc; p->have_volume = true; p->built_width = 0;
    return B200RC_OK;
}
extern "C" int b200host_svrc_set_transfer_function(b200host_svrc* p, int channel, const b200host_tf_key* keys, int n_keys,
                                                   float domain_lo, float domain_hi, const float threshold[2], float gamma) {
    if (!p || channel < 0 || channel >= B200RC_MAX_CHANNELS || (!keys && n_keys > 0) || n_keys < 0) return B200RC_ERR_INVALID;
    TransFunc& t = p->tf[channel];
    const float thr0 = threshold ? threshold[0] : 0.f, thr1 = threshold ? threshold[1] : 1.f;
    // a property set to the value it already has invalidates nothing (Property::set compares first): the per-frame hand-over of an
    // unchanged transfer function costs this comparison, not a texture / pre-integration table rebuild
    if (t.set && (int)t.keys.size() == n_keys && (n_keys == 0 || std::memcmp(t.keys.data(), keys, (size_t)n_keys * sizeof(b200host_tf_key)) == 0) &&
        t.domain[0] == domain_lo && t.domain[1] == domain_hi && t.threshold[0] == thr0 && t.threshold[1] == thr1 && t.gamma == gamma) return B200RC_OK;
    t.keys.assign(keys, keys + n_keys);
    t.domain[0] = domain_lo; t.domain[1] = domain_hi;
    t.threshold[0] = thr0; t.threshold[1] = thr1;
    t.gamma = gamma; t.set = true; t.dirty = true;
    return B200RC_OK;
}
extern "C" int b200host_svrc_set_camera(b200host_svrc* p, const float position[3], const float focus[3], const float up[3],
                                        float fovy_deg, float near_dist, float far_dist) {
    if (!p || !position || !focus || !up) return B200RC_ERR_INVALID;
    for (int i = 0; i < 3; ++i) { p->cam_pos[i] = position[i]; p->cam_focus[i] = focus[i]; p->cam_up[i] = up[i]; }
    p->fovy = fovy_deg; p->near_d = near_dist; p->far_d = far_dist; p->have_camera = true;
    return B200RC_OK;
}
extern "C" int b200host_svrc_connect_output(b200host_svrc* p, int port, int connected) {
    if (!p || port < 0 || port >= B200RC_MAX_OUTPUTS) return B200RC_ERR_INVALID;
    p->props.outputConnected[port] = connected != 0;
    return B200RC_OK;
}
extern "C" int b200host_svrc_set_entry_exit(b200host_svrc* p, int width, int height, const float* entry_rgba, const float* entry_depth,
                                            const float* exit_rgba, const float* exit_depth) {
    if (!p) return B200RC_ERR_INVALID;
    p->eep_w = width; p->eep_h = height;
    p->entry = entry_rgba; p->entry_depth = entry_depth; p->exitp = exit_rgba; p->exit_depth = exit_depth;
    return B200RC_OK;
}
extern "C" int b200host_svrc_is_ready(b200host_svrc* p) {
    if (!p || !p->have_volume) return 0;
    if (!p->props.outputConnected[0] && !p->props.outputConnected[1] && !p->props.outputConnected[2]) return 0;
    for (int c = 0; c < p->desc.channels; ++c)
        if (!p->tf[c].set) return 0;
    return 1;
}

// beforeProcess(): (re)build TF textures / pre-integration tables the current properties need
static int before_process(b200host_svrc* p, const b200rc_params& P, int builder) {
    int width = 0;
    b200host_tf1d_width_for(p->desc.dtype, &width);
    const std::string cls = b200rc::effective_classification(p->props, false);
    const bool rebuild_all = width != p->built_width || cls != p->built_cls || (builder >= 0 && P.sampling_step != p->built_step);
    for (int c = 0; c < p->desc.channels; ++c) {
        TransFunc& t = p->tf[c];
        if (!t.dirty && !rebuild_all) continue;
        if (builder >= 0) {
            // PreIntegrationTableMap: resolution min(tfWidth, 1024) (preintegrationtablemap.cpp:66-67).  The CPU builders
            // index the TF at table resolution, the GPU shader samples the full-width texture (preintegration.frag:64).
            const int res = width < 1024 ? width : 1024;
            const int w_build = builder == B200RC_PI_GPU ? width : res;
            p->tex.resize((size_t)w_build * 4);
            b200host_tf1d_keys_texture(t.keys.data(), (int)t.keys.size(), t.threshold, t.gamma, 1, w_build, p->tex.data());
            int rc = forward(p, b200rc_set_tf1d(p->ctx, c, p->tex.data(), w_build, t.domain[0], t.domain[1]), "set_tf1d");
            if (rc) return rc;
            rc = forward(p, b200rc_build_preint(p->ctx, c, res, P.sampling_step, builder), "build_preint");
            if (rc) return rc;
        }
        p->tex.resize((size_t)width * 4);
        b200host_tf1d_keys_texture(t.keys.data(), (int)t.keys.size(), t.threshold, t.gamma, 1, width, p->tex.data());
        const int rc = forward(p, b200rc_set_tf1d(p->ctx, c, p->tex.data(), width, t.domain[0], t.domain[1]), "set_tf1d");
        if (rc) return rc;
        t.dirty = false;
    }
    p->built_width = width; p->built_cls = cls; p->built_step = P.sampling_step;
    return B200RC_OK;
}

extern "C" int b200host_svrc_process(b200host_svrc* p, int width, int height, void* out_color[B200RC_MAX_OUTPUTS], float* out_depth) {
    if (!p || width < 1 || height < 1) return B200RC_ERR_INVALID;
    if (!b200host_svrc_is_ready(p)) return fail(p, B200RC_ERR_NOT_READY, "processor not ready (volume, transfer function or outport missing)");
    b200rc_params P; int builder = -1; std::string err;
    if (!b200rc::to_params(p->props, p->desc.dims, false, &P, &builder, &err)) return fail(p, B200RC_ERR_INVALID, err);
    int rc = before_process(p, P, builder);
    if (rc) return rc;
    b200rc_camera cam;
    b200host_camera(p->cam_pos, p->cam_focus, p->cam_up, p->fovy, 1.f, p->near_d, p->far_d, width, height, &cam);
    if (p->entry && p->exitp && p->entry_depth && p->exit_depth) {
        if (p->eep_w != width || p->eep_h != height) return fail(p, B200RC_ERR_INVALID, "entry/exit size differs from the output size");
        return forward(p, b200rc_render(p->ctx, &P, &cam, width, height, p->entry, p->entry_depth, p->exitp, p->exit_depth, out_color, out_depth), "render");
    }
    b200rc_eep_params E;
    b200rc::to_eep_params(p->props, &E);
    return forward(p, b200rc_render_camera(p->ctx, &P, &E, &cam, width, height, out_color, out_depth), "render_camera");
}
