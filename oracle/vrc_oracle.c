/*
 * vrc_oracle.c -- CPU restatement (C99 + OpenMP) of Voreen's direct-volume-ray-casting path.
 * TEST INFRASTRUCTURE ONLY -- see vrc_oracle.h for the rules and the parity status ("parity unpinned" at image
 * level; camera math pinned against the reference's own tgt headers, everything else pinned by cited constants).
 *
 * Build: see oracle/Makefile (gcc -O2 -fopenmp -mfma -ffp-contract=off: fmaf() is used only where written).
 * Every function cites the reference file:line it follows.  V/ = /root/reference/voreen/
 */
#include "vrc_oracle.h"

#include <math.h>
#include <stdlib.h>
#include <string.h>
#include <float.h>
#ifdef _OPENMP
#include <omp.h>
#endif

/* ------------------------------------------------------------------------------------------------------------ */
/* small vector helpers                                                                                          */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct { float x, y, z; } v3;
typedef struct { float x, y, z, w; } v4;

static inline v3 V3(float x, float y, float z) { v3 r = {x, y, z}; return r; }
static inline v4 V4(float x, float y, float z, float w) { v4 r = {x, y, z, w}; return r; }
static inline v3 add3(v3 a, v3 b) { return V3(a.x + b.x, a.y + b.y, a.z + b.z); }
static inline v3 sub3(v3 a, v3 b) { return V3(a.x - b.x, a.y - b.y, a.z - b.z); }
static inline v3 mul3(v3 a, v3 b) { return V3(a.x * b.x, a.y * b.y, a.z * b.z); }
static inline v3 scale3(v3 a, float s) { return V3(a.x * s, a.y * s, a.z * s); }
static inline float dot3(v3 a, v3 b) { return (a.x * b.x + a.y * b.y) + a.z * b.z; }
static inline v3 cross3(v3 a, v3 b) { return V3(a.y * b.z - a.z * b.y, a.z * b.x - a.x * b.z, a.x * b.y - a.y * b.x); }
static inline float length3(v3 a) { return sqrtf(dot3(a, a)); }
/* GLSL normalize(): x / length(x); a zero vector yields NaN exactly like 0 * rsq(0) does on the GL drivers. */
static inline v3 normalize3(v3 a) { float l = length3(a); return V3(a.x / l, a.y / l, a.z / l); }
/* GLSL max()/min() as the drivers the reference runs on implement them (maxps / FMNMX): a NaN operand is dropped. */
static inline float gl_max(float a, float b) { return fmaxf(a, b); }
static inline float gl_min(float a, float b) { return fminf(a, b); }
static inline float clampf(float v, float lo, float hi) { return v < lo ? lo : (v > hi ? hi : v); }
static inline v3 mix3(v3 a, v3 b, float t) { return V3(a.x * (1.f - t) + b.x * t, a.y * (1.f - t) + b.y * t, a.z * (1.f - t) + b.z * t); }

static int g_threads = 0;
int orc_abi_version(void) { return B200RC_ABI_VERSION; }
void orc_set_threads(int n) {
    g_threads = n;
#ifdef _OPENMP
    if (n > 0) omp_set_num_threads(n);
#endif
}

/* ------------------------------------------------------------------------------------------------------------ */
/* tgt matrices (row-major, column vectors)                                                                      */
/* ------------------------------------------------------------------------------------------------------------ */

/* V/ext/tgt/matrix.h operator*(Matrix4, Matrix4): result(r,c) = sum_k a(r,k) b(k,c), accumulated left to right */
void orc_mat4_mul(const float a[16], const float b[16], float out[16]) {
    float r[16];
    for (int i = 0; i < 4; ++i)
        for (int j = 0; j < 4; ++j) {
            float s = 0.f;
            for (int k = 0; k < 4; ++k) s += a[i * 4 + k] * b[k * 4 + j];
            r[i * 4 + j] = s;
        }
    memcpy(out, r, sizeof r);
}

/* V/ext/tgt/matrix.h:1327-1376 Matrix4<T>::invert: Gauss-Jordan; note the pivot search compares against the
 * diagonal element, not against the running maximum (kept as is). */
int orc_mat4_invert(const float m[16], float out[16]) {
    float tmp[16], res[16];
    memcpy(tmp, m, sizeof tmp);
    for (int i = 0; i < 16; ++i) res[i] = (i % 5 == 0) ? 1.f : 0.f;
    for (int i = 0; i < 4; ++i) {
        int swap = i;
        for (int j = i + 1; j < 4; ++j)
            if (fabsf(tmp[j * 4 + i]) > fabsf(tmp[i * 4 + i])) swap = j;
        if (swap != i)
            for (int k = 0; k < 4; ++k) {
                float t = tmp[i * 4 + k]; tmp[i * 4 + k] = tmp[swap * 4 + k]; tmp[swap * 4 + k] = t;
                t = res[i * 4 + k]; res[i * 4 + k] = res[swap * 4 + k]; res[swap * 4 + k] = t;
            }
        if (tmp[i * 4 + i] == 0.f) return 0;
        float t = tmp[i * 4 + i];
        for (int k = 0; k < 4; ++k) { tmp[i * 4 + k] /= t; res[i * 4 + k] /= t; }
        for (int j = 0; j < 4; ++j)
            if (j != i) {
                t = tmp[j * 4 + i];
                for (int k = 0; k < 4; ++k) {
                    tmp[j * 4 + k] -= tmp[i * 4 + k] * t;
                    res[j * 4 + k] -= res[i * 4 + k] * t;
                }
            }
    }
    memcpy(out, res, sizeof res);
    return 1;
}

static inline v4 mat4_mul_v4(const float m[16], v4 v) {
    v4 r;
    r.x = ((m[0] * v.x + m[1] * v.y) + m[2] * v.z) + m[3] * v.w;
    r.y = ((m[4] * v.x + m[5] * v.y) + m[6] * v.z) + m[7] * v.w;
    r.z = ((m[8] * v.x + m[9] * v.y) + m[10] * v.z) + m[11] * v.w;
    r.w = ((m[12] * v.x + m[13] * v.y) + m[14] * v.z) + m[15] * v.w;
    return r;
}
static inline v3 mat4_mul_point(const float m[16], v3 p) { v4 r = mat4_mul_v4(m, V4(p.x, p.y, p.z, 1.f)); return V3(r.x, r.y, r.z); }

/* tgt normalize: v * (1/length(v))  (V/ext/tgt/vector.h:815-820); lengthSq accumulates 0+x*x+y*y+z*z (:802-808) */
static inline v3 tgt_normalize(v3 v) {
    float l2 = 0.f; l2 += v.x * v.x; l2 += v.y * v.y; l2 += v.z * v.z;
    float r = 1.f / sqrtf(l2);
    return V3(v.x * r, v.y * r, v.z * r);
}

/* V/ext/tgt/matrix.h:1196-1210 createLookAt */
void orc_look_at(const float eye[3], const float focus[3], const float up[3], float out[16]) {
    v3 e = V3(eye[0], eye[1], eye[2]);
    v3 look = tgt_normalize(sub3(V3(focus[0], focus[1], focus[2]), e));
    v3 strafe = tgt_normalize(cross3(look, tgt_normalize(V3(up[0], up[1], up[2]))));
    v3 up2 = cross3(strafe, look);
    float m[16] = { strafe.x, strafe.y, strafe.z, 0.f,
                    up2.x, up2.y, up2.z, 0.f,
                    -look.x, -look.y, -look.z, 0.f,
                    0.f, 0.f, 0.f, 1.f };
    float t[16] = { 1.f, 0.f, 0.f, -e.x,  0.f, 1.f, 0.f, -e.y,  0.f, 0.f, 1.f, -e.z,  0.f, 0.f, 0.f, 1.f };
    orc_mat4_mul(m, t, out);
}

/* V/ext/tgt/matrix.h:1230-1248 createPerspective */
void orc_perspective(float fov, float aspect, float pnear, float pfar, float out[16]) {
    /* `tan` inside tgt resolves to the C double overload (verified against the reference's own header by
     * oracle/_ref/tgt_probe): T(1) / tan(fov / T(2)) is evaluated in double and rounded once */
    float f = (float)(1.0 / tan((double)(fov / 2.f)));
    float m[16] = { f / aspect, 0.f, 0.f, 0.f,
                    0.f, f, 0.f, 0.f,
                    0.f, 0.f, (pnear + pfar) / (pnear - pfar), (2.f * pfar * pnear) / (pnear - pfar),
                    0.f, 0.f, -1.f, 0.f };
    memcpy(out, m, sizeof m);
}

/* V/ext/tgt/camera.cpp:188-217 Camera::getProjectionMatrix, PERSPECTIVE branch (fovy clamp 6..175 degrees);
 * deg2rad = (PI / 180) * f with PIf (V/ext/tgt/tgt_math.h) */
void orc_camera_projection(float fovy_deg, float frustum_ratio, float pnear, float pfar, float window_ratio, float out[16]) {
    const float PIf = 3.14159265358979323846f;
    float fovy = fovy_deg;
    if (fovy < 6.f) fovy = 6.f;
    if (fovy > 175.f) fovy = 175.f;
    if (window_ratio >= 1.0f)
        orc_perspective((PIf / 180.f) * fovy, frustum_ratio * window_ratio, pnear, pfar, out);
    else
        orc_perspective((float)(atan(tan((double)((PIf / 180.f) * (fovy / 2.f))) / (double)(window_ratio * frustum_ratio)) * 2),
                        frustum_ratio * window_ratio, pnear, pfar, out);
}

/* volumebase.cpp:190-231: textureToWorld = physicalToWorld * (translate(offset) * scale(spacing)) *
 *                                          (translate(-0.5) * scale(dims)) */
void orc_texture_to_world(const b200rc_volume_desc* d, float out[16]) {
    float tex2vox[16] = { (float)d->dims[0], 0, 0, -0.5f,  0, (float)d->dims[1], 0, -0.5f,  0, 0, (float)d->dims[2], -0.5f,  0, 0, 0, 1 };
    float vox2phys[16] = { d->spacing[0], 0, 0, d->offset[0],  0, d->spacing[1], 0, d->offset[1],  0, 0, d->spacing[2], d->offset[2],  0, 0, 0, 1 };
    /* createTranslation(t) * createScale(s) evaluated by the generic 4x4 product, as tgt does */
    float T1[16] = { 1, 0, 0, -0.5f, 0, 1, 0, -0.5f, 0, 0, 1, -0.5f, 0, 0, 0, 1 };
    float S1[16] = { (float)d->dims[0], 0, 0, 0, 0, (float)d->dims[1], 0, 0, 0, 0, (float)d->dims[2], 0, 0, 0, 0, 1 };
    float T2[16] = { 1, 0, 0, d->offset[0], 0, 1, 0, d->offset[1], 0, 0, 1, d->offset[2], 0, 0, 0, 1 };
    float S2[16] = { d->spacing[0], 0, 0, 0, 0, d->spacing[1], 0, 0, 0, 0, d->spacing[2], 0, 0, 0, 0, 1 };
    (void)tex2vox; (void)vox2phys;
    float a[16], b[16], vox2world[16];
    orc_mat4_mul(T1, S1, a);            /* getTextureToVoxelMatrix */
    orc_mat4_mul(T2, S2, b);            /* getVoxelToPhysicalMatrix */
    orc_mat4_mul(d->phys_to_world, b, vox2world);   /* getVoxelToWorldMatrix */
    orc_mat4_mul(vox2world, a, out);    /* getTextureToWorldMatrix */
}
void orc_world_to_physical(const b200rc_volume_desc* d, float out[16]) {
    if (!orc_mat4_invert(d->phys_to_world, out)) {   /* volumebase.cpp:206-210: identity when singular */
        for (int i = 0; i < 16; ++i) out[i] = (i % 5 == 0) ? 1.f : 0.f;
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* framebuffer formats                                                                                           */
/* ------------------------------------------------------------------------------------------------------------ */
/* GL_RGBA16 store (clamp, round to nearest) followed by a fetch (c / 65535) */
float orc_unorm16(float v) {
    if (!(v > 0.f)) return 0.f;             /* NaN and negatives -> 0 */
    if (v > 1.f) v = 1.f;
    float q = floorf(v * 65535.f + 0.5f);
    return q / 65535.f;
}
/* GL_DEPTH_COMPONENT24 */
float orc_depth24(float v) {
    if (!(v > 0.f)) return 0.f;
    if (v > 1.f) v = 1.f;
    double q = floor((double)v * 16777215.0 + 0.5);
    return (float)(q / 16777215.0);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* transfer functions                                                                                            */
/* ------------------------------------------------------------------------------------------------------------ */
static inline int iround_tgt(float f) { return (int)floorf(f + 0.5f); }   /* V/ext/tgt/tgt_math.h:213-225 */

/* TransFunc1DKeys::getMappingForValueUByte (transfunc1dkeys.cpp:119-154): piecewise linear in uint8 WITH TRUNCATION */
static void tf_keys_mapping_ubyte(const orc_tf1d* tf, float value, uint8_t out[4]) {
    if (tf->n_keys == 0) { out[0] = out[1] = out[2] = out[3] = 0; return; }
    value = (value < 0.f) ? 0.f : value;
    value = (value > 1.f) ? 1.f : value;
    int it = 0;
    while (it != tf->n_keys && value > tf->keys[it].intensity) it++;
    if (it == 0) { memcpy(out, tf->keys[0].color_l, 4); return; }
    if (it == tf->n_keys) { memcpy(out, tf->keys[it - 1].color_r, 4); return; }
    const orc_tf_key* lk = &tf->keys[it - 1];
    const orc_tf_key* rk = &tf->keys[it];
    float fraction = (value - lk->intensity) / (rk->intensity - lk->intensity);
    for (int c = 0; c < 4; ++c) {
        /* color.r += (rightDest.r - leftDest.r) * fraction;  -- GLubyte += float: float sum, truncating conversion */
        float sum = (float)lk->color_r[c] + (float)((int)rk->color_l[c] - (int)lk->color_r[c]) * fraction;
        out[c] = (uint8_t)(int)sum;
    }
}

/* TransFunc1D::applyGammaToIntensity (transfunc1d.cpp:144-149) */
static inline float tf_apply_gamma(const orc_tf1d* tf, float v) { return tf->gamma != 1.f ? powf(v, tf->gamma) : v; }

/* TransFunc1D::createTextureData (transfunc1d.cpp:85-116) + TransFuncBase::applyAlpha (transfuncbase.cpp:177-210).
 * Quirk kept: the threshold is compared against round(thr*width) with an inclusive upper end (x > back_start). */
void orc_tf1d_texture(const orc_tf1d* tf, int width, uint8_t* out) {
    int front_end = iround_tgt(tf->threshold[0] * (float)width);
    int back_start = iround_tgt(tf->threshold[1] * (float)width);
    for (int x = 0; x < width; ++x) {
        uint8_t* px = out + 4 * x;
        if (x < front_end || x > back_start) { px[0] = px[1] = px[2] = px[3] = 0; }
        else tf_keys_mapping_ubyte(tf, tf_apply_gamma(tf, (float)x / (float)(width - 1)), px);
        if (tf->alpha_mode == ORC_TF_ONE_ALPHA) { if (px[3]) px[3] = 255; }
        else if (tf->alpha_mode == ORC_TF_ZERO_ALPHA) px[3] = 0;
    }
}

/* PreIntegrationTable::computeTable (preintegrationtable.cpp:67-204). table[sb*res + sf]. */
void orc_preint_cpu(const orc_tf1d* tf, int res, float step, int use_integral, float* out) {
    if (res <= 1) res = 256;        /* :42-45 */
    if (step <= 0.f) step = 1.f;
    v4* tfb = (v4*)malloc(sizeof(v4) * (size_t)res);
    int front_end = iround_tgt(tf->threshold[0] * (float)res);
    int back_start = iround_tgt(tf->threshold[1] * (float)res);
    for (int i = 0; i < res; ++i) tfb[i] = V4(0, 0, 0, 0);
    for (int i = front_end; i < back_start && i < res; ++i) {
        if (i < 0) continue;
        float intensity = (float)i / (float)(res - 1);
        uint8_t ub[4];
        tf_keys_mapping_ubyte(tf, tf_apply_gamma(tf, intensity), ub);
        v4 value = V4(ub[0] / 255.f, ub[1] / 255.f, ub[2] / 255.f, ub[3] / 255.f);
        if (tf->alpha_mode == ORC_TF_ONE_ALPHA && value.w != 0.f) value.w = 1.f;
        else if (tf->alpha_mode == ORC_TF_ZERO_ALPHA) value.w = 0.f;
        tfb[i] = value;
    }
    v4* table = (v4*)out;
    if (!use_integral) {
#pragma omp parallel for schedule(dynamic, 4)
        for (int sb = 0; sb < res; ++sb) {
            for (int sf = 0; sf < res; ++sf) {
                v4 result;
                if (sb != sf) {
                    float scale = 1.0f / (fabsf((float)(sb - sf)) + 1.f);
                    int incr = (sb < sf) ? -1 : 1;
                    result = V4(0, 0, 0, 0);
                    for (int s = sf; (incr == 1 ? s <= sb : s >= sb) && (result.w < 0.95f); s += incr) {
                        v4 cur = tfb[s];
                        if (cur.w > 0.0f) {
                            cur.w = 1.f - powf(1.f - cur.w, step * 200.0f * scale);
                            float k = (1.0f - result.w) * cur.w;
                            result.x = result.x + k * cur.x;
                            result.y = result.y + k * cur.y;
                            result.z = result.z + k * cur.z;
                            result.w = result.w + (1.0f - result.w) * cur.w;
                        }
                    }
                    float d = fmaxf(result.w, 0.001f);
                    result.x /= d; result.y /= d; result.z /= d;
                    result = V4(clampf(result.x, 0, 1), clampf(result.y, 0, 1), clampf(result.z, 0, 1), clampf(result.w, 0, 1));
                } else {
                    result = V4(clampf(tfb[sf].x, 0, 1), clampf(tfb[sf].y, 0, 1), clampf(tfb[sf].z, 0, 1), clampf(tfb[sf].w, 0, 1));
                    result.w = 1.f - powf(1.f - result.w, step * 200.0f);
                }
                table[(size_t)sb * res + sf] = result;
            }
        }
    } else {
        /* integral functions, Real-Time Volume Graphics p.96 (:146-201) */
        v4* integ = (v4*)malloc(sizeof(v4) * (size_t)res);
        v4 acc = V4(0, 0, 0, 0);
        for (int i = 0; i < res; ++i) {
            v4 c = tfb[i];
            if (c.w > 0.0f) { acc.x += c.x * c.w; acc.y += c.y * c.w; acc.z += c.z * c.w; acc.w += c.w; }
            integ[i] = acc;
        }
        for (int sb = 0; sb < res; ++sb)
            for (int sf = 0; sf < res; ++sf) {
                int smin = sb < sf ? sb : sf, smax = sb < sf ? sf : sb;
                v4 col;
                if (smax != smin) {
                    float factor = step * 200.f / (float)(smax - smin);
                    col.x = (integ[smax].x - integ[smin].x) * factor;
                    col.y = (integ[smax].y - integ[smin].y) * factor;
                    col.z = (integ[smax].z - integ[smin].z) * factor;
                    col.w = 1.f - expf(-(integ[smax].w - integ[smin].w) * factor);
                    float d = fmaxf(col.w, 0.001f);
                    col.x /= d; col.y /= d; col.z /= d;
                } else {
                    col = tfb[smin];
                    col.w = 1.f - powf(1.f - col.w, step * 200.0f);
                }
                table[(size_t)sb * res + sf] = V4(clampf(col.x, 0, 1), clampf(col.y, 0, 1), clampf(col.z, 0, 1), clampf(col.w, 0, 1));
            }
        free(integ);
    }
    free(tfb);
}

/* sampler1D NEAREST, CLAMP_TO_EDGE: texel floor(s*w) clamped (GL 4.x 8.14.2) */
static inline v4 tf1d_fetch(const uint8_t* tex, int width, float s) {
    int i = (int)floorf(s * (float)width);
    if (i < 0) i = 0;
    if (i > width - 1) i = width - 1;
    const uint8_t* p = tex + 4 * i;
    const float k = 1.0f / 255.0f;      /* unorm8 -> float */
    return V4(p[0] * k, p[1] * k, p[2] * k, p[3] * k);
}

/* preintegration.frag:35-87, fragment (x = sf, y = sb), target GL_RGBA16 */
void orc_preint_gpu(const uint8_t* tf, int tfw, int res, float step, float* out) {
    v4* table = (v4*)out;
    float rcp = 1.0f / (float)res;      /* texParams_.dimensionsRCP_ */
#pragma omp parallel for schedule(dynamic, 4)
    for (int y = 0; y < res; ++y)
        for (int x = 0; x < res; ++x) {
            float fx = (float)x + 0.5f, fy = (float)y + 0.5f;   /* gl_FragCoord */
            v4 c = V4(0, 0, 0, 0);
            if (fx == fy) {
                c = tf1d_fetch(tf, tfw, fx * rcp);
                c.w = 1.0f - powf(1.0f - c.w, step * 200.0f);
            } else {
                int sf = (int)fx, sb = (int)fy;
                float scale = 1.f / (fabsf((float)(sb - sf)) + 1.0f);
                int incr = (sb - sf) > 0 ? 1 : ((sb - sf) < 0 ? -1 : 0);
                for (int s = sf; ((incr * s) <= (incr * sb)) && (c.w < 0.95f); s += incr) {
                    v4 cur = tf1d_fetch(tf, tfw, (float)s * rcp);   /* no half-texel offset (quirk) */
                    if (cur.w > 0.0f) {
                        cur.w = 1.0f - powf(1.0f - cur.w, step * 200.0f * scale);
                        float k = (1.0f - c.w) * cur.w;
                        c.x = c.x + k * cur.x; c.y = c.y + k * cur.y; c.z = c.z + k * cur.z;
                        c.w = c.w + (1.0f - c.w) * cur.w;
                    }
                }
                float d = fmaxf(c.w, 0.001f);
                c.x /= d; c.y /= d; c.z /= d;
                c = V4(clampf(c.x, 0, 1), clampf(c.y, 0, 1), clampf(c.z, 0, 1), clampf(c.w, 0, 1));
            }
            table[(size_t)y * res + x] = V4(orc_unorm16(c.x), orc_unorm16(c.y), orc_unorm16(c.z), orc_unorm16(c.w));
        }
}

/* PreIntegrationTable::classify (preintegrationtable.cpp:256-279) */
void orc_preint_classify(const float* table, int res, float fs, float fe, float out[4]) {
    int i1 = iround_tgt(fs * (float)(res - 1));
    int i2 = iround_tgt(fe * (float)(res - 1));
    if (i1 < 0) i1 = 0; if (i1 > res - 1) i1 = res - 1;
    if (i2 < 0) i2 = 0; if (i2 > res - 1) i2 = res - 1;
    memcpy(out, table + 4 * ((size_t)i2 * res + i1), 4 * sizeof(float));
}

/* ------------------------------------------------------------------------------------------------------------ */
/* volume sampling (GL 4.x spec 8.14 with float weights; mod_sampler3d.frag:90-101,151-153)                        */
/* ------------------------------------------------------------------------------------------------------------ */
static inline float raw_texel(const orc_volume* v, int64_t idx) {
    switch (v->desc.dtype) {
        case B200RC_U8:  return (float)((const uint8_t*)v->data)[idx] * (1.0f / 255.0f);
        case B200RC_U16: return (float)((const uint16_t*)v->data)[idx] * (1.0f / 65535.0f);
        default:         return ((const float*)v->data)[idx];
    }
}

/* texel after the VolumeGL swizzle (volumegl.cpp:366-373): 1ch (R,0,0,R); 2ch (R,G,0,G); 3ch (R,G,B,1); 4ch RGBA.
 * Out-of-range indices: clamp-to-edge or the border colour vec4(borderIntensity) (singlevolumeraycaster.cpp:299-303). */
static inline v4 texel_fetch(const orc_volume* v, int x, int y, int z, int clamp_mode, float border) {
    const int* d = v->desc.dims;
    int nc = v->desc.channels;
    float c[4] = {0, 0, 0, 0};
    int oob = (x < 0 || y < 0 || z < 0 || x >= d[0] || y >= d[1] || z >= d[2]);
    if (oob && clamp_mode == B200RC_CLAMP_TO_BORDER) {
        for (int i = 0; i < nc; ++i) c[i] = border;
    } else {
        if (x < 0) x = 0; if (x > d[0] - 1) x = d[0] - 1;
        if (y < 0) y = 0; if (y > d[1] - 1) y = d[1] - 1;
        if (z < 0) z = 0; if (z > d[2] - 1) z = d[2] - 1;
        /* windowed storage (bench.py's bounded CPU sample): `data` holds only voxels [shard_lo - apron, shard_hi + apron)
         * of the global volume; indices outside the window are clamped into it (memory safety only -- the caller sizes
         * the window so that no sample of the rays it casts reaches outside) */
        int rl[3] = {0, 0, 0}, rd[3] = {d[0], d[1], d[2]};
        if (v->desc.shard_hi[0] > v->desc.shard_lo[0])
            for (int a = 0; a < 3; ++a) {
                int lo = v->desc.shard_lo[a] - v->desc.shard_apron, hi = v->desc.shard_hi[a] + v->desc.shard_apron;
                if (lo < 0) lo = 0; if (hi > d[a]) hi = d[a];
                rl[a] = lo; rd[a] = hi - lo;
            }
        x -= rl[0]; y -= rl[1]; z -= rl[2];
        if (x < 0) x = 0; if (x > rd[0] - 1) x = rd[0] - 1;
        if (y < 0) y = 0; if (y > rd[1] - 1) y = rd[1] - 1;
        if (z < 0) z = 0; if (z > rd[2] - 1) z = rd[2] - 1;
        int64_t base = (((int64_t)z * rd[1] + y) * rd[0] + x) * nc;
        for (int i = 0; i < nc; ++i) c[i] = raw_texel(v, base + i);
    }
    switch (nc) {
        case 1: return V4(c[0], 0.f, 0.f, c[0]);
        case 2: return V4(c[0], c[1], 0.f, c[1]);
        case 3: return V4(c[0], c[1], c[2], 1.f);
        default: return V4(c[0], c[1], c[2], c[3]);
    }
}

static inline float lerp1(float a, float b, float w) { return fmaf(w, b - a, a); }
static inline v4 lerp4(v4 a, v4 b, float w) { return V4(lerp1(a.x, b.x, w), lerp1(a.y, b.y, w), lerp1(a.z, b.z, w), lerp1(a.w, b.w, w)); }

/* texture(sampler3D, p): un-mapped normalised value (textureLookup3DUnnormalized) */
static v4 tex3d(const orc_volume* v, int filter, int clamp_mode, float border, v3 p) {
    const int* d = v->desc.dims;
    if (filter == B200RC_FILTER_NEAREST) {
        int x = (int)floorf(p.x * (float)d[0]), y = (int)floorf(p.y * (float)d[1]), z = (int)floorf(p.z * (float)d[2]);
        return texel_fetch(v, x, y, z, clamp_mode, border);
    }
    float ux = p.x * (float)d[0] - 0.5f, uy = p.y * (float)d[1] - 0.5f, uz = p.z * (float)d[2] - 0.5f;
    float fx = floorf(ux), fy = floorf(uy), fz = floorf(uz);
    float ax = ux - fx, ay = uy - fy, az = uz - fz;
    int x0 = (int)fx, y0 = (int)fy, z0 = (int)fz;
    v4 c000 = texel_fetch(v, x0, y0, z0, clamp_mode, border), c100 = texel_fetch(v, x0 + 1, y0, z0, clamp_mode, border);
    v4 c010 = texel_fetch(v, x0, y0 + 1, z0, clamp_mode, border), c110 = texel_fetch(v, x0 + 1, y0 + 1, z0, clamp_mode, border);
    v4 c001 = texel_fetch(v, x0, y0, z0 + 1, clamp_mode, border), c101 = texel_fetch(v, x0 + 1, y0, z0 + 1, clamp_mode, border);
    v4 c011 = texel_fetch(v, x0, y0 + 1, z0 + 1, clamp_mode, border), c111 = texel_fetch(v, x0 + 1, y0 + 1, z0 + 1, clamp_mode, border);
    v4 c00 = lerp4(c000, c100, ax), c10 = lerp4(c010, c110, ax), c01 = lerp4(c001, c101, ax), c11 = lerp4(c011, c111, ax);
    v4 c0 = lerp4(c00, c10, ay), c1 = lerp4(c01, c11, ay);
    return lerp4(c0, c1, az);
}

/* getVoxel = textureLookup3D: result *= rwmScale; result += rwmOffset (mod_sampler3d.frag:90-101) */
static inline v4 get_voxel(const orc_volume* v, const b200rc_volume_params* vp, v3 p) {
    v4 r = tex3d(v, vp->filter, vp->clamp, vp->border_intensity, p);
    float s = v->desc.rwm_scale, o = v->desc.rwm_offset;
    r.x = r.x * s; r.y = r.y * s; r.z = r.z * s; r.w = r.w * s;
    r.x = r.x + o; r.y = r.y + o; r.z = r.z + o; r.w = r.w + o;
    return r;
}

void orc_sample3d(const orc_volume* vol, int filter, int clamp_mode, float border, const float pos[3], float out[4]) {
    v4 r = tex3d(vol, filter, clamp_mode, border, V3(pos[0], pos[1], pos[2]));
    out[0] = r.x; out[1] = r.y; out[2] = r.z; out[3] = r.w;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* gradients (mod_gradients.frag)                                                                                */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct {
    const orc_volume* vol;
    const b200rc_volume_params* vp;
    v3 dims_rcp, spacing_rcp, dims, spacing, offset;
} vol_ctx;

static inline float tex_a(const vol_ctx* c, v3 p) { return tex3d(c->vol, c->vp->filter, c->vp->clamp, c->vp->border_intensity, p).w; }

/* calcGradientAFD (:116-126) */
static v3 grad_forward(const vol_ctx* c, v3 p) {
    v3 o = c->dims_rcp;
    float v = tex_a(c, p);
    float v0 = tex_a(c, add3(p, V3(o.x, 0, 0))), v1 = tex_a(c, add3(p, V3(0, o.y, 0))), v2 = tex_a(c, add3(p, V3(0, 0, o.z)));
    v3 g = V3(v - v0, v - v1, v - v2);
    g = mul3(g, c->spacing_rcp);
    return scale3(g, c->vol->desc.rwm_scale);
}
/* calcGradientA (:202-214) */
static v3 grad_central(const vol_ctx* c, v3 p) {
    v3 o = c->dims_rcp;
    float v0 = tex_a(c, add3(p, V3(o.x, 0, 0))), v1 = tex_a(c, add3(p, V3(0, o.y, 0))), v2 = tex_a(c, add3(p, V3(0, 0, o.z)));
    float v3_ = tex_a(c, add3(p, V3(-o.x, 0, 0))), v4_ = tex_a(c, add3(p, V3(0, -o.y, 0))), v5 = tex_a(c, add3(p, V3(0, 0, -o.z)));
    v3 g = V3(v3_ - v0, v4_ - v1, v5 - v2);
    g = mul3(g, scale3(c->spacing_rcp, 0.5f));
    return scale3(g, c->vol->desc.rwm_scale);
}
/* calcGradientSobel (:224-280) */
static v3 grad_sobel(const vol_ctx* c, v3 p) {
    v3 o = c->dims_rcp;
    float s[3][3][3];   /* s[dx+1][dy+1][dz+1] */
    for (int i = -1; i <= 1; ++i) for (int j = -1; j <= 1; ++j) for (int k = -1; k <= 1; ++k) {
        if (i == 0 && j == 0 && k == 0) { s[1][1][1] = 0.f; continue; }
        s[i + 1][j + 1][k + 1] = tex_a(c, add3(p, V3(i * o.x, j * o.y, k * o.z)));
    }
#define S(i, j, k) s[(i) + 1][(j) + 1][(k) + 1]
    /* v01..v09: x=+1 ; v10..v18: x=0 ; v19..v27: x=-1 ; inside a block: rows z=+1,0,-1 ; columns y=-1,0,+1 */
    float v01 = S(1, -1, 1), v02 = S(1, 0, 1), v03 = S(1, 1, 1), v04 = S(1, -1, 0), v05 = S(1, 0, 0), v06 = S(1, 1, 0), v07 = S(1, -1, -1), v08 = S(1, 0, -1), v09 = S(1, 1, -1);
    float v10 = S(0, -1, 1), v11 = S(0, 0, 1), v12 = S(0, 1, 1), v13 = S(0, -1, 0), v15 = S(0, 1, 0), v16 = S(0, -1, -1), v17 = S(0, 0, -1), v18 = S(0, 1, -1);
    float v19 = S(-1, -1, 1), v20 = S(-1, 0, 1), v21 = S(-1, 1, 1), v22 = S(-1, -1, 0), v23 = S(-1, 0, 0), v24 = S(-1, 1, 0), v25 = S(-1, -1, -1), v26 = S(-1, 0, -1), v27 = S(-1, 1, -1);
#undef S
    float gx = v01 + 2.0f * v02 + v03 - v19 - 2.0f * v20 - v21
             + 2.0f * v04 + 4.0f * v05 + 2.0f * v06 - 2.0f * v22 - 4.0f * v23 - 2.0f * v24
             + v07 + 2.0f * v08 + v09 - v25 - 2.0f * v26 - v27;
    float gy = -v01 + v03 - 2.0f * v10 + 2.0f * v12 - v19 + v21
             - 2.0f * v04 + 2.0f * v06 - 4.0f * v13 + 4.0f * v15 - 2.0f * v22 + 2.0f * v24
             - v07 + v09 - 2.0f * v16 + 2.0f * v18 - v25 + v27;
    float gz = v01 + 2.0f * v02 + v03 + 2.0f * v10 + 4.0f * v11 + 2.0f * v12 + v19 + 2.0f * v20 + v21
             - v07 - 2.0f * v08 - v09 - 2.0f * v16 - 4.0f * v17 - 2.0f * v18 - v25 - 2.0f * v26 - v27;
    v3 g = V3(-gx, -gy, -gz);
    g = mul3(g, scale3(c->spacing_rcp, 0.5f));
    g = V3(g.x / 16.0f, g.y / 16.0f, g.z / 16.0f);
    return scale3(g, c->vol->desc.rwm_scale);
}
/* calcGradientFiltered (:291-305) */
static v3 grad_filtered(const vol_ctx* c, v3 p) {
    v3 d = c->dims_rcp;
    v3 g0 = grad_central(c, p);
    v3 g1 = grad_central(c, add3(p, V3(-d.x, -d.y, -d.z))), g2 = grad_central(c, add3(p, V3(d.x, d.y, d.z)));
    v3 g3 = grad_central(c, add3(p, V3(-d.x, d.y, -d.z))), g4 = grad_central(c, add3(p, V3(d.x, -d.y, d.z)));
    v3 g5 = grad_central(c, add3(p, V3(-d.x, -d.y, d.z))), g6 = grad_central(c, add3(p, V3(d.x, d.y, -d.z)));
    v3 g7 = grad_central(c, add3(p, V3(-d.x, d.y, d.z))), g8 = grad_central(c, add3(p, V3(d.x, -d.y, -d.z)));
    v3 mix0 = mix3(mix3(g1, g2, 0.5f), mix3(g3, g4, 0.5f), 0.5f);
    v3 mix1 = mix3(mix3(g5, g6, 0.5f), mix3(g7, g8, 0.5f), 0.5f);
    return mix3(g0, mix3(mix0, mix1, 0.5f), 0.75f);
}
/* CALC_GRADIENT dispatch (volumeraycaster.cpp:197-207); "none" = (voxel.xyz-0.5)*2 on the mapped voxel */
static v3 calc_gradient(int mode, const vol_ctx* c, v3 p, v4 voxel) {
    switch (mode) {
        case B200RC_GRAD_NONE: return V3((voxel.x - 0.5f) * 2.0f, (voxel.y - 0.5f) * 2.0f, (voxel.z - 0.5f) * 2.0f);
        case B200RC_GRAD_FORWARD: return grad_forward(c, p);
        case B200RC_GRAD_CENTRAL: return grad_central(c, p);
        case B200RC_GRAD_SOBEL: return grad_sobel(c, p);
        default: return grad_filtered(c, p);
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* shading (mod_shading.frag)                                                                                    */
/* ------------------------------------------------------------------------------------------------------------ */
static inline float attenuation(const b200rc_light* l, float d) {    /* :51-56 */
    float att = 1.0f / (l->attenuation[0] + l->attenuation[1] * d + l->attenuation[2] * d * d);
    return gl_min(att, 1.0f);
}
static inline v3 L3(const float* p) { return V3(p[0], p[1], p[2]); }
static inline v3 ambient_term(const b200rc_light* l, v3 ka) { return mul3(ka, L3(l->ambient)); }
static inline v3 diffuse_term(const b200rc_light* l, v3 kd, v3 N, v3 L) { return scale3(mul3(kd, L3(l->diffuse)), gl_max(dot3(N, L), 0.0f)); }
static inline v3 specular_term(const b200rc_light* l, v3 ks, v3 N, v3 L, v3 V, float alpha) {
    v3 H = normalize3(add3(V, L));
    float NdotH = powf(gl_max(dot3(N, H), 0.0f), alpha);
    return scale3(mul3(ks, L3(l->specular)), NdotH);
}

/* APPLY_SHADING(n, pos, lPos, cPos, ka, kd, ks) dispatch (glsl.cpp:411-437) */
static v3 apply_shading(int mode, const b200rc_light* l, v3 n, v3 pos, v3 lPos, v3 cPos, v3 ka, v3 kd, v3 ks) {
    switch (mode) {
    case B200RC_SHADE_NONE: return ka;
    case B200RC_SHADE_PHONG: case B200RC_SHADE_PHONG_DS: case B200RC_SHADE_PHONG_S:
    case B200RC_SHADE_PHONG_DA: case B200RC_SHADE_PHONG_D: {          /* :134-267 */
        v3 N = normalize3(n);
        v3 L = sub3(lPos, pos);
        v3 V = normalize3(sub3(cPos, pos));
        float d = length3(L);
        L = V3(L.x / d, L.y / d, L.z / d);
        v3 c = V3(0, 0, 0);
        if (mode == B200RC_SHADE_PHONG || mode == B200RC_SHADE_PHONG_DA) c = add3(c, ambient_term(l, ka));
        if (mode != B200RC_SHADE_PHONG_S) c = add3(c, diffuse_term(l, kd, N, L));
        if (mode == B200RC_SHADE_PHONG || mode == B200RC_SHADE_PHONG_DS || mode == B200RC_SHADE_PHONG_S)
            c = add3(c, specular_term(l, ks, N, L, V, l->shininess));
        if (l->apply_attenuation) c = scale3(c, attenuation(l, d));
        return c;
    }
    case B200RC_SHADE_TOON: {                                          /* :278-293, numShades = 3; L NOT normalised */
        v3 N = normalize3(n);
        v3 L = sub3(lPos, pos);
        float NdotL = gl_max(dot3(N, L), 0.0f);
        v3 k = mul3(kd, L3(l->diffuse));
        for (int i = 1; i <= 3; ++i)
            if (NdotL <= (float)i / 3.0f) { NdotL = (float)i / 3.0f; break; }
        return scale3(k, NdotL);
    }
    case B200RC_SHADE_COOK_TORRANCE: {                                 /* :307-369 */
        float m = 1.0f - (l->shininess / 128.0f);
        float s = (l->shininess / 128.0f);
        float d = 1.0f - s;
        float korr = 2.0f, f = 1.0f;
        v3 N = normalize3(n), L = normalize3(sub3(lPos, pos)), V = normalize3(sub3(cPos, pos));
        v3 H = normalize3(add3(V, L));
        float dw = length3(L) * 1.0f;
        float NH = dot3(N, H), NV = dot3(N, V), NL = gl_max(0.0000001f, dot3(N, L)), VH = dot3(V, H);
        float F = f + (1.0f - f) * powf(1.0f - NV, 5.0f);
        float G1 = (2.0f * NH * NL) / VH, G2 = (2.0f * NH * NV) / VH;
        float G = gl_min(1.0f, gl_max(0.0f, gl_min(G1, G2)));
        float mq = m * m + 1e-5f, NHq = NH * NH + 1e-5f;
        float D1 = 1.0f / (4.0f * mq * NHq * NHq);
        float D2 = expf(-(1.0f - NHq) / (mq * NHq));
        float D = D1 * D2;
        v3 Rd = scale3(mul3(kd, L3(l->diffuse)), korr * NL);
        v3 Rs = scale3(mul3(ks, L3(l->specular)), F * D * G / (3.1415926f * NV * NL));
        v3 c = ambient_term(l, ka);
        return add3(c, scale3(add3(scale3(Rs, s), scale3(Rd, d)), NL * dw));
    }
    case B200RC_SHADE_OREN_NAYAR: {                                    /* :381-412 */
        float m = l->shininess / 128.0f;
        v3 N = normalize3(n), L = normalize3(sub3(lPos, pos)), V = normalize3(sub3(cPos, pos));
        float VdotN = dot3(V, N), LdotN = dot3(L, N), NdotL = gl_max(0.0000001f, dot3(N, L));
        float cosPhi = dot3(sub3(V, scale3(N, VdotN)), sub3(L, scale3(N, LdotN)));
        float alpha = gl_max(acosf(VdotN), acosf(LdotN)), beta = gl_min(acosf(VdotN), acosf(LdotN));
        float mq = m * m;
        float A = 1.0f - 0.5f * (mq / (mq + 0.57f)), B = 0.45f * (mq / (mq + 0.09f));
        float res = (A + B * gl_max(0.0f, cosPhi) * sinf(alpha) * tanf(beta));
        v3 c = ambient_term(l, ka);
        return add3(c, scale3(mul3(kd, L3(l->diffuse)), NdotL * res));
    }
    default: {                                                         /* ward :426-481 */
        float alphaX = l->shininess / 128.0f, alphaY = alphaX * 2.0f;
        v3 N = normalize3(n), L = normalize3(sub3(lPos, pos)), V = normalize3(sub3(cPos, pos));
        v3 H = normalize3(add3(V, L));
        v3 X = normalize3(cross3(N, V3(1, 0, 0))), Y = normalize3(cross3(N, X));
        float VdotN = dot3(V, N), LdotN = dot3(L, N), HdotN = dot3(H, N), HdotX = dot3(H, X), HdotY = dot3(H, Y);
        float NdotL = gl_max(0.0000001f, dot3(N, L));
        alphaX += 0.0000001f; alphaY += 0.0000001f;
        v3 Rd = mul3(kd, L3(l->diffuse));
        float A = HdotX / alphaX; A *= A;
        float B = HdotY / alphaY; B *= B;
        float exponent = -2.0f * ((A + B) / (1.0f + HdotN));
        float nenner = 4.0f * 3.14159f; nenner *= alphaX; nenner *= alphaY; nenner *= sqrtf(LdotN * VdotN);
        if (nenner < 0.0f) nenner = gl_min(0.0000001f, nenner); else nenner = gl_max(0.0000001f, nenner);
        v3 Rs = scale3(mul3(ks, L3(l->specular)), expf(exponent) / nenner);
        v3 c = ambient_term(l, ka);
        return add3(c, scale3(add3(Rs, Rd), NdotL));
    }
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* classification (mod_transfunc.frag)                                                                           */
/* ------------------------------------------------------------------------------------------------------------ */
static inline float rw_to_tex(float lo, float hi, float v) {     /* realWorldToTexture :38-47 */
    if (v <= lo) return 0.0f;
    else if (v >= hi) return 1.0f;
    else if (hi == lo) return 0.0f;
    else return (v - lo) / (hi - lo);
}

/* sampler2D LINEAR, CLAMP_TO_EDGE, float texels */
static v4 tex2d_linear_f32(const float* t, int w, int h, float s, float tt) {
    float u = s * (float)w - 0.5f, v = tt * (float)h - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int x0 = (int)fu, y0 = (int)fv, x1 = x0 + 1, y1 = y0 + 1;
    if (x0 < 0) x0 = 0; if (x0 > w - 1) x0 = w - 1; if (x1 < 0) x1 = 0; if (x1 > w - 1) x1 = w - 1;
    if (y0 < 0) y0 = 0; if (y0 > h - 1) y0 = h - 1; if (y1 < 0) y1 = 0; if (y1 > h - 1) y1 = h - 1;
    const v4* p = (const v4*)t;
    v4 c0 = lerp4(p[(size_t)y0 * w + x0], p[(size_t)y0 * w + x1], a);
    v4 c1 = lerp4(p[(size_t)y1 * w + x0], p[(size_t)y1 * w + x1], a);
    return lerp4(c0, c1, b);
}
static v4 tex2d_linear_u8(const uint8_t* t, int w, int h, float s, float tt) {
    float u = s * (float)w - 0.5f, v = tt * (float)h - 0.5f;
    float fu = floorf(u), fv = floorf(v);
    float a = u - fu, b = v - fv;
    int x0 = (int)fu, y0 = (int)fv, x1 = x0 + 1, y1 = y0 + 1;
    if (x0 < 0) x0 = 0; if (x0 > w - 1) x0 = w - 1; if (x1 < 0) x1 = 0; if (x1 > w - 1) x1 = w - 1;
    if (y0 < 0) y0 = 0; if (y0 > h - 1) y0 = h - 1; if (y1 < 0) y1 = 0; if (y1 > h - 1) y1 = h - 1;
    const float k = 1.0f / 255.0f;
#define PX(x, y) V4(t[4 * ((size_t)(y) * w + (x))] * k, t[4 * ((size_t)(y) * w + (x)) + 1] * k, t[4 * ((size_t)(y) * w + (x)) + 2] * k, t[4 * ((size_t)(y) * w + (x)) + 3] * k)
    v4 c0 = lerp4(PX(x0, y0), PX(x1, y0), a);
    v4 c1 = lerp4(PX(x0, y1), PX(x1, y1), a);
#undef PX
    return lerp4(c0, c1, b);
}

/* RC_APPLY_CLASSIFICATION (classificationmodes.cpp:46-54): applyTF(vec4) :71-78, applyTFpi :97-106,
 * applyTF(sampler2D, vec4) :88-94.  `ig` = vec4(gradient, intensity). */
static v4 classify(int mode, const orc_classifier* c, v4 ig, float last) {
    if (mode == B200RC_CLASS_TF) {
        float s = rw_to_tex(c->domain[0], c->domain[1], ig.w);
        return tf1d_fetch(c->tf_rgba8, c->tf_width, s);
    } else if (mode == B200RC_CLASS_PREINT) {
        float s = rw_to_tex(c->domain[0], c->domain[1], ig.w);
        float p = rw_to_tex(c->domain[0], c->domain[1], last);
        return tex2d_linear_f32(c->preint, c->preint_res, c->preint_res, s, p);   /* vec2(intensity, prev) */
    } else {
        float gm = length3(V3(ig.x, ig.y, ig.z));
        float s = rw_to_tex(c->domain[0], c->domain[1], ig.w);
        float g = rw_to_tex(c->domain_y[0], c->domain_y[1], gm);
        return tex2d_linear_u8(c->tf2d_rgba8, c->tf2d_w, c->tf2d_h, s, g);
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* compositing (mod_compositing.frag)                                                                            */
/* ------------------------------------------------------------------------------------------------------------ */
#define SAMPLING_BASE_INTERVAL_RCP 200.0f     /* :40 */
#define ERT_OPACITY 0.95f                     /* mod_raysetup.frag:36 */

static v4 composite_dvr(v4 cur, v4 color, float t, float step, float* tDepth, int opacity_correction) {   /* :47-78 */
    if (opacity_correction) color.w = 1.0f - powf(1.0f - color.w, step * SAMPLING_BASE_INTERVAL_RCP);
    v4 r = cur;
    float k = (1.0f - r.w) * color.w;
    r.x = r.x + k * color.x; r.y = r.y + k * color.y; r.z = r.z + k * color.z;
    r.w = r.w + (1.0f - r.w) * color.w;
    if (*tDepth < 0.0f) *tDepth = t;
    return r;
}
static v4 composite_mip(v4 cur, float voxel, v4 color, float t, float* tDepth, float* maxI) {          /* :85-95 */
    if (voxel > *maxI) { *maxI = voxel; *tDepth = t; return color; }
    return cur;
}
static v4 composite_mop(v4 cur, v4 color, float t, float* tDepth) {                                     /* :97-106 */
    if (color.w > cur.w) { *tDepth = t; return color; }
    return cur;
}
static v4 composite_mida_helper(v4 cur, float voxel, v4 color, float* maxI, float t, float* tDepth, float gamma) {   /* :109-127 */
    v4 r = cur;
    float delta = 0.0f;
    if (voxel > *maxI) delta = voxel - *maxI;
    float beta = 1.0f - delta * (1.0f + gamma);
    float k = (1.0f - beta * r.w) * color.w;
    r.x = beta * r.x + k * color.x; r.y = beta * r.y + k * color.y; r.z = beta * r.z + k * color.z;
    r.w = beta * r.w + (1.0f - beta * r.w) * color.w;
    if (*tDepth < 0.0f) *tDepth = t;
    return r;
}
static v4 composite_mida(v4 cur, float voxel, v4 color, float* maxI, float t, float step, float* tDepth, float gamma) {  /* :134-149 */
    color.w = 1.0f - powf(1.0f - color.w, step * SAMPLING_BASE_INTERVAL_RCP);
    v4 r = cur;
    if (gamma <= 0.0f) r = composite_mida_helper(r, voxel, color, maxI, t, tDepth, gamma);
    else {
        /* note: compositeMIP updates maxIntensity BEFORE the helper reads it (inout), so delta_i is 0 in the helper */
        v4 mip = composite_mip(r, voxel, color, t, tDepth, maxI);
        v4 mida = composite_mida_helper(r, voxel, color, maxI, t, tDepth, 0.0f);
        r = V4(gamma * mip.x + (1.0f - gamma) * mida.x, gamma * mip.y + (1.0f - gamma) * mida.y,
               gamma * mip.z + (1.0f - gamma) * mida.z, gamma * mip.w + (1.0f - gamma) * mida.w);
    }
    return r;
}
static v4 composite_iso(v4 cur, v4 color, float t, float* tDepth, float iso) {                          /* :156-166 */
    const float eps = 0.02f;
    if (color.w >= iso - eps && color.w <= iso + eps) { v4 r = color; r.w = 1.0f; *tDepth = t; return r; }
    return cur;
}
static v4 composite_fhp(v3 pos, v4 cur, float t, float* tDepth) {                                       /* :171-182 */
    if (cur.x == 0.f && cur.y == 0.f && cur.z == 0.f) {
        if (*tDepth < 0.0f) *tDepth = t;
        return V4(pos.x, pos.y, pos.z, 1.0f);
    }
    return cur;
}
static v4 composite_fhn(v3 g, v4 cur, float t, float* tDepth) {                                         /* :187-198 */
    if (cur.x == 0.f && cur.y == 0.f && cur.z == 0.f) {
        v3 n = normalize3(g);
        if (*tDepth < 0.0f) *tDepth = t;
        return V4(n.x / 2.0f + 0.5f, n.y / 2.0f + 0.5f, n.z / 2.0f + 0.5f, 1.0f);
    }
    return cur;
}

/* RC_APPLY_COMPOSITINGn macro as generated by VolumeRaycaster::generateCompositionHeader (volumeraycaster.cpp:250-272) */
static v4 apply_compositing(int mode, int connected, int preint, v4 result, float voxel, v4 color, v3 samplePos,
                            v3 gradient, float t, float step, float* tDepth, float* maxI, float gamma, float iso) {
    if (!connected) return V4(0, 0, 0, 1);
    switch (mode) {
        case B200RC_COMP_DVR: return composite_dvr(result, color, t, step, tDepth, !preint);
        case B200RC_COMP_MIP: return composite_mip(result, voxel, color, t, tDepth, maxI);
        case B200RC_COMP_MOP: return composite_mop(result, color, t, tDepth);
        case B200RC_COMP_MIDA: return composite_mida(result, voxel, color, maxI, t, step, tDepth, gamma);
        case B200RC_COMP_ISO: return composite_iso(result, color, t, tDepth, iso);
        case B200RC_COMP_FHP: return composite_fhp(samplePos, result, t, tDepth);
        default: return composite_fhn(gradient, result, t, tDepth);
    }
}

/* mod_depth.frag:35-83 with the constants of volumeraycaster.cpp:280-287 */
static float get_depth_value(float t, float tEnd, float entryDepth, float exitDepth, float n, float f) {
    if (!(t >= 0.0f)) return 1.0f;
    float const_to_z_e_1 = 0.5f + 0.5f * ((f + n) / (f - n));
    float const_to_z_e_2 = ((f - n) / (f * n));
    float const_to_z_w_1 = ((f * n) / (f - n));
    float const_to_z_w_2 = 0.5f * ((f + n) / (f - n)) + 0.5f;
    float tt = t / tEnd;
    float ze_front = 1.0f / ((entryDepth - const_to_z_e_1) * const_to_z_e_2);
    float ze_back = 1.0f / ((exitDepth - const_to_z_e_1) * const_to_z_e_2);
    float ze_current = ze_front + tt * (ze_back - ze_front);
    return (1.0f / ze_current) * const_to_z_w_1 + const_to_z_w_2;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* the single-volume ray marcher (rc_singlevolume.frag)                                                          */
/* ------------------------------------------------------------------------------------------------------------ */
static inline v4 entry_fetch(const float* img, int w, int h, int x, int y) {    /* GL_NEAREST, CLAMP_TO_EDGE */
    if (x < 0) x = 0; if (x > w - 1) x = w - 1;
    if (y < 0) y = 0; if (y > h - 1) y = h - 1;
    const float* p = img + 4 * ((size_t)y * w + x);
    return V4(p[0], p[1], p[2], p[3]);
}

/* fixClipBorderGradient (mod_gradients.frag:39-46) */
static v3 fix_clip_border_gradient(v3 samplePos, v3 rayDir, const float* entry, int w, int h, int px, int py) {
    /* vec2(gl_FragCoord.x+2, y) * screenDimRCP_, NEAREST: texel floor(((px+0.5)+2)/w * w) = px+2 (clamped) */
    v4 e0 = entry_fetch(entry, w, h, px + 2, py), e1 = entry_fetch(entry, w, h, px, py + 2);
    v3 v0 = normalize3(sub3(V3(e0.x, e0.y, e0.z), samplePos));
    v3 v1 = normalize3(sub3(V3(e1.x, e1.y, e1.z), samplePos));
    v3 g = cross3(v0, v1);
    /* faceforward(N, I, Nref) = dot(Nref, I) < 0 ? N : -N */
    if (!(dot3(g, rayDir) < 0.0f)) g = V3(-g.x, -g.y, -g.z);
    return g;
}

static inline float comp4(v4 v, int i) { return i == 0 ? v.x : i == 1 ? v.y : i == 2 ? v.z : v.w; }

static void vol_ctx_init(vol_ctx* c, const orc_volume* vol, const b200rc_volume_params* vp) {
    c->vol = vol; c->vp = vp;
    const b200rc_volume_desc* d = &vol->desc;
    c->dims = V3((float)d->dims[0], (float)d->dims[1], (float)d->dims[2]);
    c->dims_rcp = V3(1.f / c->dims.x, 1.f / c->dims.y, 1.f / c->dims.z);       /* glsl.cpp:44 */
    c->spacing = V3(d->spacing[0], d->spacing[1], d->spacing[2]);
    c->spacing_rcp = V3(1.f / c->spacing.x, 1.f / c->spacing.y, 1.f / c->spacing.z);
    c->offset = V3(d->offset[0], d->offset[1], d->offset[2]);
}
/* texToPhysical (mod_sampler3d.frag:179-181) */
static inline v3 tex_to_physical(const vol_ctx* c, v3 p) {
    return add3(mul3(sub3(mul3(p, c->dims), V3(0.5f, 0.5f, 0.5f)), c->spacing), c->offset);
}

typedef struct { uint64_t rays, classified, shaded, nominal; } loc_stats;

/* rgb2hsv(...).r of rc_rgb.frag:44-63 */
static float rgb_hue(v3 c) {
    float mn = gl_min(c.x, gl_min(c.y, c.z)), mx = gl_max(c.x, gl_max(c.y, c.z));
    float delta = mx - mn, sat = (mx != 0.0f) ? delta / mx : 0.0f, h = 0.0f;
    if (sat == 0.0f) return 0.0f;
    if (c.x == mx) h = (c.y - c.z) / delta;
    else if (c.y == mx) h = 2 + (c.z - c.x) / delta;
    else if (c.z == mx) h = 4 + (c.x - c.y) / delta;
    h *= 60.0f;
    if (h < 0.0f) h += 360.0f;
    return h / 360.0f;
}

/* SimpleRaycaster (V/modules/base/glsl/rc_simple.frag:50-122; host side simpleraycaster.cpp:71-150) and
 * RGBRaycaster (rc_rgb.frag:65-140, rgbraycaster.cpp:63-125): same loop, different sample colour */
static void raycast_simple(const orc_volume* vol, const orc_classifier* cls, const b200rc_params* P, const b200rc_camera* cam, int W, int H,
                           const float* entry, const float* entryD, const float* exitp, const float* exitD,
                           float* out_color[3], float* out_depth, b200rc_stats* stats) {
    const float step = P->sampling_step;
    uint64_t s_rays = 0, s_class = 0, s_shaded = 0, s_nominal = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : s_rays, s_class, s_shaded, s_nominal)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            size_t pix = (size_t)py * W + px;
            v3 first = V3(entry[4 * pix], entry[4 * pix + 1], entry[4 * pix + 2]);
            v3 last = V3(exitp[4 * pix], exitp[4 * pix + 1], exitp[4 * pix + 2]);
            v4 result = V4(0, 0, 0, 0);
            float depth = 1.0f;
            if (!(first.x == last.x && first.y == last.y && first.z == last.z)) {      /* :113-116 */
                s_rays++;
                float depthT = -1.0f, t = 0.0f;
                v3 dir = sub3(last, first);
                float tend = length3(dir);
                dir = V3(dir.x / tend, dir.y / tend, dir.z / tend);
                s_nominal += (uint64_t)floorf(tend / step) + 1u;
                int finished = 0;
                for (int loop = 0; !finished && loop < 255 * 255; ++loop) {
                    v3 samplePos = add3(first, scale3(dir, t));
                    v4 color;
                    s_class++;
                    if (P->raycaster == B200RC_RAYCASTER_RGB) {
                        /* texture(volume_, sampleVal): no real-world mapping; alpha from the TF at the colour's hue (:84-91) */
                        v4 voxel = tex3d(vol, P->vol[0].filter, P->vol[0].clamp, P->vol[0].border_intensity, samplePos);
                        v4 tfColor = classify(B200RC_CLASS_TF, &cls[0], V4(0, 0, 0, rgb_hue(V3(voxel.x, voxel.y, voxel.z))), 0.f);
                        color = voxel;
                        if (P->rgb_color_modulation) { color.x *= tfColor.x; color.y *= tfColor.y; color.z *= tfColor.z; }
                        color.w = tfColor.w;
                    } else {
                        v4 voxel = get_voxel(vol, &P->vol[0], samplePos);
                        color = classify(B200RC_CLASS_TF, &cls[0], V4(0, 0, 0, voxel.w), 0.f);      /* applyTF(tf, tex, vec4 voxel): voxel.a */
                    }
                    if (color.w > 0.0f) {
                        s_shaded++;
                        color.w = 1.0f - powf(1.0f - color.w, step * 200.0f);
                        result.x = result.x + (1.0f - result.w) * color.w * color.x;
                        result.y = result.y + (1.0f - result.w) * color.w * color.y;
                        result.z = result.z + (1.0f - result.w) * color.w * color.z;
                        result.w = result.w + (1.0f - result.w) * color.w;
                    }
                    if (depthT < 0.0f && result.w > 0.0f) depthT = t;
                    if (result.w >= 1.0f) { result.w = 1.0f; finished = 1; }
                    t += step;
                    finished = finished || (t > tend);
                }
                if (depthT >= 0.0f) depth = get_depth_value(depthT, tend, entryD[pix], exitD[pix], cam->near_dist, cam->far_dist);
            }
            if (out_color[0]) {
                float* d = out_color[0] + 4 * pix;
                v4 r = result;
                if (P->out_format == B200RC_OUT_RGBA16) r = V4(orc_unorm16(r.x), orc_unorm16(r.y), orc_unorm16(r.z), orc_unorm16(r.w));
                d[0] = r.x; d[1] = r.y; d[2] = r.z; d[3] = r.w;
            }
            if (out_depth) out_depth[pix] = P->quantize_depth24 ? orc_depth24(depth) : depth;
        }
    if (stats) { stats->rays_cast = s_rays; stats->samples_classified = s_class; stats->samples_shaded = s_shaded; stats->samples_nominal = s_nominal; }
}

static void raycast_single(const orc_volume* vol, const orc_classifier* cls, const b200rc_params* P,
                           const b200rc_camera* cam, int W, int H,
                           const float* entry, const float* entryD, const float* exitp, const float* exitD,
                           float* out_color[3], float* out_depth, b200rc_stats* stats) {
    vol_ctx vc; vol_ctx_init(&vc, vol, &P->vol[0]);
    const int nch = vol->desc.channels;
    const int cmode = P->vol[0].classification;
    const int preint = (cmode == B200RC_CLASS_PREINT);
    const int requires_gradient = (cmode == B200RC_CLASS_TF2D);   /* CLASSIFICATION_REQUIRES_GRADIENT */
    const float step = P->sampling_step;
    /* SegmentationRaycaster (rc_segmentation.frag:104-160): the same loop with ka = kd = ks = colour (:137), early ray termination on
     * result.a alone (:145) and, with applySegmentation, the per-segment classification of :67-98 */
    const int seg = (P->raycaster == B200RC_RAYCASTER_SEGMENTATION), seg_apply = seg && P->apply_segmentation;
    const float seg_scale = P->segmentation_bit_depth == 12 ? 4095.0f : (P->segmentation_bit_depth == 16 ? 65535.0f : 255.0f);
    /* USE_EARLY_TERMINATION (volumeraycaster.cpp:225-233, usePortgroup_ = true) */
    int ert = 1;
    for (int o = 0; o < 3; ++o) {
        int m = P->compositing[o];
        int can = ((m != B200RC_COMP_MIP) && (m != B200RC_COMP_MOP) && (m != B200RC_COMP_MIDA)) || !P->output_connected[o];
        ert = ert && can;
    }
    /* SEPERATE_CHANNEL_COMPOSITION_CHANNEL_OUTPUTn (singlevolumeraycaster.cpp:450-458): tested on the mode only */
    int separate[3];
    for (int o = 0; o < 3; ++o) separate[o] = (P->compositing[o] == B200RC_COMP_MIP);
    /* camera / light in physical space (glsl.cpp:62-72) */
    float w2p[16]; orc_world_to_physical(&vol->desc, w2p);
    v3 camPhys = mat4_mul_point(w2p, V3(cam->position[0], cam->position[1], cam->position[2]));
    v4 lp = mat4_mul_v4(w2p, V4(P->light.position[0], P->light.position[1], P->light.position[2], P->light.position[3]));
    v3 lightPhys = V3(lp.x, lp.y, lp.z);
    uint64_t s_rays = 0, s_class = 0, s_shaded = 0, s_nominal = 0;

#pragma omp parallel for schedule(dynamic, 2) reduction(+ : s_rays, s_class, s_shaded, s_nominal)
    for (int py = 0; py < H; ++py) {
        for (int px = 0; px < W; ++px) {
            size_t pix = (size_t)py * W + px;
            v3 first = V3(entry[4 * pix], entry[4 * pix + 1], entry[4 * pix + 2]);
            v3 last = V3(exitp[4 * pix], exitp[4 * pix + 1], exitp[4 * pix + 2]);
            v4 result[3][4];    /* [output][channel accumulator]; accumulator 0 is the shared one */
            memset(result, 0, sizeof result);
            float depth = 1.0f;             /* glClearDepth(1) */
            int cast = !(first.x == last.x && first.y == last.y && first.z == last.z);     /* :323-325 discard */
            if (cast) {
                s_rays++;
                float t = 0.0f, tDepth = -1.0f;
                v4 lastIntensity = V4(0, 0, 0, 0);
                float maxI[3][4]; memset(maxI, 0, sizeof maxI);
                /* raySetup (mod_raysetup.frag:42-49) */
                v3 dir = sub3(last, first);
                float tEnd = length3(dir);
                dir = V3(dir.x / tEnd, dir.y / tEnd, dir.z / tEnd);
                s_nominal += (uint64_t)floorf(tEnd / step) + 1u;
                int finished = 0;
                for (int loop = 0; !finished && loop < 256 * 256; ++loop) {     /* WHILE, glsl.cpp:365 */
                    v3 samplePos = add3(first, scale3(dir, t));
                    v4 voxel;
                    if (!P->enable_channel_shift) voxel = get_voxel(vol, &P->vol[0], samplePos);
                    else {      /* getVoxelChannelShift :96-113 */
                        voxel = V4(0, 0, 0, 0);
                        voxel.x = get_voxel(vol, &P->vol[0], add3(samplePos, L3(P->channel_shift[0]))).x;
                        if (nch >= 2) voxel.y = get_voxel(vol, &P->vol[0], add3(samplePos, L3(P->channel_shift[1]))).y;
                        if (nch >= 3) voxel.z = get_voxel(vol, &P->vol[0], add3(samplePos, L3(P->channel_shift[2]))).z;
                        if (nch >= 4) voxel.w = get_voxel(vol, &P->vol[0], add3(samplePos, L3(P->channel_shift[3]))).w;
                    }
                    v3 gradient = V3(0, 0, 0);
                    if (requires_gradient) gradient = calc_gradient(P->gradient, &vc, samplePos, voxel);
                    s_class++;
                    v4 color[4];
                    int hasOpacity = 0;
                    if (seg_apply) {
                        float segValue = tex3d(&vol[1], B200RC_FILTER_NEAREST, B200RC_CLAMP_TO_EDGE, 0.f, samplePos).w;
                        float segment = gl_min(segValue * seg_scale, 255.0f);
                        int index = (int)segment;
                        float lo = cls[1].seg_domains[2 * index], hi = cls[1].seg_domains[2 * index + 1], dv;
                        if (voxel.w <= lo) dv = 0.0f;
                        else if (voxel.w >= hi) dv = 1.0f;
                        else dv = (voxel.w - lo) / (hi - lo);
                        color[0] = tex2d_linear_u8(cls[1].tf2d_rgba8, cls[1].tf2d_w, cls[1].tf2d_h, dv, (segment * 3.0f + 1.5f) / (float)cls[1].tf2d_h);
                        hasOpacity = color[0].w > 0.f;
                    } else
                    for (int c = 0; c < nch; ++c) {
                        color[c] = classify(cmode, &cls[c], V4(gradient.x, gradient.y, gradient.z, comp4(voxel, c)), comp4(lastIntensity, c));
                        hasOpacity = hasOpacity || (color[c].w > 0.f);
                    }
                    if (hasOpacity) {
                        s_shaded++;
                        if (t == 0.0f) gradient = fix_clip_border_gradient(samplePos, dir, entry, W, H, px, py);
                        else if (!requires_gradient) gradient = calc_gradient(P->gradient, &vc, samplePos, voxel);
                        v3 posPhys = tex_to_physical(&vc, samplePos);
                        for (int c = 0; c < nch; ++c) {
                            /* quirk (:245,251): channels 3 and 4 write the shaded colour into color2 -> composited unshaded */
                            v3 rgb = V3(color[c].x, color[c].y, color[c].z);
                            v3 shaded = apply_shading(P->vol[0].shading, &P->light, gradient, posPhys, lightPhys, camPhys, rgb, rgb, seg ? rgb : V3(1, 1, 1));
                            if (c < 2) { color[c].x = shaded.x; color[c].y = shaded.y; color[c].z = shaded.z; }
                            else { color[1].x = shaded.x; color[1].y = shaded.y; color[1].z = shaded.z; }
                            for (int o = 0; o < 3; ++o) {
                                v4* acc = separate[o] ? &result[o][c] : &result[o][0];
                                *acc = apply_compositing(P->compositing[o], P->output_connected[o], preint, *acc, comp4(voxel, c), color[c],
                                                         samplePos, gradient, t, step, &tDepth, &maxI[o][c], P->mida_gamma[o], P->iso_value);
                            }
                        }
                    }
                    lastIntensity = voxel;
                    if (ert) {      /* earlyRayTermination (mod_raysetup.frag:65-74) on RESULTn_1.a */
                        float a1 = result[0][0].w, a2 = result[1][0].w, a3 = result[2][0].w;
                        if (seg) {      /* the two-argument overload (mod_raysetup.frag:56-63) */
                            if (a1 >= ERT_OPACITY) { result[0][0].w = 1.0f; finished = 1; }
                        } else if (gl_min(a1, gl_min(a2, a3)) >= ERT_OPACITY) {
                            result[0][0].w = 1.0f; result[1][0].w = 1.0f; result[2][0].w = 1.0f;
                            finished = 1;
                        }
                    }
                    t += step;
                    finished = finished || (t > tEnd);
                }
                /* :272-309 add per-channel results of separately composited outputs */
                for (int o = 0; o < 3; ++o)
                    if (separate[o])
                        for (int c = 1; c < nch; ++c) {
                            result[o][0].x += result[o][c].x; result[o][0].y += result[o][c].y;
                            result[o][0].z += result[o][c].z; result[o][0].w += result[o][c].w;
                        }
                depth = get_depth_value(tDepth, tEnd, entryD[pix], exitD[pix], cam->near_dist, cam->far_dist);
            }
            for (int o = 0; o < 3; ++o)
                if (out_color[o]) {
                    float* d = out_color[o] + 4 * pix;
                    v4 r = result[o][0];
                    if (P->out_format == B200RC_OUT_RGBA16) r = V4(orc_unorm16(r.x), orc_unorm16(r.y), orc_unorm16(r.z), orc_unorm16(r.w));
                    d[0] = r.x; d[1] = r.y; d[2] = r.z; d[3] = r.w;
                }
            if (out_depth) out_depth[pix] = P->quantize_depth24 ? orc_depth24(depth) : depth;
        }
    }
    if (stats) { stats->rays_cast = s_rays; stats->samples_classified = s_class; stats->samples_shaded = s_shaded; stats->samples_nominal = s_nominal; }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* the multi-volume ray marcher (rc_multivolume.frag:97-350)                                                     */
/* ------------------------------------------------------------------------------------------------------------ */
static void raycast_multi(const orc_volume* vols, const orc_classifier* cls, const b200rc_params* P,
                          const b200rc_camera* cam, int W, int H,
                          const float* entry, const float* entryD, const float* exitp, const float* exitD,
                          float* out_color[3], float* out_depth, b200rc_stats* stats) {
    const int nv = P->num_volumes;
    const size_t plane = (size_t)W * H;
    vol_ctx vc[B200RC_MAX_VOLUMES];
    float w2t[B200RC_MAX_VOLUMES][16];
    v3 camPhys[B200RC_MAX_VOLUMES], lightPhys[B200RC_MAX_VOLUMES];
    for (int v = 0; v < nv; ++v) {
        vol_ctx_init(&vc[v], &vols[v], &P->vol[v]);
        float t2w[16]; orc_texture_to_world(&vols[v].desc, t2w);
        orc_mat4_invert(t2w, w2t[v]);                       /* getWorldToTextureMatrix */
        float w2p[16]; orc_world_to_physical(&vols[v].desc, w2p);
        camPhys[v] = mat4_mul_point(w2p, V3(cam->position[0], cam->position[1], cam->position[2]));
        v4 lp = mat4_mul_v4(w2p, V4(P->light.position[0], P->light.position[1], P->light.position[2], P->light.position[3]));
        lightPhys[v] = V3(lp.x, lp.y, lp.z);
    }
    const float step = P->sampling_step;
    int ert = 1;
    for (int o = 0; o < 3; ++o) {
        int m = P->compositing[o];
        int can = ((m != B200RC_COMP_MIP) && (m != B200RC_COMP_MOP) && (m != B200RC_COMP_MIDA)) || !P->output_connected[o];
        ert = ert && can;
    }
    uint64_t s_rays = 0, s_class = 0, s_shaded = 0, s_nominal = 0;
#pragma omp parallel for schedule(dynamic, 2) reduction(+ : s_rays, s_class, s_shaded, s_nominal)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            size_t pix = (size_t)py * W + px;
            v3 first = V3(entry[4 * pix], entry[4 * pix + 1], entry[4 * pix + 2]);
            v3 last = V3(exitp[4 * pix], exitp[4 * pix + 1], exitp[4 * pix + 2]);
            v4 result[3]; memset(result, 0, sizeof result);
            float depth = 1.0f;
            if (!(first.x == last.x && first.y == last.y && first.z == last.z)) {
                s_rays++;
                float t = 0.0f, tDepth = -1.0f;
                v3 dir = sub3(last, first);
                float tEnd = length3(dir);
                dir = V3(dir.x / tEnd, dir.y / tEnd, dir.z / tEnd);
                s_nominal += (uint64_t)floorf(tEnd / step) + 1u;
                v3 vfirst[B200RC_MAX_VOLUMES], vdir[B200RC_MAX_VOLUMES];
                float vBegin[B200RC_MAX_VOLUMES], vEnd[B200RC_MAX_VOLUMES], lastI[B200RC_MAX_VOLUMES];
                for (int v = 0; v < nv; ++v) {
                    const float* ve = entry + 4 * ((size_t)(v + 1) * plane + pix);
                    const float* vx = exitp + 4 * ((size_t)(v + 1) * plane + pix);
                    vfirst[v] = mat4_mul_point(w2t[v], first);                           /* :119-122 */
                    vdir[v] = sub3(mat4_mul_point(w2t[v], last), vfirst[v]);
                    vBegin[v] = length3(sub3(V3(ve[0], ve[1], ve[2]), first)) / tEnd;
                    vEnd[v] = length3(sub3(V3(vx[0], vx[1], vx[2]), first)) / tEnd;
                    lastI[v] = 0.0f;
                }
                float maxI[3] = {0, 0, 0};
                int finished = 0;
                for (int loop = 0; !finished && loop < 256 * 256; ++loop) {
                    float realT = t / tEnd;
                    v3 worldPos = add3(first, scale3(dir, t));
                    for (int v = 0; v < nv; ++v) {
                        v3 sp = add3(vfirst[v], scale3(vdir[v], realT));
                        if (realT >= vBegin[v] && realT <= vEnd[v]) {
                            v4 voxel = get_voxel(&vols[v], &P->vol[v], sp);
                            int cm = P->vol[v].classification;
                            int reqg = (cm == B200RC_CLASS_TF2D);
                            if (reqg) { v3 g = calc_gradient(P->gradient, &vc[v], sp, voxel); voxel.x = g.x; voxel.y = g.y; voxel.z = g.z; }
                            s_class++;
                            v4 color = classify(cm, &cls[v], voxel, lastI[v]);
                            if (color.w > 0.0f) {
                                s_shaded++;
                                if (!reqg) { v3 g = calc_gradient(P->gradient, &vc[v], sp, voxel); voxel.x = g.x; voxel.y = g.y; voxel.z = g.z; }
                                v3 g = V3(voxel.x, voxel.y, voxel.z);
                                v3 rgb = V3(color.x, color.y, color.z);
                                v3 sh = apply_shading(P->vol[v].shading, &P->light, g, tex_to_physical(&vc[v], sp), lightPhys[v], camPhys[v], rgb, rgb, V3(1, 1, 1));
                                color.x = sh.x; color.y = sh.y; color.z = sh.z;
                                /* mvOpacityCorrection (multivolumeraycaster.cpp:447-456) */
                                v4 cc = color; cc.w = 1.0f - powf(1.0f - color.w, P->mv_opacity_correction_factor);
                                /* the compositing header is generated from classificationMode_ of volume 1
                                 * (VolumeRaycaster::generateHeader, volumeraycaster.cpp:216-221) */
                                int pre = (P->vol[0].classification == B200RC_CLASS_PREINT);
                                for (int o = 0; o < 3; ++o)     /* MIDA gamma and iso are literal 0 (:180-182) */
                                    result[o] = apply_compositing(P->compositing[o], P->output_connected[o], pre, result[o], voxel.w, cc,
                                                                  worldPos, g, t, step, &tDepth, &maxI[o], 0.0f, 0.0f);
                            }
                            lastI[v] = voxel.w;
                        }
                    }
                    if (ert && gl_min(result[0].w, gl_min(result[1].w, result[2].w)) >= ERT_OPACITY) {
                        result[0].w = result[1].w = result[2].w = 1.0f;
                        finished = 1;
                    }
                    t += step;
                    finished = finished || (t > tEnd);
                }
                depth = get_depth_value(tDepth, tEnd, entryD[pix], exitD[pix], cam->near_dist, cam->far_dist);
            }
            for (int o = 0; o < 3; ++o)
                if (out_color[o]) {
                    float* d = out_color[o] + 4 * pix;
                    v4 r = result[o];
                    if (P->out_format == B200RC_OUT_RGBA16) r = V4(orc_unorm16(r.x), orc_unorm16(r.y), orc_unorm16(r.z), orc_unorm16(r.w));
                    d[0] = r.x; d[1] = r.y; d[2] = r.z; d[3] = r.w;
                }
            if (out_depth) out_depth[pix] = P->quantize_depth24 ? orc_depth24(depth) : depth;
        }
    if (stats) { stats->rays_cast = s_rays; stats->samples_classified = s_class; stats->samples_shaded = s_shaded; stats->samples_nominal = s_nominal; }
}

void orc_raycast(const orc_volume* vols, const orc_classifier* cls, const b200rc_params* params,
                 const b200rc_camera* cam, int width, int height,
                 const float* entry_rgba, const float* entry_depth, const float* exit_rgba, const float* exit_depth,
                 float* out_color[B200RC_MAX_OUTPUTS], float* out_depth, b200rc_stats* stats) {
    if (!params->multi_volume && (params->raycaster == B200RC_RAYCASTER_SIMPLE || params->raycaster == B200RC_RAYCASTER_RGB))
        raycast_simple(vols, cls, params, cam, width, height, entry_rgba, entry_depth, exit_rgba, exit_depth, out_color, out_depth, stats);
    else if (params->multi_volume)
        raycast_multi(vols, cls, params, cam, width, height, entry_rgba, entry_depth, exit_rgba, exit_depth, out_color, out_depth, stats);
    else
        raycast_single(vols, cls, params, cam, width, height, entry_rgba, entry_depth, exit_rgba, exit_depth, out_color, out_depth, stats);
}

/* ------------------------------------------------------------------------------------------------------------ */
/* proxy geometry (optimizedproxygeometry.cpp)                                                                   */
/* ------------------------------------------------------------------------------------------------------------ */
static inline float voxel_normalized(const orc_volume* v, int x, int y, int z, int c) {   /* VolumeRAM::getVoxelNormalized */
    const int* d = v->desc.dims;
    int64_t idx = ((((int64_t)z * d[1] + y) * d[0] + x) * v->desc.channels) + c;
    switch (v->desc.dtype) {
        case B200RC_U8:  return (float)((const uint8_t*)v->data)[idx] / 255.0f;
        case B200RC_U16: return (float)((const uint16_t*)v->data)[idx] / 65535.0f;
        default:         return ((const float*)v->data)[idx];
    }
}

/* computeRegionStructure (:894-987) */
void orc_region_minmax(const orc_volume* vol, int step, int32_t cells[3], float* out) {
    const int* dim = vol->desc.dims;
    int nc = vol->desc.channels;
    for (int a = 0; a < 3; ++a) cells[a] = (int)ceilf((float)dim[a] / (float)step);
    float sc = vol->desc.rwm_scale, of = vol->desc.rwm_offset;
#pragma omp parallel for collapse(2) schedule(static)
    for (int cz = 0; cz < cells[2]; ++cz)
        for (int cy = 0; cy < cells[1]; ++cy)
            for (int cx = 0; cx < cells[0]; ++cx) {
                int pos[3] = {cx, cy, cz}, llf[3], urb[3];
                for (int a = 0; a < 3; ++a) {
                    llf[a] = pos[a] * step;
                    urb[a] = (pos[a] + 1) * step - 1;
                    if (urb[a] > dim[a] - 1) urb[a] = dim[a] - 1;
                    /* one additional voxel as overlap (:944-947) */
                    llf[a] = llf[a] - 1; if (llf[a] < 0) llf[a] = 0; if (llf[a] > dim[a] - 1) llf[a] = dim[a] - 1;
                    urb[a] = urb[a] + 1; if (urb[a] < 0) urb[a] = 0; if (urb[a] > dim[a] - 1) urb[a] = dim[a] - 1;
                }
                float mn[4] = {FLT_MAX, FLT_MAX, FLT_MAX, FLT_MAX};
                float mx[4] = {FLT_MIN, FLT_MIN, FLT_MIN, FLT_MIN};     /* numeric_limits<float>::min() (quirk: smallest positive) */
                for (int z = llf[2]; z <= urb[2]; ++z)
                    for (int y = llf[1]; y <= urb[1]; ++y)
                        for (int x = llf[0]; x <= urb[0]; ++x)
                            for (int c = 0; c < nc; ++c) {
                                float cur = voxel_normalized(vol, x, y, z, c) * sc + of;    /* normalizedToRealWorld */
                                mn[c] = cur < mn[c] ? cur : mn[c];
                                mx[c] = cur > mx[c] ? cur : mx[c];
                            }
                size_t cell = ((size_t)cz * cells[1] + cy) * cells[0] + cx;
                for (int c = 0; c < nc; ++c) { out[(cell * nc + c) * 2] = mn[c]; out[(cell * nc + c) * 2 + 1] = mx[c]; }
            }
}

/* TransFuncBase::realWorldToNormalized (transfuncbase.cpp:244-259) */
static inline float tf_rw_to_norm(float rw, float lo, float hi) {
    if (rw < lo) return 0.f;
    else if (rw > hi) return 1.f;
    else if (lo == hi) return 0.f;
    else return (rw - lo) / (hi - lo);
}

int orc_proxy_boxes(const orc_volume* vol, const orc_tf1d* tfs, int n_tfs, const b200rc_eep_params* eep,
                    float* out, int max_boxes) {
    const int* dim = vol->desc.dims;
    float coordLLF[3], coordURB[3];
    for (int a = 0; a < 3; ++a) {       /* processCube (:420-448) */
        coordLLF[a] = 0.f - 0.5f;
        coordURB[a] = (float)(dim[a] - 1) + 0.5f;
        if (eep->enable_clipping) {
            coordLLF[a] = fmaxf(coordLLF[a], eep->clip_llf[a] - 0.5f);
            coordURB[a] = fminf(coordURB[a], eep->clip_urb[a] + 0.5f);
        }
    }
    if (eep->proxy_mode == B200RC_PROXY_BBOX || n_tfs <= 0) {
        if (max_boxes < 1) return 0;
        /* getVoxelToTextureMatrix = scale(1/dim) * translate(0.5)  (volumebase.cpp:227-231) */
        for (int a = 0; a < 3; ++a) {
            out[a] = (1.0f / (float)dim[a]) * (coordLLF[a] + 0.5f);
            out[3 + a] = (1.0f / (float)dim[a]) * (coordURB[a] + 0.5f);
        }
        return 1;
    }
    /* region structure + pre-integration classification (:1014-1116) */
    int step = eep->proxy_cell_size < 1 ? 1 : eep->proxy_cell_size;
    int32_t cells[3];
    for (int a = 0; a < 3; ++a) cells[a] = (int)ceilf((float)dim[a] / (float)step);
    size_t ncell = (size_t)cells[0] * cells[1] * cells[2];
    int nc = vol->desc.channels;
    float* mm = (float*)malloc(sizeof(float) * ncell * nc * 2);
    orc_region_minmax(vol, step, cells, mm);
    float* pi[4] = {0, 0, 0, 0};
    for (int c = 0; c < nc && c < n_tfs; ++c) {
        pi[c] = (float*)malloc(sizeof(float) * 4 * 256 * 256);
        orc_preint_cpu(&tfs[c], 256, 1.f, 0, pi[c]);        /* getPreIntegrationTable(1.f, 256) (:1051) */
    }
    int n = 0;
    float bmin[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, bmax[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
    int defined = 0;
    for (int cz = 0; cz < cells[2]; ++cz)
        for (int cy = 0; cy < cells[1]; ++cy)
            for (int cx = 0; cx < cells[0]; ++cx) {
                size_t cell = ((size_t)cz * cells[1] + cy) * cells[0] + cx;
                int pos[3] = {cx, cy, cz};
                int nonempty = 0;
                for (int c = 0; c < nc && c < n_tfs && !nonempty; ++c) {
                    float mn = tf_rw_to_norm(mm[(cell * nc + c) * 2], tfs[c].domain[0], tfs[c].domain[1]);
                    float mx = tf_rw_to_norm(mm[(cell * nc + c) * 2 + 1], tfs[c].domain[0], tfs[c].domain[1]);
                    float cl[4];
                    orc_preint_classify(pi[c], 256, mn, mx, cl);
                    if (!(cl[3] <= 0.0001f * eep->proxy_threshold)) nonempty = 1;       /* isRegionEmptyPi (:878-880) */
                }
                if (!nonempty) continue;
                float lo[3], hi[3];
                for (int a = 0; a < 3; ++a) {
                    int l = pos[a] * step, u = (pos[a] + 1) * step - 1;
                    if (u > dim[a] - 1) u = dim[a] - 1;
                    lo[a] = (float)l - 0.5f; hi[a] = (float)u + 0.5f;
                }
                if (eep->proxy_mode == B200RC_PROXY_MINBBOX) {
                    for (int a = 0; a < 3; ++a) { bmin[a] = fminf(bmin[a], lo[a]); bmax[a] = fmaxf(bmax[a], hi[a]); }
                    defined = 1;
                } else if (n < max_boxes) {
                    /* addCubeMesh (:480-491): tex = (coord + 0.5) / dim */
                    int separated = 0;      /* checkBounds.intersects(cubeBounds) (:1205-1216; touching counts, bounds.h:162-170) */
                    for (int a = 0; a < 3; ++a) {
                        if (eep->enable_clipping) { lo[a] = fmaxf(lo[a], coordLLF[a]); hi[a] = fminf(hi[a], coordURB[a]); if (lo[a] > hi[a]) separated = 1; }
                        out[6 * n + a] = (lo[a] + 0.5f) / (float)dim[a];
                        out[6 * n + 3 + a] = (hi[a] + 0.5f) / (float)dim[a];
                    }
                    if (!separated) n++;
                }
            }
    if (eep->proxy_mode == B200RC_PROXY_MINBBOX && defined && max_boxes >= 1) {
        for (int a = 0; a < 3; ++a) {
            float lo = bmin[a], hi = bmax[a];
            if (eep->enable_clipping) { lo = fmaxf(lo, coordLLF[a]); hi = fminf(hi, coordURB[a]); }
            out[a] = (lo + 0.5f) / (float)dim[a];
            out[3 + a] = (hi + 0.5f) / (float)dim[a];
        }
        n = 1;
    }
    for (int c = 0; c < 4; ++c) free(pi[c]);
    free(mm);
    return n;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* entry / exit points                                                                                           */
/* ------------------------------------------------------------------------------------------------------------ */
/* GL_RGBA16F store + fetch */
static float half_round(float v) {
    /* round-to-nearest-even float -> half -> float, finite values only */
    union { float f; uint32_t u; } in; in.f = v;
    uint32_t sign = in.u & 0x80000000u;
    in.u &= 0x7fffffffu;
    float a = in.f;
    if (!(a == a)) return v;
    if (a >= 65520.f) { in.u = sign | 0x7f800000u; return in.f; }
    if (a < 6.103515625e-05f) {       /* subnormal half: multiples of 2^-24 */
        float q = nearbyintf(a * 16777216.f) / 16777216.f;
        in.f = q; in.u |= sign; return in.f;
    }
    int e; frexpf(a, &e);            /* a = m * 2^e, m in [0.5,1) -> 11 significant bits */
    float scale = ldexpf(1.f, 11 - e);
    float q = nearbyintf(a * scale) / scale;
    in.f = q; in.u |= sign; return in.f;
}

typedef struct { float tmin, tmax; int hit, axis_in, axis_out; float plane_in, plane_out; } slab_hit;

/* ray o + t*d against an axis-aligned box; the parameter is along near->far, t=0 on the near plane */
static slab_hit slab(v3 o, v3 d, const float* box) {
    slab_hit h; h.tmin = -INFINITY; h.tmax = INFINITY; h.hit = 1; h.axis_in = h.axis_out = -1; h.plane_in = h.plane_out = 0.f;
    const float oo[3] = {o.x, o.y, o.z}, dd[3] = {d.x, d.y, d.z};
    for (int a = 0; a < 3; ++a) {
        float lo = box[a], hi = box[3 + a];
        if (dd[a] == 0.f) { if (oo[a] < lo || oo[a] > hi) h.hit = 0; continue; }
        float inv = 1.0f / dd[a];
        float t1 = (lo - oo[a]) * inv, t2 = (hi - oo[a]) * inv;
        float tn = t1, tf = t2, pn = lo, pf = hi;
        if (t1 > t2) { tn = t2; tf = t1; pn = hi; pf = lo; }
        if (tn > h.tmin) { h.tmin = tn; h.axis_in = a; h.plane_in = pn; }
        if (tf < h.tmax) { h.tmax = tf; h.axis_out = a; h.plane_out = pf; }
    }
    if (!(h.tmin <= h.tmax)) h.hit = 0;
    return h;
}

static inline v3 point_on(v3 o, v3 d, float t, int axis, float plane) {
    v3 p = add3(o, scale3(d, t));
    if (axis == 0) p.x = plane; else if (axis == 1) p.y = plane; else if (axis == 2) p.z = plane;
    return p;
}

void orc_entry_exit(const float* boxes, int n_boxes, const float model[16], const b200rc_eep_params* eep,
                    const b200rc_camera* cam, int W, int H,
                    float* entry_rgba, float* entry_depth, float* exit_rgba, float* exit_depth) {
    float viewInv[16], projInv[16], modelInv[16], mv[16], mvp[16];
    orc_mat4_invert(cam->view, viewInv);
    orc_mat4_invert(cam->proj, projInv);
    orc_mat4_invert(model, modelInv);               /* worldToTexture_ (meshentryexitpoints.cpp:361-364) */
    orc_mat4_mul(cam->view, model, mv);
    orc_mat4_mul(cam->proj, mv, mvp);
    /* bounding box of the whole mesh (Geometry::getBoundingBox(false)) */
    float bb[6] = {FLT_MAX, FLT_MAX, FLT_MAX, -FLT_MAX, -FLT_MAX, -FLT_MAX};
    for (int b = 0; b < n_boxes; ++b)
        for (int a = 0; a < 3; ++a) { bb[a] = fminf(bb[a], boxes[6 * b + a]); bb[3 + a] = fmaxf(bb[3 + a], boxes[6 * b + 3 + a]); }
    float rcpW = 1.0f / (float)W, rcpH = 1.0f / (float)H;      /* screenDimRCP_ */

#pragma omp parallel for schedule(dynamic, 4)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            size_t pix = (size_t)py * W + px;
            float sx = ((float)px + 0.5f) * rcpW, sy = ((float)py + 0.5f) * rcpH;
            /* unproject near (z=-1) and far (z=+1) NDC points exactly as eep_insidevolume.frag:57-66 does */
            v4 wn = V4(2.0f * sx - 1.0f, 2.0f * sy - 1.0f, -1.0f, 1.0f), wf = V4(wn.x, wn.y, 1.0f, 1.0f);
            v4 tn = mat4_mul_v4(viewInv, mat4_mul_v4(projInv, wn)), tf = mat4_mul_v4(viewInv, mat4_mul_v4(projInv, wf));
            v4 rn = V4(tn.x / tn.w, tn.y / tn.w, tn.z / tn.w, tn.w / tn.w), rf = V4(tf.x / tf.w, tf.y / tf.w, tf.z / tf.w, tf.w / tf.w);
            v3 o, f;
            if (!eep->world_coords) { v4 a = mat4_mul_v4(modelInv, rn), b = mat4_mul_v4(modelInv, rf); o = V3(a.x, a.y, a.z); f = V3(b.x, b.y, b.z); }
            else { o = V3(rn.x, rn.y, rn.z); f = V3(rf.x, rf.y, rf.z); }
            /* In world mode the boxes are still axis aligned in TEXTURE space: intersect there, emit world positions. */
            v3 ot = o, ft = f;
            if (eep->world_coords) { v4 a = mat4_mul_v4(modelInv, rn), b = mat4_mul_v4(modelInv, rf); ot = V3(a.x, a.y, a.z); ft = V3(b.x, b.y, b.z); }
            v3 d = sub3(ft, ot);
            /* rasterisation: nearest front face, nearest back face (GL_LESS), farthest back face (GL_GREATER); near/far clip */
            float tFront = INFINITY, tBack1 = INFINITY, tBackFar = -INFINITY;
            int aFront = -1, aBack1 = -1, aBackFar = -1; float pFront = 0, pBack1 = 0, pBackFar = 0;
            for (int b = 0; b < n_boxes; ++b) {
                slab_hit h = slab(ot, d, boxes + 6 * b);
                if (!h.hit) continue;
                if (h.tmin >= 0.f && h.tmin <= 1.f && h.tmin < tFront) { tFront = h.tmin; aFront = h.axis_in; pFront = h.plane_in; }
                if (h.tmax >= 0.f && h.tmax <= 1.f) {
                    if (h.tmax < tBack1) { tBack1 = h.tmax; aBack1 = h.axis_out; pBack1 = h.plane_out; }
                    if (h.tmax > tBackFar) { tBackFar = h.tmax; aBackFar = h.axis_out; pBackFar = h.plane_out; }
                }
            }
            (void)aBack1; (void)pBack1;
            v4 ecol = V4(0, 0, 0, 0), xcol = V4(0, 0, 0, 0);
            float edep = 1.0f, xdep = 0.0f;     /* exit pass clears depth to 0 (GL_GREATER) */
            /* positions and window depths */
            v3 pF = V3(0, 0, 0), pX = V3(0, 0, 0); float dF = 1.f, dX = 0.f;
            int haveFront = isfinite(tFront), haveBack = isfinite(tBackFar) && tBackFar >= 0.f, haveBack1 = isfinite(tBack1);
            if (haveFront) {
                pF = point_on(ot, d, tFront, aFront, pFront);
                v4 c = mat4_mul_v4(mvp, V4(pF.x, pF.y, pF.z, 1.f)); dF = 0.5f * (c.z / c.w) + 0.5f;
            }
            if (haveBack) {
                pX = point_on(ot, d, tBackFar, aBackFar, pBackFar);
                v4 c = mat4_mul_v4(mvp, V4(pX.x, pX.y, pX.z, 1.f)); dX = 0.5f * (c.z / c.w) + 0.5f;
            }
            if (eep->world_coords) {    /* colour = viewMatrixInverse * modelView * vertex = world position (eep_simple.vert:35-38) */
                if (haveFront) pF = mat4_mul_point(model, pF);
                if (haveBack) pX = mat4_mul_point(model, pX);
            }
            if (eep->inside_technique == B200RC_INSIDE_GPU) {       /* fillEntryPoints / eep_insidevolume.frag:44-83 */
                if (!haveBack1) { ecol = V4(0, 0, 0, 0); edep = 1.0f; }
                else {
                    v3 r = eep->world_coords ? o : ot;     /* near-plane point in output coordinates */
                    const float* llf = bb; const float* urb = bb + 3;
                    float wbb[6];
                    if (eep->world_coords) {    /* bounding box in world coords: transform the 8 corners */
                        for (int a = 0; a < 3; ++a) { wbb[a] = FLT_MAX; wbb[3 + a] = -FLT_MAX; }
                        for (int k = 0; k < 8; ++k) {
                            v3 c = mat4_mul_point(model, V3(bb[(k & 1) ? 3 : 0], bb[(k & 2) ? 4 : 1], bb[(k & 4) ? 5 : 2]));
                            wbb[0] = fminf(wbb[0], c.x); wbb[1] = fminf(wbb[1], c.y); wbb[2] = fminf(wbb[2], c.z);
                            wbb[3] = fmaxf(wbb[3], c.x); wbb[4] = fmaxf(wbb[4], c.y); wbb[5] = fmaxf(wbb[5], c.z);
                        }
                        llf = wbb; urb = wbb + 3;
                    }
                    int outside = (r.x > urb[0] || r.y > urb[1] || r.z > urb[2] || r.x < llf[0] || r.y < llf[1] || r.z < llf[2]);
                    if (outside) { if (haveFront) { ecol = V4(pF.x, pF.y, pF.z, 1.f); edep = dF; } }
                    else { ecol = V4(r.x, r.y, r.z, 1.0f); edep = 0.0f; }
                }
            } else if (haveFront) { ecol = V4(pF.x, pF.y, pF.z, 1.f); edep = dF; }
            if (haveBack) { xcol = V4(pX.x, pX.y, pX.z, 1.f); xdep = dX; }
            /* render-target formats: GL_RGBA16 (texture coords) / GL_RGBA16F (world) + DEPTH24 (meshentryexitpoints.cpp:135-148) */
            if (!eep->world_coords) {
                ecol = V4(orc_unorm16(ecol.x), orc_unorm16(ecol.y), orc_unorm16(ecol.z), orc_unorm16(ecol.w));
                xcol = V4(orc_unorm16(xcol.x), orc_unorm16(xcol.y), orc_unorm16(xcol.z), orc_unorm16(xcol.w));
            } else {
                ecol = V4(half_round(ecol.x), half_round(ecol.y), half_round(ecol.z), half_round(ecol.w));
                xcol = V4(half_round(xcol.x), half_round(xcol.y), half_round(xcol.z), half_round(xcol.w));
            }
            edep = orc_depth24(edep); xdep = orc_depth24(xdep);
            if (entry_rgba) { float* p = entry_rgba + 4 * pix; p[0] = ecol.x; p[1] = ecol.y; p[2] = ecol.z; p[3] = ecol.w; }
            if (exit_rgba) { float* p = exit_rgba + 4 * pix; p[0] = xcol.x; p[1] = xcol.y; p[2] = xcol.z; p[3] = xcol.w; }
            if (entry_depth) entry_depth[pix] = edep;
            if (exit_depth) exit_depth[pix] = xdep;
        }

    /* jitterEntryPoints / eep_jitter.frag:38-78 (reads the un-jittered entry image: done in a second pass) */
    if (eep->jitter && entry_rgba && exit_rgba) {
        float stepLen = eep->jitter_step_length;
#pragma omp parallel for schedule(static)
        for (int py = 0; py < H; ++py)
            for (int px = 0; px < W; ++px) {
                size_t pix = (size_t)py * W + px;
                float* fp = entry_rgba + 4 * pix; const float* bp = exit_rgba + 4 * pix;
                float fragDepth;
                if (bp[3] == 0.0f || fp[3] == 0.0f) fragDepth = 1.0f;
                else {
                    v3 fr = V3(fp[0], fp[1], fp[2]), bk = V3(bp[0], bp[1], bp[2]);
                    float rayLength = length3(sub3(bk, fr));
                    fragDepth = entry_depth ? entry_depth[pix] : 1.0f;
                    if (!(rayLength <= stepLen)) {
                        v3 dir = V3((bk.x - fr.x) / rayLength, (bk.y - fr.y) / rayLength, (bk.z - fr.z) / rayLength);
                        float x = (float)px + 0.5f, y = (float)py + 0.5f;
                        /* permute(i) = mod(62 i^2 + i, 961), GLSL mod(a,b) = a - b*floor(a/b) (eep_jitter.frag:38-40,69) */
                        float permY = y; permY = (62.0f * permY * permY + permY); permY = permY - 961.0f * floorf(permY / 961.0f);
                        float arg = x + permY;
                        float perm = (62.0f * arg * arg + arg); perm = perm - 961.0f * floorf(perm / 961.0f);
                        float jitterValue = perm / 961.0f;
                        v3 nf = add3(fr, scale3(dir, jitterValue * stepLen));
                        float a = 1.0f - jitterValue * stepLen;
                        if (!eep->world_coords) { fp[0] = orc_unorm16(nf.x); fp[1] = orc_unorm16(nf.y); fp[2] = orc_unorm16(nf.z); fp[3] = orc_unorm16(a); }
                        else { fp[0] = half_round(nf.x); fp[1] = half_round(nf.y); fp[2] = half_round(nf.z); fp[3] = half_round(a); }
                    }
                }
                if (entry_depth) entry_depth[pix] = orc_depth24(fragDepth);
            }
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* sort-last merge reference and the reference's image comparator                                                */
/* ------------------------------------------------------------------------------------------------------------ */
void orc_composite(int mode, const float* fc, const float* fd, const float* bc, const float* bd, float* dc, float* dd, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        const float* F = fc + 4 * i; const float* B = bc + 4 * i; float* D = dc + 4 * i;
        float fdp = fd ? fd[i] : 1.f, bdp = bd ? bd[i] : 1.f;
        int ert = (mode & B200RC_COMPOSITE_ERT) != 0;
        int m = mode & 0xff;
        if (m == B200RC_COMP_DVR) {
            float k = (ert && F[3] >= 0.95f) ? 0.0f : 1.0f - F[3];
            float r0 = F[0] + k * B[0], r1 = F[1] + k * B[1], r2 = F[2] + k * B[2], r3 = F[3] + k * B[3];
            D[0] = r0; D[1] = r1; D[2] = r2; D[3] = r3;
            if (dd) dd[i] = fdp < bdp ? fdp : bdp;
        } else if (m == B200RC_COMP_MIP || m == B200RC_COMP_MOP) {
            /* the key (max intensity / max alpha) travels in the depth slot of the partial images */
            int takeB = (bdp > fdp);
            const float* S = takeB ? B : F;
            float r0 = S[0], r1 = S[1], r2 = S[2], r3 = S[3];
            D[0] = r0; D[1] = r1; D[2] = r2; D[3] = r3;
            if (dd) dd[i] = takeB ? bdp : fdp;
        } else {
            /* first hit: the front partial image wins wherever it has a hit */
            int takeB = !(F[3] > 0.f);
            const float* S = takeB ? B : F;
            float r0 = S[0], r1 = S[1], r2 = S[2], r3 = S[3];
            D[0] = r0; D[1] = r1; D[2] = r2; D[3] = r3;
            if (dd) dd[i] = takeB ? bdp : fdp;
        }
    }
}

int64_t orc_image_diff_count(const float* a, const float* b, int W, int H, float tol, int radius) {
    int64_t bad = 0;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const float* o = b + 4 * ((size_t)y * W + x);
            const float* r = a + 4 * ((size_t)y * W + x);
            float dist = sqrtf((r[0] - o[0]) * (r[0] - o[0]) + (r[1] - o[1]) * (r[1] - o[1]) + (r[2] - o[2]) * (r[2] - o[2]) + (r[3] - o[3]) * (r[3] - o[3])) / 2.f;
            if (dist > tol) {
                int match = 0;
                for (int ny = y - radius; ny <= y + radius && !match; ++ny)
                    for (int nx = x - radius; nx <= x + radius && !match; ++nx) {
                        if (nx < 0 || nx >= W || ny < 0 || ny >= H) continue;
                        const float* q = a + 4 * ((size_t)ny * W + nx);
                        float nd = sqrtf((q[0] - o[0]) * (q[0] - o[0]) + (q[1] - o[1]) * (q[1] - o[1]) + (q[2] - o[2]) * (q[2] - o[2]) + (q[3] - o[3]) * (q[3] - o[3])) / 2.f;
                        if (nd <= tol) match = 1;
                    }
                if (!match) bad++;
            }
        }
    return bad;
}

/* ------------------------------------------------------------------------------------------------------------ */
/* synthetic "CT-like" benchmark volume -- no reference counterpart; this is the CPU twin  */
/* of b200rc_synth_volume (voreen_b200/csrc/volume.cu synth_kernel), same integer noise and the same un-contracted */
/* fp32 sequence, so that both produce identical bytes (tests/test_gpu_parity.py::test_synth_volume_matches).     */
/* ------------------------------------------------------------------------------------------------------------ */
static inline uint32_t synth_hash32(uint32_t x) {
    x ^= x >> 16; x *= 0x7feb352du; x ^= x >> 15; x *= 0x846ca68bu; x ^= x >> 16;
    return x;
}

void orc_synth_volume(const int32_t dims[3], const int32_t lo[3], const int32_t rdims[3], uint64_t seed, uint16_t* out) {
    float cx[64], cy[64], cz[64], r2[64];
    const uint32_t sd = (uint32_t)seed;
    for (int i = 0; i < 64; ++i) {
        const uint32_t h0 = synth_hash32(sd * 0x9e3779b9u + (uint32_t)(i * 4 + 0)), h1 = synth_hash32(sd * 0x9e3779b9u + (uint32_t)(i * 4 + 1));
        const uint32_t h2 = synth_hash32(sd * 0x9e3779b9u + (uint32_t)(i * 4 + 2)), h3 = synth_hash32(sd * 0x9e3779b9u + (uint32_t)(i * 4 + 3));
        cx[i] = ((float)(h0 & 0xffffu) * (1.0f / 65536.0f) - 0.5f) * 0.76f;
        cy[i] = ((float)(h1 & 0xffffu) * (1.0f / 65536.0f) - 0.5f) * 0.76f;
        cz[i] = ((float)(h2 & 0xffffu) * (1.0f / 65536.0f) - 0.5f) * 0.76f;
        const float r = 0.012f + (float)(h3 & 0xffffu) * (1.0f / 65536.0f) * 0.02f;
        r2[i] = r * r;
    }
    const int mind = dims[0] < dims[1] ? (dims[0] < dims[2] ? dims[0] : dims[2]) : (dims[1] < dims[2] ? dims[1] : dims[2]);
    const float fall_scale = 0.36f * (float)mind;
    const float rcpx = 1.0f / (float)dims[0], rcpy = 1.0f / (float)dims[1], rcpz = 1.0f / (float)dims[2];
#pragma omp parallel for schedule(dynamic, 1)
    for (int zz_ = 0; zz_ < rdims[2]; ++zz_) {
        const int z = zz_ + lo[2];
        const float pz = ((float)z + 0.5f) * rcpz - 0.5f;
        const float zz = pz * pz;
        /* spheres that can touch this slice / row: az*az <= r2 is implied by d2 <= r2 (rounded sums of non-negative
         * terms are monotone), so the candidate lists are exactly conservative */
        int cz_list[64], ncz = 0;
        for (int i = 0; i < 64; ++i) { const float az = pz - cz[i]; if (az * az <= r2[i]) cz_list[ncz++] = i; }
        for (int yy_ = 0; yy_ < rdims[1]; ++yy_) {
            const int y = yy_ + lo[1];
            const float py = ((float)y + 0.5f) * rcpy - 0.5f;
            const float yy = py * py;
            int cy_list[64], ncy = 0;
            for (int k = 0; k < ncz; ++k) { const int i = cz_list[k]; const float ay = py - cy[i]; if (ay * ay <= r2[i]) cy_list[ncy++] = i; }
            uint16_t* row = out + ((size_t)zz_ * rdims[1] + yy_) * rdims[0];
            for (int xx_ = 0; xx_ < rdims[0]; ++xx_) {
                const int x = xx_ + lo[0];
                const float px = ((float)x + 0.5f) * rcpx - 0.5f;
                const float xx = px * px;
                const float rr = (xx + yy) + zz;
                const float q2 = ((xx * 5.6689343f) + (yy * 7.7160494f)) + (zz * 6.25f);
                const float q = sqrtf(q2);
                float fall = ((1.0f - q) * fall_scale) * 0.25f;
                fall = fminf(fmaxf(fall, 0.0f), 1.0f);
                const float s = (fall * fall) * (3.0f - (2.0f * fall));
                float f = 0.35f * s;
                int hit = 0;
                for (int k = 0; k < ncy && !hit; ++k) {
                    const int i = cy_list[k];
                    const float ax = px - cx[i], ay = py - cy[i], az = pz - cz[i];
                    const float d2 = ((ax * ax) + (ay * ay)) + (az * az);
                    hit = d2 <= r2[i];
                }
                if (hit) f = fmaxf(f, 0.6f);
                if (rr >= 0.09f && rr <= 0.1024f) f = fmaxf(f, 0.85f);
                int val = (int)floorf((f * 65535.0f) + 0.5f);
                const uint32_t h = synth_hash32((uint32_t)x * 73856093u ^ (uint32_t)y * 19349663u ^ (uint32_t)z * 83492791u ^ sd * 0x9e3779b9u);
                const uint32_t hb = synth_hash32(h ^ 0x68bc21ebu);
                const int sum = (int)(h & 0xffffu) + (int)(h >> 16) + (int)(hb & 0xffffu) + (int)(hb >> 16);
                val += ((sum - 131070) * 1135) >> 16;
                val = val < 0 ? 0 : (val > 65535 ? 65535 : val);
                row[xx_] = (uint16_t)val;
            }
        }
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* VolumeRaycaster::rescaleRendering (volumeraycaster.cpp:348-377) + copyimage.frag:37-55                         */
/* ------------------------------------------------------------------------------------------------------------ */
void orc_rescale_rgba16(const uint16_t* src, const float* src_depth, int sw, int sh, uint16_t* dst, float* dst_depth, int dw, int dh) {
    #pragma omp parallel for schedule(static)
    for (int y = 0; y < dh; ++y)
        for (int x = 0; x < dw; ++x) {
            /* fragCoord = gl_FragCoord.xy * screenDimRCP_ (copyimage.frag:43-47); GL_LINEAR: u = s * size - 0.5 */
            float fx = ((float)x + 0.5f) * (1.0f / (float)dw), fy = ((float)y + 0.5f) * (1.0f / (float)dh);
            float u = fx * (float)sw - 0.5f, v = fy * (float)sh - 0.5f;
            float fu = floorf(u), fv = floorf(v);
            float a = u - fu, b = v - fv;
            int x0 = (int)fu, x1 = x0 + 1, y0 = (int)fv, y1 = y0 + 1;
            x0 = x0 < 0 ? 0 : (x0 > sw - 1 ? sw - 1 : x0); x1 = x1 < 0 ? 0 : (x1 > sw - 1 ? sw - 1 : x1);
            y0 = y0 < 0 ? 0 : (y0 > sh - 1 ? sh - 1 : y0); y1 = y1 < 0 ? 0 : (y1 > sh - 1 ? sh - 1 : y1);
            size_t o = (size_t)y * dw + x;
            for (int c = 0; c < 4; ++c) {
                float c00 = (float)src[((size_t)y0 * sw + x0) * 4 + c] / 65535.f, c10 = (float)src[((size_t)y0 * sw + x1) * 4 + c] / 65535.f;
                float c01 = (float)src[((size_t)y1 * sw + x0) * 4 + c] / 65535.f, c11 = (float)src[((size_t)y1 * sw + x1) * 4 + c] / 65535.f;
                float r = lerp1(lerp1(c00, c10, a), lerp1(c01, c11, a), b);
                r = r > 0.f ? (r > 1.f ? 1.f : r) : 0.f;
                dst[o * 4 + c] = (uint16_t)floorf(r * 65535.f + 0.5f);          /* GL_RGBA16 store */
            }
            if (src_depth && dst_depth) {
                float d00 = src_depth[(size_t)y0 * sw + x0], d10 = src_depth[(size_t)y0 * sw + x1], d01 = src_depth[(size_t)y1 * sw + x0], d11 = src_depth[(size_t)y1 * sw + x1];
                dst_depth[o] = orc_depth24(lerp1(lerp1(d00, d10, a), lerp1(d01, d11, a), b));
            }
        }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* image-space post chain: Compositor (compositor.frag:41-140) and Background (background.frag:35-57 over the       */
/* "monochrome" / "gradient" generators of background.cpp:262-300).  Per-pixel; RGBA16 unorm or RGBA32F images.     */
/* ------------------------------------------------------------------------------------------------------------ */
static inline void post_fetch(const void* img, int format, size_t i, float c[4]) {
    if (format == B200RC_OUT_RGBA16) { const uint16_t* q = (const uint16_t*)img + 4 * i; for (int k = 0; k < 4; ++k) c[k] = (float)q[k] / 65535.f; }
    else memcpy(c, (const float*)img + 4 * i, 4 * sizeof(float));
}
static inline uint16_t post_unorm16(float v) { if (!(v > 0.f)) return 0; if (v > 1.f) v = 1.f; return (uint16_t)floorf(v * 65535.f + 0.5f); }
static inline void post_store(void* img, int format, size_t i, const float c[4]) {
    if (format == B200RC_OUT_RGBA16) { uint16_t* q = (uint16_t*)img + 4 * i; for (int k = 0; k < 4; ++k) q[k] = post_unorm16(c[k]); }
    else memcpy((float*)img + 4 * i, c, 4 * sizeof(float));
}

void orc_image_compositor(int mode, float wf, float w0, float w1, int add_depth, const void* c0p, const float* d0p, const void* c1p, const float* d1p,
                          int format, void* dc, float* dd, size_t n) {
    for (size_t i = 0; i < n; ++i) {
        float a[4], b[4], f[4] = {0.f, 0.f, 0.f, 0.f}, fd = 1.0f;
        post_fetch(c0p, format, i, a); post_fetch(c1p, format, i, b);
        const float da = d0p ? d0p[i] : 1.0f, db = d1p ? d1p[i] : 1.0f;
        switch (mode) {
        case B200RC_COMPOSITOR_DEPTH_TEST:
            if (da < db) { memcpy(f, a, sizeof f); fd = da; } else { memcpy(f, b, sizeof f); fd = db; }
            break;
        case B200RC_COMPOSITOR_ALPHA_COMPOSITING:
            if (db <= da) { for (int k = 0; k < 3; ++k) f[k] = b[k] + a[k] * a[3] * (1.0f - b[3]); f[3] = b[3] + a[3] * (1.0f - b[3]); }
            else { for (int k = 0; k < 3; ++k) f[k] = a[k] + b[k] * b[3] * (1.0f - a[3]); f[3] = a[3] + b[3] * (1.0f - a[3]); }
            fd = da < db ? da : db;
            break;
        case B200RC_COMPOSITOR_ALPHA_BLENDING:
            for (int k = 0; k < 3; ++k) f[k] = a[3] * a[k] + (1.0f - a[3]) * b[k];
            f[3] = a[3] + b[3] * (1.0f - a[3]);
            fd = a[3] > 0.f ? da : db;
            break;
        case B200RC_COMPOSITOR_ALPHA_BLENDING_B_OVER_A:
            for (int k = 0; k < 3; ++k) f[k] = b[3] * b[k] + (1.0f - b[3]) * a[k];
            f[3] = b[3] + a[3] * (1.0f - b[3]);
            fd = b[3] > 0.f ? db : da;
            break;
        case B200RC_COMPOSITOR_WEIGHTED_AVERAGE:
            fd = (wf == 0.0f) ? db : (wf == 1.0f ? da : (da < db ? da : db));
            for (int k = 0; k < 4; ++k) f[k] = wf * a[k] + (1.0f - wf) * b[k];
            break;
        case B200RC_COMPOSITOR_TAKE_FIRST: case B200RC_COMPOSITOR_TAKE_SECOND_IF_READY: memcpy(f, a, sizeof f); fd = da; break;
        case B200RC_COMPOSITOR_TAKE_SECOND: memcpy(f, b, sizeof f); fd = db; break;
        case B200RC_COMPOSITOR_FIRST_HAS_PRIORITY: if (a[3] > 0.f) { memcpy(f, a, sizeof f); fd = da; } else { memcpy(f, b, sizeof f); fd = db; } break;
        case B200RC_COMPOSITOR_SECOND_HAS_PRIORITY: if (b[3] > 0.f) { memcpy(f, b, sizeof f); fd = db; } else { memcpy(f, a, sizeof f); fd = da; } break;
        case B200RC_COMPOSITOR_MAXIMUM_ALPHA:
            fd = da < db ? da : db;
            for (int k = 0; k < 3; ++k) f[k] = a[k] * a[3] + b[k] * b[3];
            f[3] = a[3] > b[3] ? a[3] : b[3];
            break;
        case B200RC_COMPOSITOR_DIFFERENCE:
            fd = fabsf(da - db);
            for (int k = 0; k < 3; ++k) f[k] = 1.0f - fabsf(a[k] - b[k]);
            f[3] = a[3] > b[3] ? a[3] : b[3];
            break;
        case B200RC_COMPOSITOR_DIFFERENCE_ALPHA_MUL:
            fd = da < db ? da : db;
            for (int k = 0; k < 3; ++k) f[k] = fabsf(a[k] * a[3] - b[k] * b[3]);
            f[3] = 1.0f;
            break;
        case B200RC_COMPOSITOR_ADD:
            for (int k = 0; k < 4; ++k) f[k] = w0 * a[k] + w1 * b[k];
            fd = add_depth ? w0 * da + w1 * db : da;
            break;
        default: break;
        }
        post_store(dc, format, i, f);
        if (dd) dd[i] = fd;
    }
}

void orc_image_background(const b200rc_background_params* bp, const void* src, const float* sdepth, int format, void* dc, float* dd, int W, int H) {
    const int ang = bp->angle_deg;
    const float rad = (float)ang * 0.017453292519943295f;
    const float s = 1.0f + (float)(45 - abs(ang % 90 - 45)) / 45.0f * 0.8284271247461900976033774484194f;
    const float sin_a = sinf(rad), cos_a = cosf(rad), scale_rcp = 1.0f / s;
    for (int y = 0; y < H; ++y)
        for (int x = 0; x < W; ++x) {
            const size_t i = (size_t)y * W + x;
            float bg[4];
            memcpy(bg, bp->color1, sizeof bg);
            if (bp->mode == B200RC_BACKGROUND_GRADIENT) {
                const float X = ((float)x + 0.5f) * (2.0f / (float)W) - 1.0f, Y = ((float)y + 0.5f) * (2.0f / (float)H) - 1.0f;
                const float ly = (-sin_a * X + cos_a * Y) * scale_rcp;
                float w = 0.5f * ly + 0.5f;
                w = w < 0.f ? 0.f : (w > 1.f ? 1.f : w);
                for (int k = 0; k < 4; ++k) bg[k] = fmaf(w, bp->color2[k] - bp->color1[k], bp->color1[k]);
            }
            if (format == B200RC_OUT_RGBA16) for (int k = 0; k < 4; ++k) bg[k] = (float)post_unorm16(bg[k]) / 65535.f;
            float f[4], fd = 1.0f;
            memcpy(f, bg, sizeof f);
            if (src) {
                float c[4];
                post_fetch(src, format, i, c);
                const float d = sdepth ? sdepth[i] : 1.0f, om = 1.0f - c[3];
                if (bp->blend_mode == B200RC_BACKGROUND_BLEND_ALPHA) {
                    for (int k = 0; k < 3; ++k) f[k] = c[k] * c[3] + bg[k] * bg[3] * om;
                    f[3] = c[3] + bg[3] * om;
                } else
                    for (int k = 0; k < 4; ++k) { f[k] = c[k] + bg[k] * om; if (f[k] > 1.0f) f[k] = 1.0f; }
                fd = d < 1.0f ? d : 1.0f;
            }
            post_store(dc, format, i, f);
            if (dd) dd[i] = fd;
        }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* TransFunc1DGaussian texture (transfunc1dgaussian.cpp:122-178, transfuncmappingcurve.cpp:257-288,                 */
/* transfunc1d.cpp:85-116)                                                                                          */
/* ------------------------------------------------------------------------------------------------------------ */
static float gcurve_opacity(const orc_tf_curve* c, float x) {
    /* exp() is the double overload there: the product and the sum are evaluated in double, the result is returned as float */
    double e = exp((double)(-(x - c->intensity) * (x - c->intensity) / (2 * c->width)));
    float a = (float)(x < c->intensity ? c->color_l[3] : c->color_r[3]) / 255.f;
    return (float)(e * (double)(a - c->base_value) + (double)c->base_value);
}
void orc_tf1d_gaussian_texture(const orc_tf_curve* curves, int n, const float thr[2], float gamma, int alpha_mode, int width, uint8_t* out) {
    int front_end = iround_tgt(thr[0] * (float)width), back_start = iround_tgt(thr[1] * (float)width);
    for (int x = 0; x < width; ++x) {
        uint8_t* px = out + 4 * x;
        px[0] = px[1] = px[2] = px[3] = 0;
        if (!(x < front_end || x > back_start) && n > 0) {
            float v = (float)x / (float)(width - 1);
            if (gamma != 1.f) v = powf(v, gamma);
            float col[4] = {0.f, 0.f, 0.f, 0.f}, alphaSum = 0.f;
            for (int i = 0; i < n; ++i) {
                const orc_tf_curve* c = curves + i;
                if (!c->active) continue;
                float yRatio = (float)exp((double)(-(v - c->intensity) * (v - c->intensity) / (2 * c->width)));
                const uint8_t* peak = v < c->intensity ? c->color_l : c->color_r;
                const uint8_t* base = v < c->intensity ? c->base_color_l : c->base_color_r;
                uint8_t cc[4];
                for (int k = 0; k < 3; ++k) cc[k] = (uint8_t)((float)peak[k] * yRatio + (float)base[k] * (1.f - yRatio));
                cc[3] = (uint8_t)(gcurve_opacity(c, v) * 255);
                float alpha = gcurve_opacity(c, v);
                if (alpha < FLT_MIN) alpha = FLT_MIN;
                alphaSum += alpha;
                for (int k = 0; k < 3; ++k) col[k] += ((float)cc[k] * alpha);
                col[3] = ((float)cc[3] > col[3] ? (float)cc[3] : col[3]);
            }
            if (alphaSum > 1.f) alphaSum = 1.f;
            for (int k = 0; k < 3; ++k) col[k] = col[k] / alphaSum;
            for (int k = 0; k < 4; ++k) { float t = col[k]; t = t < 0.f ? 0.f : (t > 255.f ? 255.f : t); px[k] = (uint8_t)t; }
        }
        if (alpha_mode == 2) { if (px[3]) px[3] = 255; }
        else if (alpha_mode == 0) px[3] = 0;
    }
}

/* ------------------------------------------------------------------------------------------------------------ */
/* SingleOctreeRaycasterCL, DVR (singleoctreeraycaster.cl:190-300; octree_config.cl:150-155,313-368,402-413,         */
/* 437-506,590-646; octree_dvr.cl:52-80), one channel, over the dense per-level node arrays of oracle/octree_ref.py   */
/* (all levels concatenated at (8^k - 1) / 7; flags bit 2 = has brick, bit 3 = has children)                          */
/* ------------------------------------------------------------------------------------------------------------ */
static inline size_t oct_off(int k) { return (((size_t)1 << (3 * k)) - 1) / 7; }
static v4 oct_tf_linear(const uint8_t* tf, int w, float s) {      /* read_imagef, CLK_FILTER_LINEAR, CLK_ADDRESS_CLAMP_TO_EDGE, UNORM_INT8 */
    float u = s * (float)w - 0.5f, fu = floorf(u), a = u - fu;
    int x0 = (int)fu, x1 = x0 + 1;
    if (x0 < 0) x0 = 0; if (x0 > w - 1) x0 = w - 1; if (x1 < 0) x1 = 0; if (x1 > w - 1) x1 = w - 1;
    v4 c0 = V4(tf[4 * x0] / 255.f, tf[4 * x0 + 1] / 255.f, tf[4 * x0 + 2] / 255.f, tf[4 * x0 + 3] / 255.f);
    v4 c1 = V4(tf[4 * x1] / 255.f, tf[4 * x1 + 1] / 255.f, tf[4 * x1 + 2] / 255.f, tf[4 * x1 + 3] / 255.f);
    return lerp4(c0, c1, a);
}
void orc_octree_raycast(const uint16_t* avg, const uint8_t* flags, const uint16_t* bricks, int levels, int B, const int32_t dims[3], int octree_dim,
                        const uint8_t* tf, int tf_width, float dom_lo, float dom_hi, float rwm_scale, float rwm_offset,
                        float step, int lod, int filter_linear, int compositing, const float* entry, const float* exitp, int W, int H, float* out, float* out_firsthit) {
    (void)levels;
    const float sc[3] = {(float)dims[0] / (float)octree_dim, (float)dims[1] / (float)octree_dim, (float)dims[2] / (float)octree_dim};
#pragma omp parallel for schedule(dynamic, 2)
    for (int py = 0; py < H; ++py)
        for (int px = 0; px < W; ++px) {
            size_t pix = (size_t)py * W + px;
            v3 en = V3(entry[4 * pix] * sc[0], entry[4 * pix + 1] * sc[1], entry[4 * pix + 2] * sc[2]);
            v3 ex = V3(exitp[4 * pix] * sc[0], exitp[4 * pix + 1] * sc[1], exitp[4 * pix + 2] * sc[2]);
            v4 color = V4(0, 0, 0, 0); float firsthit = 1.0f;
            v3 dir = sub3(ex, en);
            float tEnd = sqrtf((dir.x * dir.x + dir.y * dir.y) + dir.z * dir.z);
            int moving = tEnd > 0.f;
            if (moving) dir = V3(dir.x / tEnd, dir.y / tEnd, dir.z / tEnd);
            if (entry[4 * pix + 3] > 0.f || exitp[4 * pix + 3] > 0.f) {
                float param = 0.f, tEndNode = -1.f;
                v3 nllf = V3(0, 0, 0), nrcp = V3(1, 1, 1); size_t node = 0; int has_brick = 0;
                /* RayInfo (singleoctreeraycaster.cl:360-368): current.intensity -1, pending.intensity 0, first hits 1 */
                float cur_i = -1.0f, pend_i = 0.0f, cur_fh = 1.0f, pend_fh = 1.0f;
                v4 cur_c = V4(0, 0, 0, 0), pend_c = V4(0, 0, 0, 0);
                if (compositing == B200RC_COMP_MOP) {       /* :232-237 */
                    cur_c = oct_tf_linear(tf, tf_width, ((cur_i * rwm_scale + rwm_offset) - dom_lo) / (dom_hi - dom_lo));
                    pend_c = oct_tf_linear(tf, tf_width, ((pend_i * rwm_scale + rwm_offset) - dom_lo) / (dom_hi - dom_lo));
                }
                while (param <= tEnd) {
                    v3 s = V3(en.x + param * dir.x, en.y + param * dir.y, en.z + param * dir.z);
                    if (param > (tEndNode + 1.e-6f)) {
                        int k = 0, ix = 0, iy = 0, iz = 0; v3 llf = V3(0, 0, 0); float half = 1.0f;
                        node = 0;
                        while (k < lod && (flags[node] & 8)) {          /* getNodeAtSamplePos */
                            half = half * 0.5f;
                            int cx = s.x >= llf.x + half, cy = s.y >= llf.y + half, cz = s.z >= llf.z + half;
                            ix = ix * 2 + cx; iy = iy * 2 + cy; iz = iz * 2 + cz;
                            llf = V3(llf.x + half * (float)cx, llf.y + half * (float)cy, llf.z + half * (float)cz);
                            ++k;
                            int G = 1 << k;
                            node = oct_off(k) + ((size_t)iz * G + iy) * G + ix;
                        }
                        nllf = llf;
                        v3 urb = V3(llf.x + half, llf.y + half, llf.z + half);
                        has_brick = (flags[node] & 4) != 0;
                        nrcp = V3(1.f / (urb.x - llf.x), 1.f / (urb.y - llf.y), 1.f / (urb.z - llf.z));
                        if (moving) {       /* fetchNextRayNode over computeNodeExitPoint */
                            float ex_ = dir.x >= 0.f ? urb.x : llf.x, ey_ = dir.y >= 0.f ? urb.y : llf.y, ez_ = dir.z >= 0.f ? urb.z : llf.z;
                            float tx = dir.x != 0.f ? (ex_ - s.x) / dir.x : 1e6f, ty = dir.y != 0.f ? (ey_ - s.y) / dir.y : 1e6f, tz = dir.z != 0.f ? (ez_ - s.z) / dir.z : 1e6f;
                            float tmin = gl_min(tx, gl_min(ty, tz)) - 1e-6f;
                            v3 nexit = V3(s.x + tmin * dir.x, s.y + tmin * dir.y, s.z + tmin * dir.z);
                            float toff = gl_min((nexit.x - s.x) / dir.x, gl_min((nexit.y - s.y) / dir.y, (nexit.z - s.z) / dir.z));
                            toff = floorf(toff / step) * step;
                            tEndNode = param + toff;
                        } else tEndNode = param;
                    }
                    float intensity;
                    if (s.x < 0.f || s.x > 1.f || s.y < 0.f || s.y > 1.f || s.z < 0.f || s.z > 1.f) intensity = 0.f;
                    else if (has_brick) {
                        const uint16_t* brick = bricks + node * (size_t)B * B * B;
                        v3 q = V3((s.x - nllf.x) * nrcp.x, (s.y - nllf.y) * nrcp.y, (s.z - nllf.z) * nrcp.z);
                        if (filter_linear) {
                            float cxf = q.x * (float)(B - 1), cyf = q.y * (float)(B - 1), czf = q.z * (float)(B - 1);
                            int fx = (int)floorf(cxf), fy = (int)floorf(cyf), fz = (int)floorf(czf);
                            int lx = fx < 0 ? 0 : (fx > B - 1 ? B - 1 : fx), ly = fy < 0 ? 0 : (fy > B - 1 ? B - 1 : fy), lz = fz < 0 ? 0 : (fz > B - 1 ? B - 1 : fz);
                            int hx = fx + 1 > B - 1 ? B - 1 : fx + 1, hy = fy + 1 > B - 1 ? B - 1 : fy + 1, hz = fz + 1 > B - 1 ? B - 1 : fz + 1;
                            float pxw = cxf - (float)lx, pyw = cyf - (float)ly, pzw = czf - (float)lz;
#define OV(x, y, z) ((float)brick[((size_t)(z) * B + (y)) * B + (x)])
                            float acc = OV(lx, ly, lz) * (1.f - pxw) * (1.f - pyw) * (1.f - pzw);
                            acc = acc + OV(lx, ly, hz) * (1.f - pxw) * (1.f - pyw) * pzw;
                            acc = acc + OV(lx, hy, lz) * (1.f - pxw) * pyw * (1.f - pzw);
                            acc = acc + OV(lx, hy, hz) * (1.f - pxw) * pyw * pzw;
                            acc = acc + OV(hx, ly, lz) * pxw * (1.f - pyw) * (1.f - pzw);
                            acc = acc + OV(hx, ly, hz) * pxw * (1.f - pyw) * pzw;
                            acc = acc + OV(hx, hy, lz) * pxw * pyw * (1.f - pzw);
                            acc = acc + OV(hx, hy, hz) * pxw * pyw * pzw;
#undef OV
                            intensity = acc * (1.f / 65536.f);
                        } else {
                            int cx = (int)rintf(q.x * (float)B - 0.5f), cy = (int)rintf(q.y * (float)B - 0.5f), cz = (int)rintf(q.z * (float)B - 0.5f);
                            cx = cx < 0 ? 0 : (cx > B - 1 ? B - 1 : cx); cy = cy < 0 ? 0 : (cy > B - 1 ? B - 1 : cy); cz = cz < 0 ? 0 : (cz > B - 1 ? B - 1 : cz);
                            intensity = (float)brick[((size_t)cz * B + cy) * B + cx] * (1.f / 65536.f);
                        }
                    } else intensity = clampf((float)avg[node] * 1.f / 65535.f, 0.f, 1.f);
                    if (compositing == B200RC_COMP_MIP) {           /* octree_mip.cl applyTFandCombineColors */
                        if (pend_i < intensity) {
                            pend_i = intensity;
                            pend_fh = gl_min(pend_fh, param / tEnd);
                            if (cur_i < pend_i) { cur_i = pend_i; cur_fh = pend_fh; }
                        }
                        param += step;
                        continue;
                    }
                    float rw = intensity * rwm_scale + rwm_offset;
                    v4 c = oct_tf_linear(tf, tf_width, (rw - dom_lo) / (dom_hi - dom_lo));
                    if (compositing == B200RC_COMP_MOP) {           /* octree_mop.cl applyTFandCombineColors */
                        if (pend_c.w < c.w) {
                            pend_c = c; pend_i = intensity;
                            pend_fh = gl_min(pend_fh, param / tEnd);
                            if (cur_c.w < pend_c.w) { cur_c = pend_c; cur_i = pend_i; cur_fh = pend_fh; }
                        }
                        param += step;
                        continue;
                    }
                    if (c.w > 0.f) {
                        v3 rgb = V3(gl_min(c.x, 1.f), gl_min(c.y, 1.f), gl_min(c.z, 1.f));
                        float a = 1.f - powf(1.f - c.w, step * 200.f);
                        float k = (1.f - color.w) * a;
                        color.x += k * rgb.x; color.y += k * rgb.y; color.z += k * rgb.z; color.w += k;
                        firsthit = gl_min(param / tEnd, firsthit);
                    }
                    if (color.w >= 0.95f) { color.w = 1.f; param = tEnd + step; }
                    param += step;
                }
                if (compositing != B200RC_COMP_DVR) {       /* postRaycastingLoop, rayFinished (full-frame mode) */
                    cur_i = pend_i; cur_fh = pend_fh;
                    if (compositing == B200RC_COMP_MOP) cur_c = pend_c;
                    else cur_c = oct_tf_linear(tf, tf_width, ((cur_i * rwm_scale + rwm_offset) - dom_lo) / (dom_hi - dom_lo));
                    color = V4(0, 0, 0, 0);
                    if (cur_c.w > 0.f) color = cur_c;
                    color = V4(gl_min(color.x, 1.f), gl_min(color.y, 1.f), gl_min(color.z, 1.f), gl_min(color.w, 1.f));
                    firsthit = cur_fh;
                }
            }
            out[4 * pix] = color.x; out[4 * pix + 1] = color.y; out[4 * pix + 2] = color.z; out[4 * pix + 3] = color.w;
            if (out_firsthit) out_firsthit[pix] = firsthit;
        }
}


/* ------------------------------------------------------------------------------------------------------------ */
/* TransFunc2DPrimitives                                                                                         */
/* ------------------------------------------------------------------------------------------------------------ */
typedef struct { float p[2]; float c[4]; } tfp_vertex;
typedef struct { tfp_vertex* v; int n, cap; } tfp_list;

static void tfp_push(tfp_list* l, const float p[2], const float c[4], int zero_alpha) {
    if (l->n == l->cap) { l->cap = l->cap ? l->cap * 2 : 256; l->v = (tfp_vertex*)realloc(l->v, sizeof(tfp_vertex) * (size_t)l->cap); }
    tfp_vertex* v = &l->v[l->n++];
    v->p[0] = p[0]; v->p[1] = p[1];
    v->c[0] = c[0]; v->c[1] = c[1]; v->c[2] = c[2]; v->c[3] = zero_alpha ? 0.f : c[3];
}
/* primitives of an IMode.begin(TRIANGLE_STRIP / TRIANGLE_FAN / TRIANGLES) ... end() block -> independent triangles */
static void tfp_emit(tfp_list* tris, const tfp_list* blk, int mode /* 0 strip, 1 fan, 2 triangles */) {
    for (int i = 0; i + 2 < blk->n; i += (mode == 2 ? 3 : 1)) {
        const tfp_vertex* a = mode == 1 ? &blk->v[0] : &blk->v[i];
        const tfp_vertex *b = &blk->v[i + 1], *c = &blk->v[i + 2];
        tfp_push(tris, a->p, a->c, 0); tfp_push(tris, b->p, b->c, 0); tfp_push(tris, c->p, c->c, 0);
    }
}

/* TransFuncTriangle::paint (:98-189) / TransFuncQuad::paint (:240-352) */
static void tfp_polygon(tfp_list* tris, const orc_tf2d_primitive* P, int n) {
    float col[4][4], center[2] = {0.f, 0.f}, ccol[4] = {0.f, 0.f, 0.f, 0.f};
    for (int i = 0; i < n; ++i) {
        for (int k = 0; k < 4; ++k) col[i][k] = (float)P->color[i][k] / 255.f;
        center[0] += P->position[i][0]; center[1] += P->position[i][1];
        for (int k = 0; k < 4; ++k) ccol[k] += col[i][k];
    }
    center[0] /= (float)n; center[1] /= (float)n;
    for (int k = 0; k < 4; ++k) ccol[k] /= (float)n;
    float mid[4][2], midc[4][4], in[4][2];
    for (int i = 0; i < n; ++i) {
        const int j = (i + 1) % n;
        for (int k = 0; k < 2; ++k) mid[i][k] = 0.5f * P->position[i][k] + 0.5f * P->position[j][k];
        for (int k = 0; k < 3; ++k) midc[i][k] = 0.5f * col[i][k] + 0.5f * col[j][k];
        midc[i][3] = 0.f;
        for (int k = 0; k < 2; ++k) in[i][k] = P->fuzziness * P->position[i][k] + (1.f - P->fuzziness) * center[k];
    }
    /* the strips, as sequences of (kind, index): O outer corner (alpha 0), I inner corner, M edge middle */
    static const char* tri_blocks[] = {"O0I0M2I2O2M1", "I2I1M1O1", "O1I1M0I0O0", 0};
    static const char* quad_blocks[] = {"O0I0M3I3O3M2", "I3I2M2O2", "O2I2M1I1O1M0", "I1I0M0O0", 0};
    const char** blocks = n == 3 ? tri_blocks : quad_blocks;
    for (int b = 0; blocks[b]; ++b) {
        tfp_list blk = {0, 0, 0};
        for (const char* s = blocks[b]; *s; s += 2) {
            const int i = s[1] - '0';
            if (s[0] == 'O') tfp_push(&blk, P->position[i], col[i], 1);
            else if (s[0] == 'I') tfp_push(&blk, in[i], col[i], 0);
            else tfp_push(&blk, mid[i], midc[i], 0);
        }
        tfp_emit(tris, &blk, 0);
        free(blk.v);
    }
    tfp_list blk = {0, 0, 0};
    if (n == 3) {
        for (int i = 0; i < 3; ++i) tfp_push(&blk, in[i], col[i], 0);
        tfp_emit(tris, &blk, 2);
    } else {
        tfp_push(&blk, center, ccol, 0);
        for (int i = 0; i < 4; ++i) tfp_push(&blk, in[i], col[i], 0);
        tfp_push(&blk, in[0], col[0], 0);
        tfp_emit(tris, &blk, 1);
    }
    free(blk.v);
}

/* TransFuncBanana::paintInner (:405-490) */
static void tfp_banana(tfp_list* tris, const orc_tf2d_primitive* P) {
    const int steps = 20;
    float c[4][4];
    for (int i = 0; i < 4; ++i) for (int k = 0; k < 4; ++k) c[i][k] = (float)P->color[i][k] / 255.f;
    float t1[2], t2[2], tc[2], t3[2], t4[2], t1c[4], t2c[4], tcc[4], t3c[4], t4c[4];
    for (int k = 0; k < 2; ++k) {
        t1[k] = (2.f * P->position[1][k]) - (0.5f * P->position[0][k]) - (0.5f * P->position[3][k]);
        t2[k] = (2.f * P->position[2][k]) - (0.5f * P->position[0][k]) - (0.5f * P->position[3][k]);
        tc[k] = (t1[k] + t2[k]) / 2.f;
        t3[k] = P->fuzziness * t1[k] + (1.f - P->fuzziness) * tc[k];
        t4[k] = P->fuzziness * t2[k] + (1.f - P->fuzziness) * tc[k];
    }
    for (int k = 0; k < 4; ++k) {
        t1c[k] = (2.f * c[1][k]) - (0.5f * c[0][k]) - (0.5f * c[3][k]);
        t2c[k] = (2.f * c[2][k]) - (0.5f * c[0][k]) - (0.5f * c[3][k]);
        tcc[k] = (t1c[k] + t2c[k]) / 2.f;
        t3c[k] = P->fuzziness * t1c[k] + (1.f - P->fuzziness) * tcc[k];
        t4c[k] = P->fuzziness * t2c[k] + (1.f - P->fuzziness) * tcc[k];
    }
    const float* upper[3] = {t1, t3, t4}; const float* lower[3] = {t3, t4, t2};
    const float* upper_c[3] = {t1c, t3c, t4c}; const float* lower_c[3] = {t3c, t4c, t2c};
    tfp_list blk = {0, 0, 0};
    for (int band = 0; band < 3; ++band) {
        tfp_push(&blk, P->position[0], c[0], 0);
        for (int i = 0; i < steps; ++i) {
            const float t = (float)i / (float)(steps - 1);
            float v1[2], v2[2], v1c[4], v2c[4];
            for (int k = 0; k < 2; ++k) {
                v1[k] = (((1 - t) * (1 - t)) * P->position[0][k]) + ((2 * (1 - t) * t) * upper[band][k]) + ((t * t) * P->position[3][k]);
                v2[k] = (((1 - t) * (1 - t)) * P->position[0][k]) + ((2 * (1 - t) * t) * lower[band][k]) + ((t * t) * P->position[3][k]);
            }
            for (int k = 0; k < 4; ++k) {
                v1c[k] = (((1 - t) * (1 - t)) * c[0][k]) + ((2 * (1 - t) * t) * upper_c[band][k]) + ((t * t) * c[3][k]);
                v2c[k] = (((1 - t) * (1 - t)) * c[0][k]) + ((2 * (1 - t) * t) * lower_c[band][k]) + ((t * t) * c[3][k]);
            }
            tfp_push(&blk, v1, v1c, band == 0);
            tfp_push(&blk, v2, v2c, band == 2);
        }
        tfp_push(&blk, P->position[3], c[3], 0);
    }
    tfp_emit(tris, &blk, 0);
    free(blk.v);
}

void orc_tf2d_primitives_texture(const orc_tf2d_primitive* prims, int n, int width, int height, const float thr_x[2], const float thr_y[2],
                                 const float gamma[2], int alpha_mode, float* out) {
    tfp_list tris = {0, 0, 0};
    for (int i = 0; i < n; ++i) {
        if (prims[i].type == 0) tfp_polygon(&tris, &prims[i], 3);
        else if (prims[i].type == 1) tfp_polygon(&tris, &prims[i], 4);
        else tfp_banana(&tris, &prims[i]);
    }
    const size_t npix = (size_t)width * height;
    float* help = (float*)calloc(npix * 4, sizeof(float));      /* glClear: (0, 0, 0, 0) */
    /* every texel centre against every triangle, in paint order (GL_MAX is order independent) */
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            const double px = ((double)x + 0.5) / (double)width, py = ((double)y + 0.5) / (double)height;
            float* d = help + ((size_t)y * width + x) * 4;
            for (int k = 0; k + 2 < tris.n; k += 3) {
                const tfp_vertex *A = &tris.v[k], *B = &tris.v[k + 1], *Cv = &tris.v[k + 2];
                const double area = ((double)B->p[0] - A->p[0]) * ((double)Cv->p[1] - A->p[1]) - ((double)B->p[1] - A->p[1]) * ((double)Cv->p[0] - A->p[0]);
                if (area == 0.0) continue;
                const double w0 = (((double)B->p[0] - px) * ((double)Cv->p[1] - py) - ((double)B->p[1] - py) * ((double)Cv->p[0] - px)) / area;
                const double w1 = (((double)Cv->p[0] - px) * ((double)A->p[1] - py) - ((double)Cv->p[1] - py) * ((double)A->p[0] - px)) / area;
                const double w2 = (((double)A->p[0] - px) * ((double)B->p[1] - py) - ((double)A->p[1] - py) * ((double)B->p[0] - px)) / area;
                if (w0 < 0.0 || w1 < 0.0 || w2 < 0.0) continue;
                for (int ch = 0; ch < 4; ++ch) {
                    const float f = (float)(w0 * A->c[ch] + w1 * B->c[ch] + w2 * Cv->c[ch]);
                    if (f > d[ch]) d[ch] = f;
                }
            }
        }
    free(tris.v);
    const int xf = iround_tgt(thr_x[0] * (float)width), xb = iround_tgt(thr_x[1] * (float)width);
    const int yf = iround_tgt(thr_y[0] * (float)height), yb = iround_tgt(thr_y[1] * (float)height);
    for (int y = 0; y < height; ++y)
        for (int x = 0; x < width; ++x) {
            float* o = out + ((size_t)y * width + x) * 4;
            o[0] = o[1] = o[2] = o[3] = 0.f;
            if (y < yf || y > yb || x < xf || x > xb) continue;
            float vx = (float)x / (float)(width - 1), vy = (float)y / (float)(height - 1);        /* applyGammaToValues (:140-149) */
            if (gamma[0] != 1.f) vx = powf(vx, gamma[0]);
            if (gamma[1] != 1.f) vy = powf(vy, gamma[1]);
            int tx = (int)(vx * (float)width), ty = (int)(vy * (float)height);                      /* getMappingForValueFloat (:150-160) */
            tx = tx < 0 ? 0 : (tx > width - 1 ? width - 1 : tx); ty = ty < 0 ? 0 : (ty > height - 1 ? height - 1 : ty);
            const float* h = help + ((size_t)ty * width + tx) * 4;
            o[0] = h[0]; o[1] = h[1]; o[2] = h[2]; o[3] = h[3];
            if (alpha_mode == 2) { if (o[3]) o[3] = 1.f; }
            else if (alpha_mode == 0) o[3] = 0.f;
        }
    free(help);
}
