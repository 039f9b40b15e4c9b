// voreen_probe.cpp -- pins the oracle's transfer-function, pre-integration and voxel arithmetic against the reference's
// OWN C++ sources, compiled where they lie (oracle/Makefile `ref`; nothing of the reference is copied into this repository):
//   V/src/core/datastructures/transfunc/1d/1dkeys/transfunc1dkeys.cpp        TransFunc1DKeys::getMappingForValueUByte (:119-154)
//   V/src/core/datastructures/transfunc/1d/1dkeys/utils/transfuncmappingkey.cpp
//   V/src/core/datastructures/transfunc/1d/transfunc1d.cpp                   TransFunc1D::createTextureData (:85-116), gamma
//   V/src/core/datastructures/transfunc/transfuncbase.cpp                    TransFuncBase::applyAlpha (:177-210)
//   V/src/core/datastructures/transfunc/1d/preintegrationtable.cpp           PreIntegrationTable::computeTable (:67-204), classify (:255-277)
//   V/src/core/datastructures/transfunc/1d/preintegrationtablemap.cpp
//   V/src/core/datastructures/volume/volumeram.cpp                           VolumeRAM::getVoxelNormalizedLinear (:78-97)
//   V/include/voreen/core/datastructures/volume/volumeatomic.h               VolumeAtomic<T>::getVoxelNormalized
// TEST INFRASTRUCTURE.  The GL-facing classes those sources mention (tgt::Texture, RenderTarget, Shader, serialization) are
// never reached on these paths; the link leaves them unresolved (static link, --unresolved-symbols=ignore-all) apart from the
// two trivial RenderTarget constructors defined below.  Built with -fno-access-control: the functions of interest are protected.
//
// Protocol: commands on stdin, one per line; results on stdout as integers or C99 hex floats (exact).
//   tf W N thr_lo thr_hi gamma alpha_mode dom_lo dom_hi     followed by N lines: intensity lr lg lb la rr rg rb ra split
//   texture                     -> 4 W integers: TransFunc1D::createTextureData + TransFuncBase::applyAlpha
//   mapping x                   -> 4 integers:  TransFunc1DKeys::getMappingForValueUByte(x)
//   preint res step integral    -> res*res*4 hex floats: PreIntegrationTable(tf, res, step, integral, false).getTable()
//   classify fs fe              -> 4 hex floats: PreIntegrationTable::classify on the last table
//   gauss W N thr_lo thr_hi gamma alpha_mode                followed by N lines: intensity width base lr lg lb la rr rg rb ra blr blg blb brr brg brb active
//                               -> 4 W integers: the texture of a TransFunc1DGaussian (transfunc1dgaussian.cpp:122-178,
//                                  transfuncmappingcurve.cpp:257-288) through TransFunc1D::createTextureData + applyAlpha
//   vol fmt dx dy dz            followed by one line of dx*dy*dz values (fmt: uint8 uint16 float)
//   voxel x y z                 -> hex float: getVoxelNormalized
//   linear x y z                -> hex float: getVoxelNormalizedLinear (voxel coordinates)
#include "voreen/core/datastructures/transfunc/1d/1dkeys/transfunc1dkeys.h"
#include "voreen/core/datastructures/transfunc/1d/1dkeys/utils/transfuncmappingkey.h"
#include "voreen/core/datastructures/transfunc/1d/preintegrationtable.h"
#include "voreen/core/datastructures/transfunc/1d/1dgaussian/transfunc1dgaussian.h"
#include "voreen/core/datastructures/transfunc/1d/1dgaussian/utils/transfuncmappingcurve.h"
#include "voreen/core/datastructures/rendertarget/rendertarget.h"
#include "voreen/core/datastructures/volume/volumeatomic.h"

#include <cstdio>
#include <iostream>
#include <memory>
#include <sstream>
#include <string>

namespace voreen {      // a PreIntegrationTable owns a RenderTarget member; its GL work happens in initialize(), never called here
RenderTarget::RenderTarget() : fbo_(0), colorTex_(0), depthTex_(0), numUpdates_(0) {}
RenderTarget::~RenderTarget() {}
}

using namespace voreen;

int main() {
    std::unique_ptr<TransFunc1DKeys> tf;
    std::unique_ptr<PreIntegrationTable> table;
    std::unique_ptr<VolumeRAM> vol;
    std::string line;
    while (std::getline(std::cin, line)) {
        std::istringstream is(line);
        std::string cmd;
        if (!(is >> cmd)) continue;
        if (cmd == "tf") {
            int W, N, alphaMode; float tlo, thi, gamma, dlo, dhi;
            is >> W >> N >> tlo >> thi >> gamma >> alphaMode >> dlo >> dhi;
            table.reset();
            tf.reset(new TransFunc1DKeys(W));
            tf->clearKeys();
            for (int i = 0; i < N; ++i) {
                std::getline(std::cin, line);
                std::istringstream ks(line);
                float x; int l[4], r[4], split;
                ks >> x >> l[0] >> l[1] >> l[2] >> l[3] >> r[0] >> r[1] >> r[2] >> r[3] >> split;
                TransFuncMappingKey* k = new TransFuncMappingKey(x, tgt::col4((uint8_t)l[0], (uint8_t)l[1], (uint8_t)l[2], (uint8_t)l[3]));
                if (split) { k->setSplit(true); k->setColorR(tgt::col4((uint8_t)r[0], (uint8_t)r[1], (uint8_t)r[2], (uint8_t)r[3])); }
                tf->addKey(k);
            }
            tf->setThreshold(tlo, thi);
            tf->setGammaValue(gamma);
            tf->setAlphaMode((TransFuncBase::AlphaMode)alphaMode);
            tf->setDomain(dlo, dhi);
            std::printf("ok %d\n", tf->getNumKeys());
        } else if (cmd == "gauss") {
            int W, N, alphaMode; float tlo, thi, gamma;
            is >> W >> N >> tlo >> thi >> gamma >> alphaMode;
            TransFunc1DGaussian g(W);
            g.clearCurves();
            for (int i = 0; i < N; ++i) {
                std::getline(std::cin, line);
                std::istringstream ks(line);
                float x, w, base; int c[14], active;
                ks >> x >> w >> base;
                for (int k = 0; k < 14; ++k) ks >> c[k];
                ks >> active;
                TransFuncMappingCurve* cv = new TransFuncMappingCurve(x, w, base, tgt::col4((uint8_t)c[0], (uint8_t)c[1], (uint8_t)c[2], (uint8_t)c[3]));
                cv->colorL_ = tgt::col4((uint8_t)c[0], (uint8_t)c[1], (uint8_t)c[2], (uint8_t)c[3]);
                cv->colorR_ = tgt::col4((uint8_t)c[4], (uint8_t)c[5], (uint8_t)c[6], (uint8_t)c[7]);
                cv->baseColorL_ = tgt::col4((uint8_t)c[8], (uint8_t)c[9], (uint8_t)c[10], 0);
                cv->baseColorR_ = tgt::col4((uint8_t)c[11], (uint8_t)c[12], (uint8_t)c[13], 0);
                cv->active_ = active != 0;
                g.addCurve(cv);
            }
            g.setThreshold(tlo, thi);
            g.setGammaValue(gamma);
            g.setAlphaMode((TransFuncBase::AlphaMode)alphaMode);
            GLvoid* data = g.createTextureData();
            g.applyAlpha(data);
            const unsigned char* b = reinterpret_cast<const unsigned char*>(data);
            const int n = g.getDimensions().x * 4;
            for (int i = 0; i < n; ++i) std::printf("%d%c", (int)b[i], i == n - 1 ? '\n' : ' ');
            delete[] reinterpret_cast<tgt::Vector4<GLubyte>*>(data);
        } else if (cmd == "texture" && tf) {
            GLvoid* data = tf->createTextureData();
            tf->applyAlpha(data);
            const unsigned char* b = reinterpret_cast<const unsigned char*>(data);
            const int n = tf->getDimensions().x * 4;
            for (int i = 0; i < n; ++i) std::printf("%d%c", (int)b[i], i == n - 1 ? '\n' : ' ');
            delete[] reinterpret_cast<tgt::Vector4<GLubyte>*>(data);
        } else if (cmd == "mapping" && tf) {
            float x; is >> x;
            const tgt::Vector4<GLubyte> c = tf->getMappingForValueUByte(x);
            std::printf("%d %d %d %d\n", (int)c.r, (int)c.g, (int)c.b, (int)c.a);
        } else if (cmd == "preint" && tf) {
            int res, integral; float step;
            is >> res >> step >> integral;
            table.reset(new PreIntegrationTable(tf.get(), (size_t)res, step, integral != 0, false, 0));
            const tgt::vec4* t = table->getTable();
            const size_t n = table->getDimension() * table->getDimension();
            for (size_t i = 0; i < n; ++i) std::printf("%a %a %a %a%c", t[i].x, t[i].y, t[i].z, t[i].w, i == n - 1 ? '\n' : ' ');
        } else if (cmd == "classify" && table) {
            float fs, fe; is >> fs >> fe;
            const tgt::vec4 c = table->classify(fs, fe);
            std::printf("%a %a %a %a\n", c.x, c.y, c.z, c.w);
        } else if (cmd == "vol") {
            std::string fmt; size_t dx, dy, dz;
            is >> fmt >> dx >> dy >> dz;
            std::getline(std::cin, line);
            std::istringstream vs(line);
            const tgt::svec3 dim(dx, dy, dz);
            const size_t n = dx * dy * dz;
            if (fmt == "uint8") { VolumeRAM_UInt8* v = new VolumeRAM_UInt8(dim); for (size_t i = 0; i < n; ++i) { int t; vs >> t; v->voxel(i) = (uint8_t)t; } vol.reset(v); }
            else if (fmt == "uint16") { VolumeRAM_UInt16* v = new VolumeRAM_UInt16(dim); for (size_t i = 0; i < n; ++i) { int t; vs >> t; v->voxel(i) = (uint16_t)t; } vol.reset(v); }
            else { VolumeRAM_Float* v = new VolumeRAM_Float(dim); for (size_t i = 0; i < n; ++i) { float t; vs >> t; v->voxel(i) = t; } vol.reset(v); }
            std::printf("ok %zu\n", n);
        } else if (cmd == "voxel" && vol) {
            size_t x, y, z; is >> x >> y >> z;
            std::printf("%a\n", vol->getVoxelNormalized(x, y, z));
        } else if (cmd == "linear" && vol) {
            float x, y, z; is >> x >> y >> z;
            std::printf("%a\n", vol->getVoxelNormalizedLinear(tgt::vec3(x, y, z)));
        } else {
            std::printf("error\n");
        }
        std::fflush(stdout);
    }
    return 0;
}
