// Voreen module class of the B200 ray-casting drop-in (pattern: V/modules/sample/samplemodule.h).
#ifndef VRN_B200RAYCASTMODULE_H
#define VRN_B200RAYCASTMODULE_H

#include "voreen/core/voreenmodule.h"

namespace voreen {

class B200RaycastModule : public VoreenModule {
public:
    B200RaycastModule(const std::string& modulePath);
    virtual std::string getDescription() const;
};

} // namespace

#endif
