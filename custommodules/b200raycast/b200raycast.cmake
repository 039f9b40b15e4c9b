# Voreen module "b200raycast": drop-in SingleVolumeRaycaster + MultiVolumeRaycaster backed by libb200rc.so (hand-written sm_100a CUDA).
# Consumed by V/cmake/commonconf.cmake:456-497 exactly like V/modules/sample/sample.cmake; enable with
#   cmake -DVRN_CUSTOM_MODULEDIR=<repo>/custommodules -DVRN_MODULE_B200RAYCAST=ON -DB200RC_ROOT=<repo> ...
# The module links only the C ABI (include/b200rc.h); no CUDA headers or nvcc are needed to build Voreen.

IF(NOT B200RC_ROOT)
    GET_FILENAME_COMPONENT(B200RC_ROOT "${MOD_DIR}/../.." ABSOLUTE)
ENDIF()

SET(MOD_CORE_MODULECLASS B200RaycastModule)

SET(MOD_CORE_SOURCES
    ${MOD_DIR}/processors/b200singlevolumeraycaster.cpp
    ${MOD_DIR}/processors/b200multivolumeraycaster.cpp
    ${MOD_DIR}/processors/b200simpleraycaster.cpp
    ${MOD_DIR}/processors/b200rgbraycaster.cpp
)
SET(MOD_CORE_HEADERS
    ${MOD_DIR}/processors/b200singlevolumeraycaster.h
    ${MOD_DIR}/processors/b200multivolumeraycaster.h
    ${MOD_DIR}/processors/b200simpleraycaster.h
    ${MOD_DIR}/processors/b200rgbraycaster.h
)

SET(MOD_INCLUDE_DIRECTORIES ${B200RC_ROOT}/include)
SET(MOD_LIBRARIES ${B200RC_ROOT}/voreen_b200/libb200rc.so)
# lets basemodule.cpp drop its own registration of the type name (patches/basemodule_guard.patch)
SET(MOD_DEFINITIONS -DVRN_MODULE_B200RAYCAST)
