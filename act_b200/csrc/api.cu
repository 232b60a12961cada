// Library-level entry points of libact_b200.so.
#include "common.cuh"

extern "C" int act_version(void) { return 100; }

extern "C" const char *act_error_string(int code) {
    switch (code) {
        case ACT_OK: return "ok";
        case ACT_EINVAL: return "act_b200: invalid argument (null pointer or bad shape)";
        case ACT_EUNSUPPORTED: return "act_b200: shape not supported by the sm_100a kernels";
        case ACT_EALIGN: return "act_b200: pointer alignment requirement not met";
        default: break;
    }
    if (code > 0) return cudaGetErrorString((cudaError_t)code);
    return "act_b200: unknown error";
}

namespace act {
int &attn_tc_mode();
int &gemm_sm_cap();
}

extern "C" int act_set_option(int key, int value) {
    if (key == ACT_OPT_PDL) {
        act::pdl_flag() = value ? 1 : 0;
        return ACT_OK;
    }
    if (key == ACT_OPT_ATTN_TC) {
        if (value < 0 || value > 2) return ACT_EINVAL;
        act::attn_tc_mode() = value;
        return ACT_OK;
    }
    if (key == ACT_OPT_GEMM_SM_CAP) {
        if (value < 0) return ACT_EINVAL;
        act::gemm_sm_cap() = value;
        return ACT_OK;
    }
    return ACT_EINVAL;
}
