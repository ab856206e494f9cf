// Library-level state: thread-local error string, device check, SM count.
#include "flb_common.cuh"
#include "../../include/flb.h"
#include <string.h>

static thread_local char g_err[512] = "";
static int g_sms = 0;

void flb_set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

int flb_num_sms() {
    if (g_sms > 0) return g_sms;
    int dev = 0, n = 0;
    if (cudaGetDevice(&dev) == cudaSuccess &&
        cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) == cudaSuccess && n > 0) {
        g_sms = n;
        return n;
    }
    return FLB_NUM_SMS_B200;
}

// A second stream per host thread (and device) on which the weight-gradient kernels of a step run beside the
// activation-gradient chain; fork / join go through events, so the pair is capturable into one CUDA graph.
SideLane* flb_side_lane() {
    static thread_local SideLane lanes[16];
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 16) return nullptr;
    SideLane& l = lanes[dev];
    if (!l.ready) {
        if (cudaStreamCreateWithFlags(&l.s, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        if (cudaStreamCreateWithFlags(&l.s2, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
        for (int i = 0; i < 6; ++i)
            if (cudaEventCreateWithFlags(&l.ev[i], cudaEventDisableTiming) != cudaSuccess) return nullptr;
        l.ready = true;
    }
    return &l;
}

extern "C" const char* flb_last_error(void) { return g_err; }

extern "C" int flb_version(void) { return 100; }

// Idempotent.  Fails (no CPU fallback exists) when the device is absent or is not sm_100.
extern "C" int flb_init(int device) {
    int count = 0;
    cudaError_t e = cudaGetDeviceCount(&count);
    if (e != cudaSuccess || count == 0) {
        flb_set_error("flb_init: no CUDA device (%s); this library has no CPU path",
                      e != cudaSuccess ? cudaGetErrorString(e) : "device count 0");
        return FLB_ERR_NODEV;
    }
    FLB_CHECK_ARG(device >= 0 && device < count, "flb_init: device %d out of range (0..%d)", device, count - 1);
    cudaDeviceProp prop;
    FLB_CUDA(cudaGetDeviceProperties(&prop, device));
    if (prop.major != 10) {
        flb_set_error("flb_init: device %d is sm_%d%d; kernels are built for sm_100a only", device, prop.major, prop.minor);
        return FLB_ERR_UNSUPPORTED;
    }
    g_sms = prop.multiProcessorCount;
    return FLB_OK;
}
