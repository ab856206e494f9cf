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

// L2 residency hint for the optimizer moments.  At ~10 resident clients the Adam moments M | V (2 x K x ld fp32, 34 MB for
// SimpleCNN) are touched once per step, by the optimizer kernel only, and would fit in the 126 MB L2 -- but the ~55 MB of
// activations that stream through between two optimizer launches evict them, so every step re-reads and re-writes them
// through HBM (57 % of the optimizer's 118 MB).  An access-policy window on the launching stream marks [base, base + bytes) as
// persisting (captured into the kernel nodes of the epoch graph); bytes <= 0, or more bytes than the device's L2 set-aside
// holds, clears the window.  Returns the number of bytes covered (0: no window).  Safe to call repeatedly.
extern "C" long long flb_l2_persist_window(const void* base, long long bytes, void* stream) {
    int dev = 0, max_persist = 0, max_window = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return -1;
    cudaDeviceGetAttribute(&max_persist, cudaDevAttrMaxPersistingL2CacheSize, dev);
    cudaDeviceGetAttribute(&max_window, cudaDevAttrMaxAccessPolicyWindowSize, dev);
    // all or nothing: a window that covers a fraction of a larger state would only take L2 away from the convolutions
    const bool fits = base && bytes > 0 && max_persist > 0 && bytes <= max_persist && bytes <= max_window;
    // The set-aside is DEVICE state and outlives the engine that asked for it: it is sized for the caller of the moment, and
    // given back (persisting lines demoted) as soon as a caller without a window runs -- a 34 MB set-aside left behind by a
    // 10-client engine cost a following 100-client CIFAR10CNN engine 4 % (measured).  Only changed when the value changes, so
    // the calls made under stream capture (same engine, same value) never touch the device limit.
    static long long cur_limit[16] = {0};
    const long long want = fits ? bytes : 0;
    if (dev >= 0 && dev < 16 && cur_limit[dev] != want) {
        if (want == 0) (void)cudaCtxResetPersistingL2Cache();
        if (cudaDeviceSetLimit(cudaLimitPersistingL2CacheSize, (size_t)want) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
        cur_limit[dev] = want;
    }
    cudaStreamAttrValue v;
    memset(&v, 0, sizeof(v));
    if (fits) {
        v.accessPolicyWindow.base_ptr = const_cast<void*>(base);
        v.accessPolicyWindow.num_bytes = (size_t)bytes;
        v.accessPolicyWindow.hitRatio = 1.0f;
        v.accessPolicyWindow.hitProp = cudaAccessPropertyPersisting;
        v.accessPolicyWindow.missProp = cudaAccessPropertyStreaming;
    }
    if (cudaStreamSetAttribute((cudaStream_t)stream, cudaStreamAttributeAccessPolicyWindow, &v) != cudaSuccess) { (void)cudaGetLastError(); return 0; }
    return want;
}
