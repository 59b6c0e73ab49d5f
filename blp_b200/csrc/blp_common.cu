// blp_common.cu -- host-side plumbing shared by the C-ABI entry points:
// thread-local error text, launch accounting, device check.
#include <stdarg.h>
#include <string.h>

#include "blp_common.cuh"

namespace blp {

static thread_local char tl_error[512] = "";
static thread_local int tl_launches = 0;

void set_error(const char *fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(tl_error, sizeof(tl_error), fmt, ap);
    va_end(ap);
}

static thread_local int tl_prof_which = 0;
static thread_local cudaEvent_t tl_prof_start = nullptr, tl_prof_stop = nullptr;

void prof_begin(int which, cudaStream_t st) {
    if (which == tl_prof_which && tl_prof_start) cudaEventRecord(tl_prof_start, st);
}
void prof_end(int which, cudaStream_t st) {
    if (which == tl_prof_which && tl_prof_stop) cudaEventRecord(tl_prof_stop, st);
}

void count_launch(int n) { tl_launches += n; }
void reset_launch_count() { tl_launches = 0; }

int check_cuda(cudaError_t e, const char *what) {
    if (e == cudaSuccess) return BLP_OK;
    set_error("CUDA error in %s: %s", what, cudaGetErrorString(e));
    return BLP_ECUDA;
}

}  // namespace blp

extern "C" int blp_version(void) { return BLP_B200_VERSION; }

extern "C" const char *blp_last_error(void) { return blp::tl_error; }

extern "C" int blp_last_launch_count(void) { return blp::tl_launches; }

extern "C" int blp_profile_events(int which, void *start_event, void *stop_event) {
    if (which < 0 || which > 2) { blp::set_error("which must be 0 (off), 1 (eval sweep kernel) or 2 (train kernel)"); return BLP_EINVAL; }
    blp::tl_prof_which = which;
    blp::tl_prof_start = which ? (cudaEvent_t)start_event : nullptr;
    blp::tl_prof_stop = which ? (cudaEvent_t)stop_event : nullptr;
    return BLP_OK;
}

extern "C" int blp_device_check(int device) {
    cudaDeviceProp p;
    cudaError_t e = cudaGetDeviceProperties(&p, device);
    if (e != cudaSuccess) return blp::check_cuda(e, "cudaGetDeviceProperties");
    if (p.major != 10) {
        blp::set_error("device %d is sm_%d%d; libblp_b200 is built for sm_100a only", device, p.major, p.minor);
        return BLP_EARCH;
    }
    return BLP_OK;
}
