// ABI plumbing shared by every entry point of libim2im_uq.so.
#include "common.cuh"

namespace im2im {

char* last_error_buffer() {
    static thread_local char buf[512] = "";
    return buf;
}

std::atomic<unsigned long long>& launch_counter() {
    static std::atomic<unsigned long long> n{0};
    return n;
}

int sm_count() {
    static int cached[64] = {0};
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= 64) return 148;
    if (cached[dev] == 0) {
        int n = 0;
        if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
        cached[dev] = n;
    }
    return cached[dev];
}

}  // namespace im2im

extern "C" {

int im2im_abi_version(void) { return IM2IM_ABI_VERSION; }

const char* im2im_last_error(void) { return im2im::last_error_buffer(); }

unsigned long long im2im_launch_count(void) { return im2im::launch_counter().load(std::memory_order_relaxed); }

}  // extern "C"
