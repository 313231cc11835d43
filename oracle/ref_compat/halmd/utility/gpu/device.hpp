// TEST INFRASTRUCTURE ONLY (oracle/Makefile): stand-in for halmd/utility/gpu/device.hpp, whose original pulls in
// lua.hpp (absent here).  The reference's host-side algorithms that the GPU harness compiles in place
// (halmd/algorithm/gpu/radix_sort.hpp:207-214) only ask for the properties of the current device.
#ifndef HALMD_UTILITY_GPU_DEVICE_HPP
#define HALMD_UTILITY_GPU_DEVICE_HPP

#include <cuda_runtime.h>
#include <cuda_wrapper/cuda_wrapper.hpp>

namespace halmd {

class device
{
public:
    static cuda::device const& get()
    {
        static cuda::device* dev = nullptr;
        if (!dev) {
            int num = 0;
            cudaFree(nullptr);          // make sure the runtime's primary context exists
            cudaGetDevice(&num);
            dev = new cuda::device;     // (never destroyed: the primary context belongs to the runtime)
            dev->set(num);
        }
        return *dev;
    }
};

} // namespace halmd

#endif
