// TEST INFRASTRUCTURE — stand-in for rmm/exec_policy.hpp (see cuda_async_memory_resource.hpp).
#pragma once
#include <thrust/execution_policy.h>
#include <thrust/system/cuda/execution_policy.h>
#include <rmm/mr/device/cuda_async_memory_resource.hpp>
namespace rmm {
// asynchronous thrust policy whose temporary storage comes from the given pool
inline auto exec_policy_nosync(cudaStream_t stream, rmm::mr::cuda_async_memory_resource* mr)
{
    return thrust::cuda::par_nosync(rmm::mr::thrust_allocator<char>(stream, mr)).on(stream);
}
inline auto exec_policy(cudaStream_t stream, rmm::mr::cuda_async_memory_resource* mr)
{
    return thrust::cuda::par(rmm::mr::thrust_allocator<char>(stream, mr)).on(stream);
}
}  // namespace rmm
