// TEST INFRASTRUCTURE — stand-in for rmm/device_vector.hpp (see cuda_async_memory_resource.hpp).
#pragma once
#include <thrust/device_vector.h>
#include <rmm/mr/device/cuda_async_memory_resource.hpp>
namespace rmm {
template <typename T>
using device_vector = thrust::device_vector<T, rmm::mr::thrust_allocator<T>>;
}
