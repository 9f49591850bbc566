// TEST INFRASTRUCTURE — minimal stand-in for RMM v24.12.00 (not vendored by the reference and
// absent offline), just enough API for /root/reference/spgemm.cu:33-35,809-817 to compile
// unmodified.  Written from the call sites, not from RMM's sources.
#pragma once
#include <cuda_runtime.h>
#include <thrust/device_malloc_allocator.h>
#include <thrust/device_ptr.h>
#include <cstddef>
#include <cstdint>

namespace rmm {
namespace mr {

// stream-ordered pool resource over cudaMemPool_t (what rmm::mr::cuda_async_memory_resource wraps)
class cuda_async_memory_resource {
public:
    explicit cuda_async_memory_resource(std::size_t initial_pool_size = 0)
    {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaMemPoolProps props{};
        props.allocType = cudaMemAllocationTypePinned;
        props.handleTypes = cudaMemHandleTypeNone;
        props.location.type = cudaMemLocationTypeDevice;
        props.location.id = dev;
        cudaMemPoolCreate(&pool_, &props);
        std::uint64_t threshold = UINT64_MAX;  // rmm keeps memory in the pool (release threshold = max)
        cudaMemPoolSetAttribute(pool_, cudaMemPoolAttrReleaseThreshold, &threshold);
        if (initial_pool_size) {  // rmm primes the pool by allocating and freeing the initial size
            void* p = nullptr;
            if (cudaMallocFromPoolAsync(&p, initial_pool_size, pool_, nullptr) == cudaSuccess) cudaFreeAsync(p, nullptr);
            cudaStreamSynchronize(nullptr);
        }
    }
    ~cuda_async_memory_resource() { if (pool_) cudaMemPoolDestroy(pool_); }
    cuda_async_memory_resource(const cuda_async_memory_resource&) = delete;
    cuda_async_memory_resource(cuda_async_memory_resource&& o) noexcept : pool_(o.pool_) { o.pool_ = nullptr; }

    void* allocate_async(std::size_t bytes, cudaStream_t s)
    {
        void* p = nullptr;
        if (bytes == 0) return nullptr;
        cudaMallocFromPoolAsync(&p, bytes, pool_, s);
        return p;
    }
    void deallocate_async(void* p, std::size_t, cudaStream_t s) { if (p) cudaFreeAsync(p, s); }

private:
    cudaMemPool_t pool_ = nullptr;
};

template <typename T>
class thrust_allocator : public thrust::device_malloc_allocator<T> {
public:
    using Base = thrust::device_malloc_allocator<T>;
    using pointer = typename Base::pointer;
    using size_type = typename Base::size_type;
    template <typename U>
    struct rebind { using other = thrust_allocator<U>; };

    thrust_allocator() = default;
    thrust_allocator(cudaStream_t s, cuda_async_memory_resource& mr) : stream_(s), mr_(&mr) {}
    thrust_allocator(cudaStream_t s, cuda_async_memory_resource* mr) : stream_(s), mr_(mr) {}
    template <typename U>
    thrust_allocator(thrust_allocator<U> const& o) : stream_(o.stream()), mr_(o.resource()) {}

    pointer allocate(size_type n)
    {
        return thrust::device_pointer_cast(static_cast<T*>(mr_->allocate_async(n * sizeof(T), stream_)));
    }
    void deallocate(pointer p, size_type n) { mr_->deallocate_async(thrust::raw_pointer_cast(p), n * sizeof(T), stream_); }
    cudaStream_t stream() const { return stream_; }
    cuda_async_memory_resource* resource() const { return mr_; }

private:
    cudaStream_t stream_ = nullptr;
    cuda_async_memory_resource* mr_ = nullptr;
};

}  // namespace mr
}  // namespace rmm
