// qb2_pass_inst.cu -- one kernel variant per translation unit (the Makefile compiles this file once per
// value of QB2_INST, in parallel: a single unit with all fourteen variants takes six minutes).
#include "qb2_internal.cuh"

#if QB2_INST < 12
#include "qb2_pass_kernel.cuh"
#else
#include "qb2_pass_tma.cuh"
#endif

namespace qb2 {

#define QB2_VARIANT(k, ...) \
    int launch_pass_variant_##k(const PassLaunch &L) { return launch_kernel<__VA_ARGS__>(L); }

#if QB2_INST == 0
QB2_VARIANT(0, float, 4, 256, QB2_MINB4, false)
#elif QB2_INST == 1
QB2_VARIANT(1, float, 4, 256, QB2_MINB4, true)
#elif QB2_INST == 2
QB2_VARIANT(2, float, 4, 512, 2, false)
#elif QB2_INST == 3
QB2_VARIANT(3, float, 4, 512, 2, true)
#elif QB2_INST == 4
QB2_VARIANT(4, float, 5, 256, 2, false)
#elif QB2_INST == 5
QB2_VARIANT(5, float, 5, 256, 2, true)
#elif QB2_INST == 6
QB2_VARIANT(6, float, 3, 256, 6, false)
#elif QB2_INST == 7
QB2_VARIANT(7, float, 3, 256, 6, true)
#elif QB2_INST == 8
QB2_VARIANT(8, double, 3, 256, 4, false)
#elif QB2_INST == 9
QB2_VARIANT(9, double, 3, 256, 4, true)
#elif QB2_INST == 10
QB2_VARIANT(10, double, 4, 256, 2, false)
#elif QB2_INST == 11
QB2_VARIANT(11, double, 4, 256, 2, true)
#elif QB2_INST == 12
int launch_pass_variant_12(const PassLaunch &L) { return launch_kernel_tma<false>(L); }
#elif QB2_INST == 13
int launch_pass_variant_13(const PassLaunch &L) { return launch_kernel_tma<true>(L); }
#else
#error "QB2_INST must be 0..13"
#endif

#if defined(QB2_TRACE) && QB2_INST == 4
extern "C" int qb2_debug_read_trace(unsigned long long *out, int count) {
    if (count > QB2_TRACE_TILES * QB2_TRACE_STAMPS) count = QB2_TRACE_TILES * QB2_TRACE_STAMPS;
    cudaDeviceSynchronize();
    const cudaError_t e = cudaMemcpyFromSymbol(out, qb2_trace_buf, sizeof(unsigned long long) * count);
    static unsigned long long zero[QB2_TRACE_TILES * QB2_TRACE_STAMPS];
    cudaMemcpyToSymbol(qb2_trace_buf, zero, sizeof(zero));
    return e == cudaSuccess ? count : -1;
}
#endif

}  // namespace qb2
