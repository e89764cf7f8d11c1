// Shared device helpers for the qmprs_b200 kernels (complex128 arithmetic on double2).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

typedef double2 cplx;

// kernel classes for the launch counter / event profiler (api.cu)
enum { QM_CLS_GEMM = 0, QM_CLS_SVD_GRAM, QM_CLS_SVD_EIG, QM_CLS_SVD_APPLY, QM_CLS_SVD_LAYOUT, QM_CLS_QR_VEC,
       QM_CLS_QR_APPLY, QM_CLS_SMALL, QM_CLS_GATE, QM_CLS_ENV, QM_CLS_SVD_ROUND, QM_NCLS };
void qm_prof_pre(int cls, cudaStream_t st);
void qm_prof_post(int cls, cudaStream_t st);
// algorithmic work of the launches of a class (flops for compute kernels, bytes for streaming ones)
void qm_prof_work(int cls, double work);
// true while the event profiler brackets every launch (multi-stream schedules fall back to one stream)
bool qm_prof_active();
// every kernel launch of the library goes through this macro
#define QM_LAUNCH(cls, st, ...)          \
    do {                                 \
        qm_prof_pre((cls), (st));        \
        __VA_ARGS__;                     \
        qm_prof_post((cls), (st));       \
    } while (0)

#define QM_CHECK_LAUNCH()                                  \
    do {                                                   \
        cudaError_t e__ = cudaGetLastError();              \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)

#define QM_CUDA(call)                                      \
    do {                                                   \
        cudaError_t e__ = (call);                          \
        if (e__ != cudaSuccess) return (int)e__;           \
    } while (0)

__host__ __device__ __forceinline__ cplx mk(double re, double im) { cplx z; z.x = re; z.y = im; return z; }
__host__ __device__ __forceinline__ cplx cconj(cplx a) { return mk(a.x, -a.y); }
__host__ __device__ __forceinline__ cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
__host__ __device__ __forceinline__ cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
__host__ __device__ __forceinline__ cplx cmul(cplx a, cplx b) { return mk(a.x * b.x - a.y * b.y, a.x * b.y + a.y * b.x); }
// a * conj(b)
__host__ __device__ __forceinline__ cplx cmulc(cplx a, cplx b) { return mk(a.x * b.x + a.y * b.y, a.y * b.x - a.x * b.y); }
// conj(a) * b
__host__ __device__ __forceinline__ cplx ccmul(cplx a, cplx b) { return mk(a.x * b.x + a.y * b.y, a.x * b.y - a.y * b.x); }
__host__ __device__ __forceinline__ cplx cscale(cplx a, double s) { return mk(a.x * s, a.y * s); }
__host__ __device__ __forceinline__ double cabs2(cplx a) { return a.x * a.x + a.y * a.y; }
// acc += a*b
__host__ __device__ __forceinline__ void cfma(cplx& acc, cplx a, cplx b) {
    acc.x += a.x * b.x - a.y * b.y;
    acc.y += a.x * b.y + a.y * b.x;
}
// acc += a*conj(b)
__host__ __device__ __forceinline__ void cfmac(cplx& acc, cplx a, cplx b) {
    acc.x += a.x * b.x + a.y * b.y;
    acc.y += a.y * b.x - a.x * b.y;
}
// acc += conj(a)*b
__host__ __device__ __forceinline__ void ccfma(cplx& acc, cplx a, cplx b) {
    acc.x += a.x * b.x + a.y * b.y;
    acc.y += a.x * b.y - a.y * b.x;
}

__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}

// Block-wide sum; result valid in every thread.  `red` must hold >= 33 doubles.
__device__ __forceinline__ double block_sum(double v, double* red) {
    int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
    int nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[w] = v;
    __syncthreads();
    if (w == 0) {
        double t = (lane < nw) ? red[lane] : 0.0;
        t = warp_sum(t);
        if (lane == 0) red[32] = t;
    }
    __syncthreads();
    return red[32];
}


// Programmatic dependent launch (sm_90+): a kernel launched through qm_launch_dep may start (CTA
// scheduling, parameter setup) while its predecessor in the stream is still draining; pdl_wait()
// blocks until the predecessor has completed and its writes are visible, so it must precede every
// global-memory access of the kernel.  pdl_trigger() lets the NEXT kernel begin its own launch.
// The dependent chains of this library (gram -> eig -> update per Jacobi round, one environment
// kernel per gate-step) are thousands of short kernels: the launch latency is what this hides.
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
bool qm_pdl_enabled();
template <typename... KArgs, typename... Args>
static inline cudaError_t qm_launch_dep(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                        Args&&... args) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 0;
    if (qm_pdl_enabled()) {                      // plain launch while the stream is being captured into a graph
        cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
        if (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone) cfg.numAttrs = 1;
    }
    return cudaLaunchKernelEx(&cfg, kernel, KArgs(args)...);
}

static inline int ceil_div(long long a, long long b) { return (int)((a + b - 1) / b); }
