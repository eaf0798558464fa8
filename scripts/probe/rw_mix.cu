// Development probe: achievable HBM bandwidth for different read:write mixes (grid-stride, 16-byte accesses).
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o rw_mix rw_mix.cu && ./rw_mix
#include <cstdio>
#include <cuda_runtime.h>

template <int R, int W>
__global__ void mix(const uint4 *__restrict__ in, uint4 *__restrict__ out, size_t n)
{
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) {
        uint4 v = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int r = 0; r < R; r++) { const uint4 t = __ldg(in + i + (size_t)r * n); v.x ^= t.x; v.y += t.y; v.z ^= t.z; v.w += t.w; }
#pragma unroll
        for (int w = 0; w < W; w++) { v.x += w; out[i + (size_t)w * n] = v; }
        if (W == 0 && v.x == 0x12345678u && v.y == 77u) out[0] = v;
    }
}

template <int R, int W>
static void run(const char *name, const uint4 *in, uint4 *out, size_t n)
{
    cudaEvent_t a, b; cudaEventCreate(&a); cudaEventCreate(&b);
    const int grid = 148 * 8;
    for (int k = 0; k < 3; k++) mix<R, W><<<grid, 256>>>(in, out, n);
    cudaEventRecord(a);
    const int reps = 10;
    for (int k = 0; k < reps; k++) mix<R, W><<<grid, 256>>>(in, out, n);
    cudaEventRecord(b); cudaEventSynchronize(b);
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double bytes = (double)reps * n * 16 * (R + W);
    printf("%-22s %7.1f GB/s (%d read : %d write)\n", name, bytes / ms / 1e6, R, W);
}

int main()
{
    const size_t n = (size_t)1 << 26;                 // 1 GiB per stream
    uint4 *in, *out;
    cudaMalloc(&in, n * 16 * 2); cudaMalloc(&out, n * 16 * 2);
    cudaMemset(in, 1, n * 16 * 2); cudaMemset(out, 0, n * 16 * 2);
    run<1, 0>("read only", in, out, n);
    run<0, 1>("write only", in, out, n);
    run<1, 1>("copy", in, out, n);
    run<1, 2>("k3 mix (1:2)", in, out, n);
    run<2, 1>("2:1", in, out, n);
    run<2, 2>("copy, 2 streams each", in, out, n);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
