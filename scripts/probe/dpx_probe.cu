// Dev probe: overflow behaviour of VIADDMNMX.S16x2.RELU and friends on sm_100a.
#include <cstdio>
#include <cuda_runtime.h>
__global__ void k(const unsigned *a, const unsigned *b, unsigned *o, int n)
{
    int i = threadIdx.x;
    if (i < n) {
        o[i] = __viaddmin_s16x2_relu(a[i], b[i], 0x00ff00ffu);
        o[n + i] = __vadd2(a[i], b[i]);
        o[2 * n + i] = __viaddmin_s32_relu((int)a[i], (int)b[i], 255);
    }
}
int main()
{
    const int n = 6;
    unsigned ha[n] = {0x00ff00ffu, 0x00ff0000u, 0x000000ffu, 0x00800080u, 0x00010001u, 0x00ff00ffu};
    unsigned hb[n] = {0x7fff7fffu, 0x80008000u, 0x7f017f00u, 0xff80ff7fu, 0xffffffffu, 0x00010001u};
    unsigned *a, *b, *o, ho[3 * n];
    cudaMalloc(&a, sizeof ha); cudaMalloc(&b, sizeof hb); cudaMalloc(&o, sizeof ho);
    cudaMemcpy(a, ha, sizeof ha, cudaMemcpyHostToDevice); cudaMemcpy(b, hb, sizeof hb, cudaMemcpyHostToDevice);
    k<<<1, 32>>>(a, b, o, n);
    cudaMemcpy(ho, o, sizeof ho, cudaMemcpyDeviceToHost);
    for (int i = 0; i < n; i++) printf("a=%08x b=%08x viaddmin_s16x2_relu=%08x vadd2=%08x s32=%08x\n", ha[i], hb[i], ho[i], ho[n + i], ho[2 * n + i]);
    printf("%s\n", cudaGetErrorString(cudaDeviceSynchronize()));
    return 0;
}
