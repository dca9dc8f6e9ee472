// Microbenchmark: issue rate of legacy mma.sync m16n8k8 TF32 (and m16n8k16 BF16) on sm_100a,
// used to decide whether the per-head 32x32x32 attention products should go to the tensor cores.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o mma_rate mma_tf32_rate.cu && ./mma_rate
#include <cstdio>
#include <cuda_runtime.h>

template <int KIND>
__global__ void k(float* out, int iters, long long* cyc) {
    float c[8][4];
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) c[i][j] = 0.f;
    unsigned a0 = threadIdx.x, a1 = 2, a2 = 3, a3 = 4, b0 = 5, b1 = 6;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            if (KIND == 0)
                asm volatile("mma.sync.aligned.m16n8k8.row.col.f32.tf32.tf32.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
            else
                asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                             : "+f"(c[i][0]), "+f"(c[i][1]), "+f"(c[i][2]), "+f"(c[i][3])
                             : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
        }
    }
    long long t1 = clock64();
    float s = 0.f;
    for (int i = 0; i < 8; ++i) for (int j = 0; j < 4; ++j) s += c[i][j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

int main() {
    float* out; long long* cyc; long long h;
    cudaMalloc(&out, 148 * 1024 * 4); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    for (int kind = 0; kind < 2; ++kind)
        for (int warps = 1; warps <= 16; warps *= 2) {
            if (kind == 0) k<0><<<148, warps * 32>>>(out, iters, cyc); else k<1><<<148, warps * 32>>>(out, iters, cyc);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            double per_sm = (double)h / (iters * 8.0 * warps);
            printf("%s warps/SM=%2d  cycles=%lld  SM-cycles per MMA=%.2f  (dense MAC/clk/SM = %.0f)\n", kind == 0 ? "tf32 m16n8k8 " : "bf16 m16n8k16",
                   warps, h, per_sm, (kind == 0 ? 1024.0 : 2048.0) / per_sm);
        }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
