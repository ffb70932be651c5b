// scratch probe: achieved store bandwidth of the block kernels' output pattern (3 x 16-byte stores per thread at a
// 48-byte thread stride + 3 x 4-byte stores at a 12-byte stride) against fully coalesced stores of the same bytes
#include <cstdio>
#include <cuda_runtime.h>
#include <stdint.h>
constexpr int W = 3840, H = 2160;
// pattern A: thread = 4 pixels of one row: 3 float4 to state, 3 u32 to out (as finish_quad)
__global__ void k_quad(float* __restrict__ state, uint8_t* __restrict__ out, float v) {
    const int q = blockIdx.x * blockDim.x + threadIdx.x;          // quad index, row-major (W / 4 quads per row)
    if (q >= W / 4 * H) return;
    float4* sp = reinterpret_cast<float4*>(state) + (size_t)q * 3;
    sp[0] = make_float4(v, v, v, v); sp[1] = make_float4(v, v, v, v); sp[2] = make_float4(v, v, v, v);
    uint32_t* op = reinterpret_cast<uint32_t*>(out) + (size_t)q * 3;
    op[0] = 1; op[1] = 2; op[2] = 3;
}
// pattern A2: as the tile kernels: 16 x 16 threads per 64 x 32 tile, thread = 4 x 2 pixels
__global__ void k_tile(float* __restrict__ state, uint8_t* __restrict__ out, float v) {
    const int tiles_x = W / 64, tile = blockIdx.x, tby = tile / tiles_x, tbx = tile - tby * tiles_x;
    const int tx = threadIdx.x & 15, ty = threadIdx.x >> 4;
    for (int r = 0; r < 2; ++r) {
        const int y = tby * 32 + 2 * ty + r, x = tbx * 64 + 4 * tx;
        if (y >= H) return;
        const size_t o = ((size_t)y * W + x) * 3;
        float4* sp = reinterpret_cast<float4*>(state + o);
        sp[0] = make_float4(v, v, v, v); sp[1] = make_float4(v, v, v, v); sp[2] = make_float4(v, v, v, v);
        uint32_t* op = reinterpret_cast<uint32_t*>(out + o);
        op[0] = 1; op[1] = 2; op[2] = 3;
    }
}
// pattern B: fully coalesced (consecutive threads write consecutive 16-byte / 4-byte words)
__global__ void k_coal(float* __restrict__ state, uint8_t* __restrict__ out, float v) {
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x, n4 = (size_t)W * H * 3 / 4, nthreads = (size_t)gridDim.x * blockDim.x;
    for (size_t j = i; j < n4; j += nthreads) reinterpret_cast<float4*>(state)[j] = make_float4(v, v, v, v);
    for (size_t j = i; j < n4; j += nthreads) reinterpret_cast<uint32_t*>(out)[j] = 7;
}
int main() {
    float* state; uint8_t* out;
    const size_t px = (size_t)W * H;
    const int NBUF = 8;                                            // rotate buffers: larger than L2
    cudaMalloc(&state, px * 12 * NBUF); cudaMalloc(&out, px * 3 * NBUF);
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    const double bytes = px * 15.0;
    for (int variant = 0; variant < 3; ++variant) {
        float best = 1e9f;
        for (int rep = 0; rep < 5; ++rep) {
            cudaEventRecord(e0);
            for (int b = 0; b < NBUF; ++b) {
                float* s = state + px * 3 * b; uint8_t* o = out + px * 3 * b;
                if (variant == 0) k_quad<<<(W / 4 * H + 255) / 256, 256>>>(s, o, 1.f);
                else if (variant == 1) k_tile<<<(W / 64) * ((H + 31) / 32), 256>>>(s, o, 1.f);
                else k_coal<<<148 * 8, 256>>>(s, o, 1.f);
            }
            cudaEventRecord(e1); cudaEventSynchronize(e1);
            float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms;
        }
        printf("%s: %.1f us per 4K frame of stores (15 B/px) = %.2f TB/s\n", variant == 0 ? "quad rows   " : variant == 1 ? "tile pattern" : "coalesced   ",
               best * 1000 / NBUF, bytes * NBUF / (best * 1e-3) / 1e12);
    }
    printf("err: %s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
