// scratch probe: one 3-D u8 tensor-map load and one 2-D f32 load, checks contents and byte counts
#include <cstdio>
#include <cstdlib>
#include <vector>
#include "crt_tma.cuh"
using namespace crt;
__global__ void k(const __grid_constant__ CUtensorMap mi, const __grid_constant__ CUtensorMap ms, int c0, int c1, int c2, int s0, int s1, int mode,
                  uint8_t* out_raw, float* out_st, int* status) {
    extern __shared__ __align__(128) unsigned char dsm[];
    __shared__ __align__(8) uint64_t bar[2];
    float* s_state = reinterpret_cast<float*>(dsm);
    uint8_t* s_raw = dsm + 24576;
    if (threadIdx.x == 0) {
        mbar_init(&bar[0], 1); mbar_init(&bar[1], 1); fence_mbar_init();
        if (mode & 1) { tma_prefetch_desc(&mi); tma_prefetch_desc(&ms); }
        if (mode & 2) { mbar_expect_tx(&bar[0], 224 * 18); if (mode & 16) tma_load_2d(s_raw, &mi, c0, c1 + c2 * 48, &bar[0]); else tma_load_3d(s_raw, &mi, c0, c1, c2, &bar[0]); }
        if (mode & 4) { mbar_expect_tx(&bar[1], 24576); tma_load_2d(s_state, &ms, s0, s1, &bar[1]); }
        if (mode & 8) printf("smem base %u raw %u\n", smem_u32(dsm), smem_u32(s_raw));
    }
    __syncthreads();
    bool ok0 = false, ok1 = false;
    for (int i = 0; i < (1 << 22) && !ok0; ++i) ok0 = mbar_try_wait(&bar[0], 0);
    for (int i = 0; i < (1 << 22) && !ok1; ++i) ok1 = mbar_try_wait(&bar[1], 0);
    if (threadIdx.x == 0) { status[0] = ok0; status[1] = ok1; }
    if (ok0) for (int i = threadIdx.x; i < 224 * 18; i += blockDim.x) out_raw[i] = s_raw[i];
    if (ok1) for (int i = threadIdx.x; i < 6144; i += blockDim.x) out_st[i] = s_state[i];
}
int main(int argc, char** argv) {
    int mode = argc > 1 ? atoi(argv[1]) : 7;
    const int W = 128, H = 96, N = 3;
    std::vector<uint8_t> h(N * H * W * 3);
    for (size_t i = 0; i < h.size(); ++i) h[i] = (uint8_t)(i * 7 + (i >> 8));
    std::vector<float> hs(H * W * 3);
    for (size_t i = 0; i < hs.size(); ++i) hs[i] = (float)i;
    uint8_t *d_in, *d_raw; float *d_st, *d_ost; int* d_status;
    cudaMalloc(&d_in, h.size()); cudaMalloc(&d_st, hs.size() * 4); cudaMalloc(&d_raw, 4096); cudaMalloc(&d_ost, 24576); cudaMalloc(&d_status, 8);
    cudaMemcpy(d_in, h.data(), h.size(), cudaMemcpyHostToDevice); cudaMemcpy(d_st, hs.data(), hs.size() * 4, cudaMemcpyHostToDevice);
    CUtensorMap mi, ms;
    const uint64_t W3 = W * 3;
    const uint64_t id[3] = {W3, H / 2, N}, is[2] = {2 * W3, W3 * H}; const uint32_t ib[3] = {224, 18, 1};
    const uint64_t sd[2] = {W3, H}, ss[1] = {W3 * 4}; const uint32_t sb[2] = {192, 32};
    bool e0;
    if (mode & 16) { const uint64_t id2[2] = {W3, (uint64_t)N * H / 2}; e0 = tma_encode(&mi, CU_TENSOR_MAP_DATA_TYPE_UINT8, 2, d_in, id2, is, ib); }
    else e0 = tma_encode(&mi, CU_TENSOR_MAP_DATA_TYPE_UINT8, 3, d_in, id, is, ib);
    bool e1 = tma_encode(&ms, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, d_st, sd, ss, sb);
    printf("encode %d %d\n", e0, e1);
    cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, 32768);
    int cases[3][5] = {{-9, -1, 1, 0, 0}, {183, 15, 2, 192, 64}, {6, 3, 0, 192, 32}};
    const int only = argc > 2 ? atoi(argv[2]) : -1; int ci = -1;
    for (auto& c : cases) {
        if (++ci != only && only >= 0) continue;
        if (mode & 32) c[0] &= ~15;
        cudaMemset(d_raw, 0xEE, 4096);
        k<<<1, 256, 32768>>>(mi, ms, c[0], c[1], c[2], c[3], c[4], mode, d_raw, d_ost, d_status);
        cudaError_t e = cudaDeviceSynchronize();
        int st[2]; cudaMemcpy(st, d_status, 8, cudaMemcpyDeviceToHost);
        std::vector<uint8_t> r(4096); cudaMemcpy(r.data(), d_raw, 4096, cudaMemcpyDeviceToHost);
        std::vector<float> o(6144); cudaMemcpy(o.data(), d_ost, 24576, cudaMemcpyDeviceToHost);
        int bad = 0, badst = 0;
        for (int row = 0; row < 18; ++row) for (int x = 0; x < 224; ++x) {
            int gy = c[1] + row, gx = c[0] + x; uint8_t want = 0;
            if (gy >= 0 && gy < H / 2 && gx >= 0 && gx < (int)W3) want = h[(size_t)c[2] * H * W3 + (size_t)(2 * gy) * W3 + gx];
            bad += r[row * 224 + x] != want;
        }
        for (int row = 0; row < 32; ++row) for (int x = 0; x < 192; ++x) {
            int gy = c[4] + row, gx = c[3] + x; float want = 0.f;
            if (gy < H && gx < (int)W3) want = hs[(size_t)gy * W3 + gx];
            badst += o[row * 192 + x] != want;
        }
        printf("case (%d,%d,%d | %d,%d): err=%s landed in=%d st=%d mismatches in=%d st=%d\n", c[0], c[1], c[2], c[3], c[4], cudaGetErrorString(e), st[0], st[1], bad, badst);
    }
    return 0;
}
