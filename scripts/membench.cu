// scripts/membench.cu -- calibration microbenchmark (not product code): what does B200 HBM sustain for
// write-only / copy / multi-plane-write access patterns at the byte counts of our workloads, timed exactly
// like bench.py (events around back-to-back launches, rotating buffers so nothing stays in L2)?
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o build/membench scripts/membench.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e), __FILE__, __LINE__); exit(1);} } while (0)

template <bool CS>
__global__ void write_planes(float4* __restrict__ out, size_t plane_vec, int planes, size_t n_vec) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vec) return;
    float4 v = make_float4((float)i, 1.f, 2.f, 3.f);
    for (int p = 0; p < planes; ++p) {
        if (CS) __stcs(out + p * plane_vec + i, v); else out[p * plane_vec + i] = v;
        v.x += 1.f;
    }
}

__global__ void write_planes_pdl(float4* __restrict__ out, size_t plane_vec, int planes, size_t n_vec) {
    asm volatile("griddepcontrol.launch_dependents;");
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_vec) {
        float4 v = make_float4((float)i, 1.f, 2.f, 3.f);
        for (int p = 0; p < planes; ++p) { __stcs(out + p * plane_vec + i, v); v.x += 1.f; }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

__global__ void copy_k(const float4* __restrict__ in, float4* __restrict__ out, size_t n_vec) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_vec) out[i] = in[i];
}

// read 1 plane, write `planes` planes (our traffic shape)
__global__ void read1_write_planes(const float4* __restrict__ in, float4* __restrict__ out, size_t plane_vec, int planes, size_t n_vec) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= n_vec) return;
    float4 v = in[i];
    for (int p = 0; p < planes; ++p) { __stcs(out + p * plane_vec + i, v); v.x += 1.f; }
}

__global__ void read1_write_planes_pdl(const float4* __restrict__ in, float4* __restrict__ out, size_t plane_vec, int planes, size_t n_vec) {
    asm volatile("griddepcontrol.launch_dependents;");
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_vec) {
        float4 v = in[i];
        for (int p = 0; p < planes; ++p) { __stcs(out + p * plane_vec + i, v); v.x += 1.f; }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// same, and every CTA first pulls a slice of the NEXT frame's input into L2 (prefetch.global.L2)
__global__ void read1_write_planes_pdl_pf(const float4* __restrict__ in, const float4* __restrict__ next_in, float4* __restrict__ out,
                                          size_t plane_vec, int planes, size_t n_vec) {
    asm volatile("griddepcontrol.launch_dependents;");
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n_vec) {
        if ((threadIdx.x & 7) == 0) asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(next_in + i));
        float4 v = in[i];
        for (int p = 0; p < planes; ++p) { __stcs(out + p * plane_vec + i, v); v.x += 1.f; }
    }
    asm volatile("griddepcontrol.wait;" ::: "memory");
}

// bulk L2 prefetch of one plane by a small grid (one 128-byte line per thread)
__global__ void l2_prefetch(const char* __restrict__ p, size_t bytes) {
    asm volatile("griddepcontrol.launch_dependents;");
    size_t i = ((size_t)blockIdx.x * blockDim.x + threadIdx.x) * 128;
    if (i < bytes) asm volatile("prefetch.global.L2::evict_last [%0];" :: "l"(p + i));
}

int main(int argc, char** argv) {
    const size_t px_list[] = {2073600 * 4 / 3, 8294400 * 4 / 3, 33177600 * 4 / 3};   // ~ sum over octaves of c2, c3, c4 pixels
    const char* names[] = {"c2-like", "c3-like", "c4-like"};
    cudaStream_t st; CK(cudaStreamCreate(&st));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    for (int w = 0; w < 3; ++w) {
        const size_t px = px_list[w] / 4 * 4, n_vec = px / 4;
        const int planes = 11;
        const size_t frame_bytes = px * 4 * planes;
        const int slots = (int)((600ull << 20) / frame_bytes) + 2;
        float4* out; CK(cudaMalloc(&out, frame_bytes * slots));
        float4* in; CK(cudaMalloc(&in, px * 4 * slots)); CK(cudaMemset(in, 0, px * 4 * slots));
        const int reps = w == 0 ? 1000 : (w == 1 ? 300 : 100);
        const int block = 256; const unsigned grid = (unsigned)((n_vec + block - 1) / block);
        auto time_it = [&](const char* what, auto launch, double bytes) {
            for (int i = 0; i < 20; ++i) launch(i % slots);
            CK(cudaStreamSynchronize(st));
            CK(cudaEventRecord(e0, st));
            for (int i = 0; i < reps; ++i) launch(i % slots);
            CK(cudaEventRecord(e1, st));
            CK(cudaStreamSynchronize(st));
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1));
            const double us = ms * 1e3 / reps;
            printf("%-8s %-34s %9.2f us  %8.1f GB/s\n", names[w], what, us, bytes / us / 1e3);
        };
        time_it("cudaMemsetAsync (write only)", [&](int s) { CK(cudaMemsetAsync((char*)out + (size_t)s * frame_bytes, 1, frame_bytes, st)); }, (double)frame_bytes);
        time_it("write 11 planes, default st", [&](int s) { write_planes<false><<<grid, block, 0, st>>>(out + (size_t)s * n_vec * planes, n_vec, planes, n_vec); }, (double)frame_bytes);
        time_it("write 11 planes, st.cs", [&](int s) { write_planes<true><<<grid, block, 0, st>>>(out + (size_t)s * n_vec * planes, n_vec, planes, n_vec); }, (double)frame_bytes);
        time_it("write 11 planes, st.cs, PDL", [&](int s) {
            cudaLaunchConfig_t cfg{}; cfg.gridDim = dim3(grid); cfg.blockDim = dim3(block); cfg.stream = st;
            cudaLaunchAttribute a[1]; a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = a; cfg.numAttrs = 1;
            CK(cudaLaunchKernelEx(&cfg, write_planes_pdl, out + (size_t)s * n_vec * planes, n_vec, planes, n_vec)); }, (double)frame_bytes);
        time_it("read 1 + write 11 planes, st.cs", [&](int s) { read1_write_planes<<<grid, block, 0, st>>>(in + (size_t)s * n_vec, out + (size_t)s * n_vec * planes, n_vec, planes, n_vec); }, (double)frame_bytes + px * 4.0);
        auto pdl_cfg = [&](cudaLaunchConfig_t& cfg, cudaLaunchAttribute* a, unsigned g) {
            cfg = cudaLaunchConfig_t{}; cfg.gridDim = dim3(g); cfg.blockDim = dim3(block); cfg.stream = st;
            a[0].id = cudaLaunchAttributeProgrammaticStreamSerialization; a[0].val.programmaticStreamSerializationAllowed = 1;
            cfg.attrs = a; cfg.numAttrs = 1; };
        time_it("read 1 + write 11, PDL", [&](int s) {
            cudaLaunchConfig_t cfg; cudaLaunchAttribute a[1]; pdl_cfg(cfg, a, grid);
            CK(cudaLaunchKernelEx(&cfg, read1_write_planes_pdl, (const float4*)(in + (size_t)s * n_vec), out + (size_t)s * n_vec * planes, n_vec, planes, n_vec)); }, (double)frame_bytes + px * 4.0);
        time_it("read 1 (L2-resident) + write 11, PDL", [&](int s) {
            cudaLaunchConfig_t cfg; cudaLaunchAttribute a[1]; pdl_cfg(cfg, a, grid);
            CK(cudaLaunchKernelEx(&cfg, read1_write_planes_pdl, (const float4*)in, out + (size_t)s * n_vec * planes, n_vec, planes, n_vec)); }, (double)frame_bytes + px * 4.0);
        time_it("read 1 + write 11, PDL, in-kernel L2 prefetch of next", [&](int s) {
            cudaLaunchConfig_t cfg; cudaLaunchAttribute a[1]; pdl_cfg(cfg, a, grid);
            CK(cudaLaunchKernelEx(&cfg, read1_write_planes_pdl_pf, (const float4*)(in + (size_t)s * n_vec), (const float4*)(in + (size_t)((s + 1) % slots) * n_vec),
                                  out + (size_t)s * n_vec * planes, n_vec, planes, n_vec)); }, (double)frame_bytes + px * 4.0);
        time_it("read 1 + write 11, PDL, separate L2 prefetch kernel", [&](int s) {
            cudaLaunchConfig_t cfg; cudaLaunchAttribute a[1];
            const size_t bytes = px * 4; pdl_cfg(cfg, a, (unsigned)((bytes / 128 + block - 1) / block));
            CK(cudaLaunchKernelEx(&cfg, l2_prefetch, (const char*)(in + (size_t)((s + 1) % slots) * n_vec), bytes));
            pdl_cfg(cfg, a, grid);
            CK(cudaLaunchKernelEx(&cfg, read1_write_planes_pdl, (const float4*)(in + (size_t)s * n_vec), out + (size_t)s * n_vec * planes, n_vec, planes, n_vec)); }, (double)frame_bytes + px * 4.0);
        time_it("write 1 plane x11 launches-equivalent", [&](int s) { write_planes<true><<<grid, block, 0, st>>>(out + (size_t)s * n_vec * planes, n_vec, 1, n_vec); }, (double)px * 4);
        const size_t cp_vec = frame_bytes / 2 / 16;
        const unsigned cgrid = (unsigned)((cp_vec + block - 1) / block);
        time_it("copy kernel (read+write bytes)", [&](int s) { copy_k<<<cgrid, block, 0, st>>>(out + (size_t)((s + 1) % slots) * n_vec * planes, out + (size_t)s * n_vec * planes, cp_vec); }, (double)cp_vec * 32);
        time_it("cudaMemcpyAsync D2D (r+w bytes)", [&](int s) { CK(cudaMemcpyAsync(out + (size_t)s * n_vec * planes, out + (size_t)((s + 1) % slots) * n_vec * planes, frame_bytes / 2, cudaMemcpyDeviceToDevice, st)); }, (double)frame_bytes);
        time_it("empty-ish launch (1 plane, 1 block)", [&](int s) { write_planes<true><<<1, block, 0, st>>>(out, n_vec, 1, 256); }, 4096.0);
        CK(cudaFree(out)); CK(cudaFree(in));
    }
    return 0;
}
