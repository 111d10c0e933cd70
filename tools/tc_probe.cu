// tc_probe.cu — hardware probe for the tcgen05 operand-layout facts the tensor-core kernels of this repo rely on.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O2 -std=c++17 -I recnext_b200/csrc -o tools/bin/tc_probe tools/tc_probe.cu
// Each case builds a shared-memory IMAGE of A and B on the host according to a layout hypothesis, runs ksteps
// tcgen05.mma (M = 128, K = 16 each) and compares the TMEM accumulator with the exact product.  Then it times the
// MMA issue / completion rate, TMEM loads and two concurrent issuers.  Results: gpurun_out/tc_probe.txt.
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include "tc05.cuh"

using namespace recnext;

struct Case {
    int N, ksteps;
    uint32_t a_off, b_off;
    uint32_t a_lbo, a_sbo, b_lbo, b_sbo;
    uint32_t a_kadv, b_kadv;
    int a_mn, b_mn;
    int a_bytes, b_bytes;
};

static constexpr int kAImg = 96 * 1024, kBImg = 96 * 1024;

__global__ void __launch_bounds__(128, 1) probe_kernel(const uint8_t* __restrict__ aimg, const uint8_t* __restrict__ bimg, Case c, float* __restrict__ out,
                                                         float* __restrict__ out2, uint32_t* __restrict__ status) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar_mem;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < c.a_bytes / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = reinterpret_cast<const uint4*>(aimg)[i];
    for (int i = tid; i < c.b_bytes / 16; i += 128) reinterpret_cast<uint4*>(smem + kAImg)[i] = reinterpret_cast<const uint4*>(bimg)[i];
    const uint32_t bar = tc::smem_u32(&bar_mem);
    if (tid == 0) { tc::mbar_init(bar, 1); tc::mbar_init_fence(); }
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_slot), 512);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    if (tid == 0) {
        const uint32_t idesc = tc::make_idesc(128, c.N, 1, c.a_mn, c.b_mn);
        const uint64_t ad = tc::make_sdesc(tc::smem_u32(smem) + c.a_off, c.a_lbo, c.a_sbo);
        const uint64_t bd = tc::make_sdesc(tc::smem_u32(smem + kAImg) + c.b_off, c.b_lbo, c.b_sbo);
        for (int ks = 0; ks < c.ksteps; ++ks) tc::mma_ss(tbase, tc::sdesc_advance(ad, ks * c.a_kadv), tc::sdesc_advance(bd, ks * c.b_kadv), idesc, ks > 0);
        tc::mma_commit(bar);
    }
    bool ok = false;
    for (int it = 0; it < (1 << 22); ++it) if (tc::mbar_try_wait(bar, 0)) { ok = true; break; }
    if (!ok) { if (tid == 0) status[0] = 1; }
    tc::fence_after_sync();
    if (ok) {
        const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
        for (int c0 = 0; c0 < c.N; c0 += 8) {
            uint32_t v[8];
            tc::tmem_ld8(trow + c0, v);
            tc::tmem_ld_wait();
            for (int j = 0; j < 8; ++j) out[(warp * 32 + lane) * c.N + c0 + j] = __uint_as_float(v[j]);
        }
        uint32_t v[8];
        tc::tmem_ld8(trow + 2, v);   // unaligned column start
        tc::tmem_ld_wait();
        for (int j = 0; j < 8; ++j) out2[(warp * 32 + lane) * 8 + j] = __uint_as_float(v[j]);
        if (tid == 0) status[1] = tbase;
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

// timing: mode 0 = one issuer, nrep MMAs of (128 x N x 16); mode 1 = two issuers (warps 0 and 1), nrep each, disjoint columns;
// mode 2 = nrep rounds of 4 warps reading all `N` columns with tcgen05.ld x32; mode 3: nrep x (1 MMA -> commit -> wait) round trips
__global__ void __launch_bounds__(128, 1) time_kernel(int N, int nrep, int mode, long long* __restrict__ res, uint32_t a_sbo, uint32_t a_lbo) {
    extern __shared__ __align__(1024) uint8_t smem[];
    __shared__ uint32_t tmem_slot;
    __shared__ __align__(8) uint64_t bar_mem[2];
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    for (int i = tid; i < (kAImg + kBImg) / 16; i += 128) reinterpret_cast<uint4*>(smem)[i] = make_uint4(0, 0, 0, 0);
    const uint32_t bar0 = tc::smem_u32(&bar_mem[0]), bar1 = tc::smem_u32(&bar_mem[1]);
    if (tid == 0) { tc::mbar_init(bar0, 1); tc::mbar_init(bar1, 1); tc::mbar_init_fence(); }
    if (warp == 0) tc::tmem_alloc(tc::smem_u32(&tmem_slot), 512);
    tc::fence_proxy_async();
    tc::fence_before_sync();
    __syncthreads();
    tc::fence_after_sync();
    const uint32_t tbase = tmem_slot;
    const uint32_t idesc = tc::make_idesc(128, N, 1, 0, 0);
    const uint64_t ad = tc::make_sdesc(tc::smem_u32(smem), a_lbo, a_sbo);
    const uint64_t bd = tc::make_sdesc(tc::smem_u32(smem + kAImg), 16 * N, 128);
    long long t0 = 0, t1 = 0, t2 = 0;
    const int uwarp = __shfl_sync(0xffffffffu, warp, 0);   // warp-uniform for the compiler: no per-lane waterfall around UTCHMMA
    if (mode == 0 || mode == 1) {
        if (uwarp == 0 || (mode == 1 && uwarp == 1)) {
            const uint32_t bar = uwarp == 0 ? bar0 : bar1;
            const uint32_t d = tbase + (uwarp == 0 ? 0 : 256);
            uint32_t elected = 0;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
            t0 = clock64();
            for (int i = 0; i < nrep; ++i) {
                const uint64_t a_i = tc::sdesc_advance(ad, (i & 7) * 16);
                if (elected) tc::mma_ss(d + (i & 1) * N, a_i, bd, idesc, 1);
            }
            if (elected) tc::mma_commit(bar);
            t1 = clock64();
            bool ok = false;
            for (int it = 0; it < (1 << 24); ++it) if (tc::mbar_try_wait(bar, 0)) { ok = true; break; }
            t2 = clock64();
            if (lane == 0) { res[uwarp * 4 + 0] = t1 - t0; res[uwarp * 4 + 1] = t2 - t0; res[uwarp * 4 + 2] = ok ? 1 : 0; }
        }
    } else if (mode == 2) {
        const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
        uint32_t acc = 0;
        __syncthreads();
        t0 = clock64();
        for (int i = 0; i < nrep; ++i) {
            for (int c0 = 0; c0 < N; c0 += 32) {
                uint32_t v[32];
                tc::tmem_ld32(trow + c0, v);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= v[j];
            }
        }
        t1 = clock64();
        if (lane == 0) { res[warp * 4 + 0] = t1 - t0; res[warp * 4 + 1] = acc; res[warp * 4 + 2] = 1; }
    } else if (mode == 4) {
        const uint32_t trow = tbase + ((uint32_t)(warp * 32) << 16);
        uint32_t acc = 0;
        __syncthreads();
        t0 = clock64();
        for (int i = 0; i < nrep; ++i) {
            for (int c0 = 0; c0 < N; c0 += 128) {
                uint32_t v0[32], v1[32], v2[32], v3[32];
                tc::tmem_ld32(trow + c0, v0);
                tc::tmem_ld32(trow + c0 + 32, v1);
                tc::tmem_ld32(trow + c0 + 64, v2);
                tc::tmem_ld32(trow + c0 + 96, v3);
                tc::tmem_ld_wait();
#pragma unroll
                for (int j = 0; j < 32; ++j) acc ^= v0[j] ^ v1[j] ^ v2[j] ^ v3[j];
            }
        }
        t1 = clock64();
        if (lane == 0) { res[warp * 4 + 0] = t1 - t0; res[warp * 4 + 1] = acc; res[warp * 4 + 2] = 1; }
    } else if (mode == 5 || mode == 6 || mode == 7) {
        // mode 5: MN-major B, idle neighbours; mode 6: warps 1-3 stream 16-byte stores into an unrelated region meanwhile;
        // mode 7: warps 1-3 stream 16-byte LOADS meanwhile
        const uint32_t idesc_mn = tc::make_idesc(128, N, 1, 0, 1);
        const uint64_t bdm = tc::make_sdesc(tc::smem_u32(smem + kAImg), 128, 64 * 16 + 16);
        __shared__ volatile int stop;
        if (tid == 0) stop = 0;
        __syncthreads();
        if (uwarp == 0) {
            uint32_t elected = 0;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
            t0 = clock64();
            for (int i = 0; i < nrep; ++i) {
                const uint64_t a_i = tc::sdesc_advance(ad, (i & 3) * 4096);
                const uint64_t b_i = tc::sdesc_advance(bdm, (i & 3) * 256);
                if (elected) tc::mma_ss(tbase + (i & 1) * N, a_i, b_i, idesc_mn, 1);
            }
            if (elected) tc::mma_commit(bar0);
            t1 = clock64();
            bool ok = false;
            for (int it = 0; it < (1 << 24); ++it) if (tc::mbar_try_wait(bar0, 0)) { ok = true; break; }
            t2 = clock64();
            if (lane == 0) { res[0] = t1 - t0; res[1] = t2 - t0; res[2] = ok ? 1 : 0; stop = 1; }
        } else if (mode != 5) {
            uint4* reg = reinterpret_cast<uint4*>(smem + kAImg + 80 * 1024) + (warp - 1) * 32 * 8 + lane;
            uint4 v = make_uint4(tid, 1, 2, 3);
            long long n = 0;
            while (!stop) {
                if (mode == 6) {
#pragma unroll
                    for (int u = 0; u < 8; ++u) reg[u * 32] = v;
                } else {
#pragma unroll
                    for (int u = 0; u < 8; ++u) { uint4 w = reg[u * 32]; v.x ^= w.x; }
                }
                n += 8;
            }
            if (lane == 0) res[warp * 4 + 3] = n + (v.x & 1);
        }
    } else if (mode == 8 || mode == 9 || mode == 10) {
        // issue-side cost of the hand-offs around a group of 4 MMAs (N columns): mode 8: 4 MMAs + tcgen05.commit per group; mode 9: the same
        // plus a try_wait on an already completed mbarrier and a tcgen05.fence::after_thread_sync; mode 10: 4 MMAs only (one commit at the end)
        if (uwarp == 0) {
            uint32_t elected = 0;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
            if (elected) tc::mbar_arrive(bar1);   // bar1 completes its phase 0: try_wait(bar1, 0) succeeds immediately from now on
            __syncwarp();
            t0 = clock64();
            for (int i = 0; i < nrep; ++i) {
                if (mode == 9) { tc::mbar_wait(bar1, 0); tc::fence_after_sync(); }
                for (int j = 0; j < 4; ++j) {
                    const uint64_t a_i = tc::sdesc_advance(ad, (uint32_t)j * 4096u);
                    if (elected) tc::mma_ss(tbase + (i & 1) * N, a_i, bd, idesc, 1);
                }
                if (mode != 10 && elected) tc::mma_commit(bar0);
            }
            if (mode == 10 && elected) tc::mma_commit(bar0);
            t1 = clock64();
            if (lane == 0) { res[0] = t1 - t0; res[1] = t1 - t0; res[2] = 1; }
        }
    } else if (mode == 11) {
        // the channel mixer's GEMM1 operand walk: A = four 16 KB K-major weight tiles (ring slots), B = MN-major activation tile with
        // chunk pitch `a_sbo` bytes (C * 16 + 16 in ffn_tc.cu), 4 K steps per slot
        if (uwarp == 0) {
            uint32_t elected = 0;
            asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(elected));
            const uint32_t idesc_mn = tc::make_idesc(128, N, 1, 0, 1);
            t0 = clock64();
            for (int i = 0; i < nrep; ++i) {
                const int kc = i & 3;
                const uint64_t a0 = tc::make_sdesc(tc::smem_u32(smem) + kc * 16384, 2048, 128);
                const uint64_t b0 = tc::make_sdesc(tc::smem_u32(smem) + 65536 + (uint32_t)(kc * 64) * 16u, 128, a_sbo);
                for (int j = 0; j < 4; ++j)
                    if (elected) tc::mma_ss(tbase + (i & 1) * N, tc::sdesc_advance(a0, (uint32_t)j * 4096u), tc::sdesc_advance(b0, (uint32_t)j * 256u), idesc_mn, 1);
                if (elected) tc::mma_commit(bar1);
            }
            if (elected) tc::mma_commit(bar0);
            t1 = clock64();
            bool ok = false;
            for (int it = 0; it < (1 << 24); ++it) if (tc::mbar_try_wait(bar0, 0)) { ok = true; break; }
            t2 = clock64();
            if (lane == 0) { res[0] = t1 - t0; res[1] = t2 - t0; res[2] = ok ? 1 : 0; }
        }
    } else if (mode == 12 || mode == 13) {
        // wake-up latency of an mbarrier wait: warp 1 arrives after `nrep` clocks; warp 0 waits (mode 12: try_wait with the suspend-time
        // hint of tc05.cuh; mode 13: try_wait without a hint in a spin loop) and stamps the clock when it gets through
        __shared__ long long t_arrive;
        if (uwarp == 1) {
            const long long s0 = clock64();
            while (clock64() - s0 < nrep) { }
            if (lane == 0) { t_arrive = clock64(); tc::mbar_arrive(bar0); }
        } else if (uwarp == 0) {
            if (mode == 12) tc::mbar_wait(bar0, 0);
            else {
                uint32_t ok = 0;
                while (!ok) asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(ok) : "r"(bar0), "r"(0) : "memory");
            }
            t1 = clock64();
            __syncwarp();
            if (lane == 0) { res[0] = t1 - t_arrive; res[2] = 1; }
        }
    } else if (mode == 3) {
        if (tid == 0) {
            t0 = clock64();
            uint32_t ph = 0;
            bool ok = true;
            for (int i = 0; i < nrep && ok; ++i) {
                tc::mma_ss(tbase, ad, bd, idesc, 1);
                tc::mma_commit(bar0);
                ok = false;
                for (int it = 0; it < (1 << 22); ++it) if (tc::mbar_try_wait(bar0, ph)) { ok = true; break; }
                ph ^= 1;
            }
            t1 = clock64();
            res[0] = t1 - t0; res[1] = 0; res[2] = ok ? 1 : 0;
        }
    }
    tc::fence_before_sync();
    __syncthreads();
    if (warp == 0) tc::tmem_dealloc(tbase, 512);
}

static inline uint16_t f2bf(float f) { __nv_bfloat16 h = __float2bfloat16(f); uint16_t u; memcpy(&u, &h, 2); return u; }

static int g_fail = 0;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// A(m, k), B(k, n) logical accessors supplied by the case builder through explicit images
static void run_case(const char* name, const Case& c, const std::vector<uint8_t>& aimg, const std::vector<uint8_t>& bimg, const std::vector<float>& expect) {
    uint8_t *da, *db; float *dout, *dout2; uint32_t* dst;
    CK(cudaMalloc(&da, kAImg)); CK(cudaMalloc(&db, kBImg)); CK(cudaMalloc(&dout, 128 * 256 * 4)); CK(cudaMalloc(&dout2, 128 * 8 * 4)); CK(cudaMalloc(&dst, 16));
    CK(cudaMemset(da, 0, kAImg)); CK(cudaMemset(db, 0, kBImg)); CK(cudaMemset(dst, 0, 16)); CK(cudaMemset(dout, 0xff, 128 * 256 * 4));
    CK(cudaMemcpy(da, aimg.data(), aimg.size(), cudaMemcpyHostToDevice));
    CK(cudaMemcpy(db, bimg.data(), bimg.size(), cudaMemcpyHostToDevice));
    probe_kernel<<<1, 128, kAImg + kBImg>>>(da, db, c, dout, dout2, dst);
    cudaError_t e = cudaDeviceSynchronize();
    if (e != cudaSuccess) { printf("[%s] KERNEL ERROR %s\n", name, cudaGetErrorString(e)); g_fail++; exit(3); }
    std::vector<float> out(128 * c.N), out2(128 * 8);
    uint32_t st[4];
    CK(cudaMemcpy(out.data(), dout, out.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(out2.data(), dout2, out2.size() * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(st, dst, 16, cudaMemcpyDeviceToHost));
    double maxerr = 0; int bad = 0, firstbad = -1;
    for (int i = 0; i < 128 * c.N; ++i) { double d = fabs((double)out[i] - expect[i]); if (d > maxerr) maxerr = d; if (d > 1e-3) { if (firstbad < 0) firstbad = i; ++bad; } }
    int bad2 = 0;
    for (int m = 0; m < 128; ++m) for (int j = 0; j < 8 && 2 + j < c.N; ++j) if (fabs(out2[m * 8 + j] - expect[m * c.N + 2 + j]) > 1e-3) ++bad2;
    printf("[%s] N=%d ksteps=%d timeout=%u tmem_base=0x%x maxerr=%.4g bad=%d/%d unaligned_col_bad=%d", name, c.N, c.ksteps, st[0], st[1], maxerr, bad, 128 * c.N, bad2);
    if (bad) printf("  first bad (m=%d,n=%d): got %.4f want %.4f", firstbad / c.N, firstbad % c.N, out[firstbad], expect[firstbad]);
    printf("  -> %s\n", (bad == 0 && st[0] == 0) ? "PASS" : "FAIL");
    if (bad || st[0]) g_fail++;
    cudaFree(da); cudaFree(db); cudaFree(dout); cudaFree(dout2); cudaFree(dst);
}

static float rnd_small() { return (float)((rand() % 17) - 8) / 4.0f; }   // exactly representable, exact fp32 sums

int main() {
    CK(cudaFuncSetAttribute(probe_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAImg + kBImg));
    CK(cudaFuncSetAttribute(time_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAImg + kBImg));
    srand(1);
    // ---- case 1: K-major A and B, SWIZZLE_NONE, K = 32
    {
        const int N = 16, K = 32;
        std::vector<float> A(128 * K), B(K * N), E(128 * N, 0.f);
        for (auto& v : A) v = rnd_small();
        for (auto& v : B) v = rnd_small();
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[k * N + n]; E[m * N + n] = s; }
        Case c{}; c.N = N; c.ksteps = K / 16; c.a_sbo = 128; c.a_lbo = 128 * 16; c.b_sbo = 128; c.b_lbo = N * 16; c.a_kadv = 2 * c.a_lbo; c.b_kadv = 2 * c.b_lbo;
        std::vector<uint8_t> ai(128 * K * 2), bi(N * K * 2);
        for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) { uint16_t h = f2bf(A[m * K + k]); memcpy(&ai[(m / 8) * c.a_sbo + (k / 8) * c.a_lbo + (m % 8) * 16 + (k % 8) * 2], &h, 2); }
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { uint16_t h = f2bf(B[k * N + n]); memcpy(&bi[(n / 8) * c.b_sbo + (k / 8) * c.b_lbo + (n % 8) * 16 + (k % 8) * 2], &h, 2); }
        c.a_bytes = (int)ai.size(); c.b_bytes = (int)bi.size();
        run_case("kmajor_plain", c, ai, bi, E);
    }
    // ---- case 2: row-shifted A: slab of 160 rows x 4 strips (8 columns each, rows contiguous at 16 bytes), start row r
    for (int r : {1, 3, 4, 13}) {
        const int N = 16, K = 32, R = 160;
        std::vector<float> S(R * K), B(K * N), E(128 * N, 0.f);
        for (auto& v : S) v = rnd_small();
        for (auto& v : B) v = rnd_small();
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += S[(m + r) * K + k] * B[k * N + n]; E[m * N + n] = s; }
        Case c{}; c.N = N; c.ksteps = 2; c.a_sbo = 128; c.a_lbo = R * 16; c.b_sbo = 128; c.b_lbo = N * 16; c.a_kadv = 2 * c.a_lbo; c.b_kadv = 2 * c.b_lbo;
        c.a_off = r * 16;
        std::vector<uint8_t> ai(R * K * 2), bi(N * K * 2);
        for (int m = 0; m < R; ++m) for (int k = 0; k < K; ++k) { uint16_t h = f2bf(S[m * K + k]); memcpy(&ai[(k / 8) * c.a_lbo + m * 16 + (k % 8) * 2], &h, 2); }
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { uint16_t h = f2bf(B[k * N + n]); memcpy(&bi[(n / 8) * c.b_sbo + (k / 8) * c.b_lbo + (n % 8) * 16 + (k % 8) * 2], &h, 2); }
        c.a_bytes = (int)ai.size(); c.b_bytes = (int)bi.size();
        char nm[64]; snprintf(nm, sizeof nm, "row_shift_r%d", r);
        run_case(nm, c, ai, bi, E);
    }
    // ---- case 3: LBO = 16 bytes: the two K chunks of one MMA are the SAME strip at rows m and m + 1
    {
        const int N = 16, R = 160;
        std::vector<float> S(R * 8), B(16 * N), E(128 * N, 0.f);
        for (auto& v : S) v = rnd_small();
        for (auto& v : B) v = rnd_small();
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < 16; ++k) s += S[(m + 2 + k / 8) * 8 + (k % 8)] * B[k * N + n]; E[m * N + n] = s; }
        Case c{}; c.N = N; c.ksteps = 1; c.a_sbo = 128; c.a_lbo = 16; c.b_sbo = 128; c.b_lbo = N * 16; c.a_off = 2 * 16;
        std::vector<uint8_t> ai(R * 16), bi(N * 16 * 2);
        for (int m = 0; m < R; ++m) for (int k = 0; k < 8; ++k) { uint16_t h = f2bf(S[m * 8 + k]); memcpy(&ai[m * 16 + k * 2], &h, 2); }
        for (int n = 0; n < N; ++n) for (int k = 0; k < 16; ++k) { uint16_t h = f2bf(B[k * N + n]); memcpy(&bi[(n / 8) * c.b_sbo + (k / 8) * c.b_lbo + (n % 8) * 16 + (k % 8) * 2], &h, 2); }
        c.a_bytes = (int)ai.size(); c.b_bytes = (int)bi.size();
        run_case("lbo16_row_pair", c, ai, bi, E);
    }
    // ---- case 4: MN-major B (activations: K = channels, N = pixels contiguous), K-major A; N = 128, K = 64
    for (int pad : {0, 16}) {
        const int N = 128, K = 64;
        std::vector<float> A(128 * K), B(K * N), E(128 * N, 0.f);
        for (auto& v : A) v = rnd_small();
        for (auto& v : B) v = rnd_small();
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += A[m * K + k] * B[k * N + n]; E[m * N + n] = s; }
        Case c{}; c.N = N; c.ksteps = K / 16; c.a_sbo = 128; c.a_lbo = 128 * 16; c.a_kadv = 2 * c.a_lbo; c.b_mn = 1;
        c.b_lbo = 128; c.b_sbo = K * 16 + pad; c.b_kadv = 2 * c.b_lbo;
        std::vector<uint8_t> ai(128 * K * 2), bi((N / 8) * c.b_sbo);
        for (int m = 0; m < 128; ++m) for (int k = 0; k < K; ++k) { uint16_t h = f2bf(A[m * K + k]); memcpy(&ai[(m / 8) * c.a_sbo + (k / 8) * c.a_lbo + (m % 8) * 16 + (k % 8) * 2], &h, 2); }
        for (int n = 0; n < N; ++n) for (int k = 0; k < K; ++k) { uint16_t h = f2bf(B[k * N + n]); memcpy(&bi[(n / 8) * c.b_sbo + (k / 8) * c.b_lbo + (k % 8) * 16 + (n % 8) * 2], &h, 2); }
        c.a_bytes = (int)ai.size(); c.b_bytes = (int)bi.size();
        run_case(pad ? "mnmajor_B_padded_sbo" : "mnmajor_B", c, ai, bi, E);
    }
    // ---- case 5: MN-major A (transposed operand, e.g. weight gradients S^T G) and MN-major B, K = 32 rows
    {
        const int N = 16, K = 32;   // A(m, k) = S[k][m]: S has K rows of 128 columns (strips of 8 columns, rows at 16 bytes)
        std::vector<float> S(K * 128), G(K * N), E(128 * N, 0.f);
        for (auto& v : S) v = rnd_small();
        for (auto& v : G) v = rnd_small();
        for (int m = 0; m < 128; ++m) for (int n = 0; n < N; ++n) { float s = 0; for (int k = 0; k < K; ++k) s += S[k * 128 + m] * G[k * N + n]; E[m * N + n] = s; }
        Case c{}; c.N = N; c.ksteps = K / 16; c.a_mn = 1; c.b_mn = 1;
        c.a_lbo = 128; c.a_sbo = K * 16; c.a_kadv = 256; c.b_lbo = 128; c.b_sbo = K * 16; c.b_kadv = 256;
        std::vector<uint8_t> ai(16 * c.a_sbo), bi((N / 8) * c.b_sbo);
        for (int k = 0; k < K; ++k) for (int m = 0; m < 128; ++m) { uint16_t h = f2bf(S[k * 128 + m]); memcpy(&ai[(m / 8) * c.a_sbo + k * 16 + (m % 8) * 2], &h, 2); }
        for (int k = 0; k < K; ++k) for (int n = 0; n < N; ++n) { uint16_t h = f2bf(G[k * N + n]); memcpy(&bi[(n / 8) * c.b_sbo + k * 16 + (n % 8) * 2], &h, 2); }
        c.a_bytes = (int)ai.size(); c.b_bytes = (int)bi.size();
        run_case("mnmajor_A_and_B", c, ai, bi, E);
    }
    // ---- timing
    long long* dres; CK(cudaMalloc(&dres, 64 * 8));
    auto timeit = [&](const char* name, int N, int nrep, int mode, uint32_t sbo, uint32_t lbo) {
        CK(cudaMemset(dres, 0, 64 * 8));
        time_kernel<<<1, 128, kAImg + kBImg>>>(N, nrep, mode, dres, sbo, lbo);
        cudaError_t e = cudaDeviceSynchronize();
        if (e != cudaSuccess) { printf("[%s] KERNEL ERROR %s\n", name, cudaGetErrorString(e)); exit(4); }
        long long r[16]; CK(cudaMemcpy(r, dres, sizeof r, cudaMemcpyDeviceToHost));
        if (mode == 12 || mode == 13) printf("[time %s] waited %d clocks, woke %lld clocks after the arrive\n", name, nrep, r[0]);
        else if (mode == 11) printf("[time %s] N=%d groups=%d B chunk pitch %u: total %lld cyc (%.1f per MMA) ok=%lld\n", name, N, nrep, sbo, r[1], (double)r[1] / (4.0 * nrep), r[2]);
        else if (mode >= 8) printf("[time %s] N=%d groups=%d issue time %lld cyc (%.1f per group of 4 MMAs; math floor %.0f)\n", name, N, nrep, r[0], (double)r[0] / nrep, 4.0 * (N >= 128 ? N / 2.0 : (N == 64 ? 48.0 : 39.0)));
        else if (mode >= 5) printf("[time %s] N=%d nrep=%d total=%lld cyc (%.1f / MMA) ok=%lld  neighbour 16-byte accesses per warp: %lld %lld %lld (%.1f B/cyc)\n", name, N, nrep, r[1], (double)r[1] / nrep, r[2], r[7], r[11], r[15], 16.0 * 32 * (r[7] + r[11] + r[15]) / (double)r[1]);
        else if (mode == 0 || mode == 3) printf("[time %s] N=%d nrep=%d issue=%lld cyc total=%lld cyc (%.1f / MMA) ok=%lld\n", name, N, nrep, r[0], r[1] ? r[1] : r[0], (double)(r[1] ? r[1] : r[0]) / nrep, r[2]);
        else if (mode == 1) printf("[time %s] N=%d nrep=%d x2 issuers: w0 total=%lld w1 total=%lld (%.1f cyc / MMA overall) ok=%lld,%lld\n", name, N, nrep, r[1], r[5], (double)(r[1] > r[5] ? r[1] : r[5]) / (2.0 * nrep), r[2], r[6]);
        else printf("[time %s] cols=%d nrep=%d per-warp cycles %lld %lld %lld %lld -> %.1f B/cyc/SM\n", name, N, nrep, r[0], r[4], r[8], r[12], 4.0 * 32 * N * 4 * nrep / (double)r[0]);
    };
    for (int rep = 0; rep < 2; ++rep) {
        timeit("mma_n16", 16, 2048, 0, 128, 2048);
        timeit("mma_n16_lbo16", 16, 2048, 0, 128, 16);
        timeit("mma_n32", 32, 2048, 0, 128, 2048);
        timeit("mma_n64", 64, 1024, 0, 128, 2048);
        timeit("mma_n128", 128, 1024, 0, 128, 2048);
        timeit("mma_n256", 256, 512, 0, 128, 2048);
        timeit("two_issuers_n16", 16, 2048, 1, 128, 2048);
        timeit("two_issuers_n128", 128, 1024, 1, 128, 2048);
        timeit("mma_n128_mnB", 128, 1024, 5, 128, 2048);
        timeit("mma_n128_mnB_plus_sts", 128, 1024, 6, 128, 2048);
        timeit("mma_n128_mnB_plus_lds", 128, 1024, 7, 128, 2048);
        timeit("mma_n256_mnB", 256, 512, 5, 128, 2048);
        timeit("mma_n256_mnB_plus_sts", 256, 512, 6, 128, 2048);
        timeit("4mma_n128_only", 128, 256, 10, 128, 2048);
        timeit("4mma_n128_commit", 128, 256, 8, 128, 2048);
        timeit("4mma_n128_wait_fence_commit", 128, 256, 9, 128, 2048);
        timeit("4mma_n64_commit", 64, 256, 8, 128, 2048);
        timeit("gemm1_walk_pitch1040", 128, 256, 11, 1040, 0);
        timeit("gemm1_walk_pitch4112", 128, 256, 11, 4112, 0);
        timeit("gemm1_walk_pitch4096", 128, 256, 11, 4096, 0);
        timeit("gemm1_walk_pitch4224", 128, 256, 11, 4224, 0);
        timeit("gemm1_walk_pitch2064", 128, 256, 11, 2064, 0);
        timeit("gemm1_walk_n64_pitch8208", 64, 256, 11, 8208, 0);
        for (int d : {300, 3000, 30000, 300000}) { timeit("wake_hint", 16, d, 12, 0, 0); timeit("wake_spin", 16, d, 13, 0, 0); }
        timeit("tmem_ld", 512, 64, 2, 128, 2048);
        timeit("tmem_ld_pipelined4", 512, 64, 4, 128, 2048);
        timeit("roundtrip_n16", 16, 256, 3, 128, 2048);
        timeit("roundtrip_n128", 128, 256, 3, 128, 2048);
    }
    printf("probe done, failures=%d\n", g_fail);
    return 0;   // layout hypotheses that fail are information, not errors
}
