// peaks.cu -- own-measured denominators that MEASURED_PEAKS.json does not carry:
//   * L2-resident streaming read bandwidth (a 48 MiB buffer read repeatedly by a full grid, 16-byte loads)
//   * FP64 issue rate without FMA (the traversal kernels are compiled -fmad=false): DADD + DMUL chains
//   * dependent-load latency through L2 (pointer chase over a 64 MiB ring, one lane)
// Build:  nvcc -gencode arch=compute_100a,code=sm_100a -O3 -fmad=false -o peaks bench/peaks.cu     Run:  ./peaks
// Prints one JSON object.  Events on the launching stream, warm-up first, best of 5.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <vector>

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { fprintf(stderr, "%s: %s\n", #x, cudaGetErrorString(e_)); exit(1); } } while (0)

__global__ void k_l2_read(const double2* __restrict__ buf, size_t n, int reps, double* sink) {
    double acc = 0;
    const size_t stride = (size_t)gridDim.x * blockDim.x;
    for (int r = 0; r < reps; r++)
        for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += stride) {
            double2 v = __ldcg(buf + i);  // L2 only: bypass L1 so that the number is the L2's
            acc += v.x + v.y;
        }
    if (acc == 12345.678) *sink = acc;
}
__global__ void k_fp64(int iters, double* sink) {
    double a0 = threadIdx.x * 1e-3, a1 = a0 + 1, a2 = a0 + 2, a3 = a0 + 3, a4 = a0 + 4, a5 = a0 + 5, a6 = a0 + 6, a7 = a0 + 7;
    const double m = 1.0000001, c = 1e-9;
    for (int i = 0; i < iters; i++) {  // 8 independent chains of one DMUL + one DADD (no FMA at -fmad=false)
        a0 = a0 * m + c; a1 = a1 * m + c; a2 = a2 * m + c; a3 = a3 * m + c;
        a4 = a4 * m + c; a5 = a5 * m + c; a6 = a6 * m + c; a7 = a7 * m + c;
    }
    double s = a0 + a1 + a2 + a3 + a4 + a5 + a6 + a7;
    if (s == 12345.678) *sink = s;
}
__global__ void k_chase(const unsigned int* __restrict__ next, int steps, unsigned int* sink, long long* cycles) {
    unsigned int p = 0;
    long long t0 = clock64();
    for (int i = 0; i < steps; i++) p = __ldcg(next + p);
    long long t1 = clock64();
    *sink = p;
    *cycles = t1 - t0;
}

static float best_ms(void (*launch)(void*), void* arg) {
    cudaEvent_t e0, e1;
    CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    launch(arg);
    CK(cudaDeviceSynchronize());
    float best = 1e30f;
    for (int r = 0; r < 5; r++) {
        CK(cudaEventRecord(e0));
        launch(arg);
        CK(cudaEventRecord(e1));
        CK(cudaEventSynchronize(e1));
        float ms;
        CK(cudaEventElapsedTime(&ms, e0, e1));
        if (ms < best) best = ms;
    }
    return best;
}

struct L2Arg { double2* buf; size_t n; int reps; double* sink; int grid; };
static void launch_l2(void* p) { L2Arg* a = (L2Arg*)p; k_l2_read<<<a->grid, 256>>>(a->buf, a->n, a->reps, a->sink); }
struct FArg { int iters; double* sink; int grid; };
static void launch_f(void* p) { FArg* a = (FArg*)p; k_fp64<<<a->grid, 256>>>(a->iters, a->sink); }

int main() {
    cudaDeviceProp prop;
    CK(cudaGetDeviceProperties(&prop, 0));
    const int sms = prop.multiProcessorCount;
    int clk_khz = 0;
    CK(cudaDeviceGetAttribute(&clk_khz, cudaDevAttrClockRate, 0));
    double* sink;
    CK(cudaMalloc(&sink, 64));
    // L2-resident read
    const size_t bytes = 48ull << 20;
    L2Arg a;
    CK(cudaMalloc(&a.buf, bytes));
    CK(cudaMemset(a.buf, 0, bytes));
    a.n = bytes / sizeof(double2); a.reps = 20; a.sink = sink; a.grid = sms * 8;
    float ms = best_ms(launch_l2, &a);
    double l2_gbs = (double)bytes * a.reps / (ms * 1e-3) / 1e9;
    // FP64 DADD+DMUL
    FArg f; f.iters = 20000; f.sink = sink; f.grid = sms * 8;
    float fms = best_ms(launch_f, &f);
    double ops = (double)f.grid * 256 * f.iters * 16.0;  // 8 DMUL + 8 DADD per iteration per thread
    double fp64_gops = ops / (fms * 1e-3) / 1e9;
    // dependent L2 latency
    const size_t nn = (64ull << 20) / 4;
    std::vector<unsigned int> h(nn);
    const size_t step = (1 << 20) / 4 + 33;  // > 1 MiB apart, co-prime with nn: every hop is a new line, ring covers the buffer
    for (size_t i = 0; i < nn; i++) h[i] = (unsigned int)((i + step) % nn);
    unsigned int* d_next; unsigned int* d_s; long long* d_c;
    CK(cudaMalloc(&d_next, nn * 4)); CK(cudaMalloc(&d_s, 4)); CK(cudaMalloc(&d_c, 8));
    CK(cudaMemcpy(d_next, h.data(), nn * 4, cudaMemcpyHostToDevice));
    k_chase<<<1, 1>>>(d_next, 2000, d_s, d_c);  // warm the ring into L2
    CK(cudaDeviceSynchronize());
    k_chase<<<1, 1>>>(d_next, 2000, d_s, d_c);
    CK(cudaDeviceSynchronize());
    long long cyc;
    CK(cudaMemcpy(&cyc, d_c, 8, cudaMemcpyDeviceToHost));
    printf("{\"device\": \"%s\", \"sms\": %d, \"sm_clock_mhz\": %.0f, \"l2_read_gbs\": %.1f, \"l2_buffer_mib\": 48, "
           "\"fp64_nofma_gops\": %.1f, \"fp64_nofma_ops_per_clk_per_sm\": %.1f, \"l2_dependent_load_cycles\": %.1f}\n",
           prop.name, sms, clk_khz / 1e3, l2_gbs, fp64_gops, fp64_gops * 1e9 / (sms * (clk_khz * 1e3)), (double)cyc / 2000.0);
    return 0;
}
