// fp64 pipe microbenchmark for B200 (sm_100a): dependent-issue latency of DFMA / DADD / DMUL and achieved fp64
// throughput as a function of resident warps per SM sub-partition and independent chains per thread.  Design input
// for k_advect: how much ILP x TLP the fp64 pipe needs to stay busy.
//   nvcc -O3 -gencode arch=compute_100a,code=sm_100a -o fp64_lat fp64_lat.cu && ./fp64_lat
#include <cstdio>
#include <cuda_runtime.h>

template <int ILP, int OP>
__global__ void chain(double* out, long long* cyc, int iters, double a, double b)
{
    double x[ILP];
#pragma unroll
    for (int i = 0; i < ILP; ++i) x[i] = a + i + threadIdx.x;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < 8; ++u) {
#pragma unroll
            for (int i = 0; i < ILP; ++i) {
                if (OP == 0) x[i] = fma(x[i], b, a);
                else if (OP == 1) x[i] = __dadd_rn(x[i], b);
                else x[i] = __dmul_rn(x[i], b);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < ILP; ++i) s += x[i];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) *cyc = t1 - t0;
}

template <int ILP, int OP>
void run(int warps_per_sm, const char* name)
{
    double* out; long long* cyc;
    cudaMalloc(&out, 148 * 2048 * 8); cudaMalloc(&cyc, 8);
    const int iters = 2000;
    chain<ILP, OP><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    chain<ILP, OP><<<148, warps_per_sm * 32>>>(out, cyc, iters, 1.0000001, 0.9999999);
    cudaDeviceSynchronize();
    long long h; cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double ops_per_warp = (double)iters * 8 * ILP;
    const double cyc_per_dep = (double)h / (iters * 8.0);                          // cycles per dependent step of one chain
    const double warp_inst_per_cyc_smsp = ops_per_warp * (warps_per_sm / 4.0) / h; // peak 0.5 (16 lanes/clk/SMSP)
    printf("%-5s warps/SM %2d (per SMSP %4.1f) ILP %d : %6.2f cycles per dependent op, fp64 pipe %5.1f %% of peak\n", name, warps_per_sm,
           warps_per_sm / 4.0, ILP, cyc_per_dep, 100.0 * warp_inst_per_cyc_smsp / 0.5);
    cudaFree(out); cudaFree(cyc);
}

int main()
{
    run<1, 0>(4, "DFMA"); run<1, 1>(4, "DADD"); run<1, 2>(4, "DMUL");
    for (int w : {4, 8, 12, 16, 24, 32}) {
        run<1, 0>(w, "DFMA"); run<2, 0>(w, "DFMA"); run<4, 0>(w, "DFMA"); run<6, 0>(w, "DFMA"); run<8, 0>(w, "DFMA");
    }
    return 0;
}
