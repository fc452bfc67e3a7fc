// Micro-benchmark (B200): latency of dependent FP64 operations, as the ray-generation recurrence issues them.
// nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_latency fp64_latency.cu && ./fp64_latency
#include <cstdio>
#include <cuda_runtime.h>
template <int MODE>
__global__ void chain(double* out, double c, double s, int n, long long* cycles) {
	double a = out[threadIdx.x], b = out[threadIdx.x + 32];
	const long long t0 = clock64();
	for (int i = 0; i < n; i++) {
		if (MODE == 0) { a = __dadd_rn(a, c); }                                   // dependent DADD
		if (MODE == 1) { a = __dmul_rn(a, c); }                                   // dependent DMUL
		if (MODE == 2) { a = __fma_rn(a, c, s); }                                 // dependent DFMA
		if (MODE == 3) {                                                          // the rotation step
			const double na = __dadd_rn(__dmul_rn(a, c), __dmul_rn(b, s)), nb = __dadd_rn(__dmul_rn(a, -s), __dmul_rn(b, c));
			a = na; b = nb;
		}
		if (MODE == 4) {                                                          // the rotation step + a store per step
			out[64 + (size_t)i * 4 + threadIdx.x * 8192] = a;
			const double na = __dadd_rn(__dmul_rn(a, c), __dmul_rn(b, s)), nb = __dadd_rn(__dmul_rn(a, -s), __dmul_rn(b, c));
			a = na; b = nb;
		}
	}
	const long long t1 = clock64();
	out[threadIdx.x] = a + b;
	if (threadIdx.x == 0) *cycles = t1 - t0;
}
int main() {
	double* d; long long* cyc; long long h;
	cudaMalloc(&d, (size_t)64 * 8192 * 8 + 4096); cudaMalloc(&cyc, 8);
	cudaMemset(d, 0, (size_t)64 * 8192 * 8 + 4096);
	const int n = 960;
	const char* names[] = {"dependent DADD", "dependent DMUL", "dependent DFMA", "rotation step (4 DMUL + 2 DADD)", "rotation step + store"};
	for (int lanes : {1, 3, 32}) {
		for (int m = 0; m < 5; m++) {
			for (int rep = 0; rep < 2; rep++) {
				if (m == 0) chain<0><<<1, lanes>>>(d, 0.9999, 1e-4, n, cyc);
				if (m == 1) chain<1><<<1, lanes>>>(d, 0.9999, 1e-4, n, cyc);
				if (m == 2) chain<2><<<1, lanes>>>(d, 0.9999, 1e-4, n, cyc);
				if (m == 3) chain<3><<<1, lanes>>>(d, 0.9999, 1e-4, n, cyc);
				if (m == 4) chain<4><<<1, lanes>>>(d, 0.9999, 1e-4, n, cyc);
				cudaDeviceSynchronize();
			}
			cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
			printf("%2d lanes  %-34s %7.1f cycles per step\n", lanes, names[m], (double)h / n);
		}
	}
	return 0;
}
