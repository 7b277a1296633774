// ubench_copy.cu -- what a block-sized PCIe copy costs on this box: back-to-back cudaMemcpyAsync of
// the cfg2 tuner block (819200 B, host -> device) and of its audio (524288 B, device -> host), each
// alone and both directions at once, plus the same bytes split over two streams.
// Build: nvcc -O2 -o build/ubench_copy tools/ubench_copy.cu ; run on the GPU box.
#include <cstdio>
#include <cuda_runtime.h>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

int main()
{
	const size_t inB = 819200, outB = 524288;
	const int N = 2000, NB = 8;
	char *h_in, *h_out, *d_in, *d_out;
	CK(cudaMallocHost(&h_in, inB * NB));
	CK(cudaMallocHost(&h_out, outB * NB));
	CK(cudaMalloc(&d_in, inB * NB));
	CK(cudaMalloc(&d_out, outB * NB));
	cudaStream_t s1, s2, s3;
	CK(cudaStreamCreateWithFlags(&s1, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&s2, cudaStreamNonBlocking));
	CK(cudaStreamCreateWithFlags(&s3, cudaStreamNonBlocking));
	cudaEvent_t e0, e1, f0, f1;
	CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1)); CK(cudaEventCreate(&f0)); CK(cudaEventCreate(&f1));
	float ms, ms2;
	for (int pass = 0; pass < 2; pass++) {
		// host -> device alone
		CK(cudaEventRecord(e0, s1));
		for (int i = 0; i < N; i++)
			CK(cudaMemcpyAsync(d_in + (i % NB) * inB, h_in + (i % NB) * inB, inB, cudaMemcpyHostToDevice, s1));
		CK(cudaEventRecord(e1, s1));
		CK(cudaEventSynchronize(e1));
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (pass) printf("H2D %zu B alone          : %.2f us per copy, %.1f GB/s\n", inB, ms * 1e3 / N, inB * N / ms / 1e6);
		// device -> host alone
		CK(cudaEventRecord(e0, s2));
		for (int i = 0; i < N; i++)
			CK(cudaMemcpyAsync(h_out + (i % NB) * outB, d_out + (i % NB) * outB, outB, cudaMemcpyDeviceToHost, s2));
		CK(cudaEventRecord(e1, s2));
		CK(cudaEventSynchronize(e1));
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (pass) printf("D2H %zu B alone          : %.2f us per copy, %.1f GB/s\n", outB, ms * 1e3 / N, outB * N / ms / 1e6);
		// both directions at once
		CK(cudaEventRecord(e0, s1));
		CK(cudaEventRecord(f0, s2));
		for (int i = 0; i < N; i++) {
			CK(cudaMemcpyAsync(d_in + (i % NB) * inB, h_in + (i % NB) * inB, inB, cudaMemcpyHostToDevice, s1));
			CK(cudaMemcpyAsync(h_out + (i % NB) * outB, d_out + (i % NB) * outB, outB, cudaMemcpyDeviceToHost, s2));
		}
		CK(cudaEventRecord(e1, s1));
		CK(cudaEventRecord(f1, s2));
		CK(cudaEventSynchronize(e1));
		CK(cudaEventSynchronize(f1));
		CK(cudaEventElapsedTime(&ms, e0, e1));
		CK(cudaEventElapsedTime(&ms2, f0, f1));
		if (pass) printf("both directions at once  : H2D %.2f us per copy, D2H %.2f us per copy\n", ms * 1e3 / N, ms2 * 1e3 / N);
		// host -> device split over two streams
		CK(cudaEventRecord(e0, s1));
		CK(cudaEventRecord(f0, s3));
		for (int i = 0; i < N; i++) {
			CK(cudaMemcpyAsync(d_in + (i % NB) * inB, h_in + (i % NB) * inB, inB / 2, cudaMemcpyHostToDevice, s1));
			CK(cudaMemcpyAsync(d_in + (i % NB) * inB + inB / 2, h_in + (i % NB) * inB + inB / 2, inB / 2, cudaMemcpyHostToDevice, s3));
		}
		CK(cudaEventRecord(e1, s1));
		CK(cudaEventRecord(f1, s3));
		CK(cudaEventSynchronize(e1));
		CK(cudaEventSynchronize(f1));
		CK(cudaEventElapsedTime(&ms, e0, e1));
		CK(cudaEventElapsedTime(&ms2, f0, f1));
		if (pass) printf("H2D halves on two streams: %.2f / %.2f us per block\n", ms * 1e3 / N, ms2 * 1e3 / N);
		// one large copy for reference
		CK(cudaEventRecord(e0, s1));
		for (int i = 0; i < 20; i++)
			CK(cudaMemcpyAsync(d_in, h_in, inB * NB, cudaMemcpyHostToDevice, s1));
		CK(cudaEventRecord(e1, s1));
		CK(cudaEventSynchronize(e1));
		CK(cudaEventElapsedTime(&ms, e0, e1));
		if (pass) printf("H2D %zu B copies        : %.1f GB/s\n", inB * NB, inB * NB * 20 / ms / 1e6);
	}
	return 0;
}
